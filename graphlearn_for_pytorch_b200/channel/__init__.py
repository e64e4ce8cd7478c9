"""Sample transport between sampling workers and trainers: shared-memory ring (native), multiprocessing queue,
and the client side of the server-client mode."""
from . import base, mp_channel, remote_channel, shm_channel

ChannelBase = base.ChannelBase
SampleMessage = base.SampleMessage
QueueTimeoutError = base.QueueTimeoutError
QueueClosedError = base.QueueClosedError
MpChannel = mp_channel.MpChannel
ShmChannel = shm_channel.ShmChannel
RemoteReceivingChannel = remote_channel.RemoteReceivingChannel

__all__ = ['ChannelBase', 'SampleMessage', 'QueueTimeoutError', 'QueueClosedError', 'MpChannel', 'ShmChannel',
           'RemoteReceivingChannel']
