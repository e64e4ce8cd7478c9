from .base import ChannelBase, SampleMessage, QueueTimeoutError, QueueClosedError
from .mp_channel import MpChannel
from .shm_channel import ShmChannel
from .remote_channel import RemoteReceivingChannel
