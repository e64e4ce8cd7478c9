import time

import pytest
import torch
import torch.multiprocessing as mp

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.channel import QueueTimeoutError, ShmChannel


def _sender(ch, rank, n, out_q):
  for i in range(n):
    ch.send({'ids': torch.full((100 + i,), rank * 1000 + i, dtype=torch.int64),
             'feat': torch.full((10, 4), float(i)), '#META.k': torch.tensor([rank, i])})
  out_q.put(('sent', rank))


def _receiver(ch, n, out_q):
  got = []
  for _ in range(n):
    m = ch.recv(timeout_ms=20000)
    r, i = m['#META.k'].tolist()
    assert m['ids'].shape == (100 + i,) and int(m['ids'][0]) == r * 1000 + i
    assert float(m['feat'][0, 0]) == float(i)
    got.append((r, i))
  out_q.put(('got', got))


def test_shm_channel_roundtrip_and_timeout():
  ch = ShmChannel(capacity=4, shm_size='1MB')
  msg = {'a': torch.arange(10), 'b': torch.randn(3, 5), 'c': torch.tensor([1.5], dtype=torch.float64),
         'd': torch.zeros(0, dtype=torch.int64), 'e': torch.arange(6, dtype=torch.int16).view(2, 3)}
  ch.send(msg)
  assert not ch.empty()
  out = ch.recv(timeout_ms=1000)
  assert set(out.keys()) == set(msg.keys())
  for k in msg:
    assert out[k].dtype == msg[k].dtype and torch.equal(out[k], msg[k])
  assert ch.empty()
  del out   # ring space is reclaimed in order: a held message pins everything behind it
  t0 = time.time()
  with pytest.raises(QueueTimeoutError):
    ch.recv(timeout_ms=200)
  assert 0.15 < time.time() - t0 < 2.0
  # ring wrap-around: many messages through a small ring, blocks recycled when tensors die
  for i in range(200):
    ch.send({'x': torch.full((5000,), i, dtype=torch.int32)})
    assert int(ch.recv(timeout_ms=1000)['x'][-1]) == i


def test_shm_channel_multi_process_exactly_once():
  """2 senders + 2 receivers in spawned processes: every message arrives exactly once
  (same scenario as the reference's test/cpp/test_shm_queue.cu:72-145)."""
  ctx = mp.get_context('spawn')
  ch = ShmChannel(capacity=8, shm_size='4MB')
  q = ctx.Queue()
  n = 50
  procs = [ctx.Process(target=_sender, args=(ch, r, n, q)) for r in range(2)]
  procs += [ctx.Process(target=_receiver, args=(ch, n, q)) for _ in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=120) for _ in range(4)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0          # the reference never checks child exit codes
  got = sorted(x for kind, v in results if kind == 'got' for x in v)
  assert got == sorted((r, i) for r in range(2) for i in range(n))


def test_message_too_large_is_an_error():
  ch = ShmChannel(capacity=2, shm_size='64KB')
  with pytest.raises(RuntimeError):
    ch.send({'x': torch.zeros(1 << 20)})


@pytest.mark.gpu
def test_shm_channel_cuda_tensors_and_pinning():
  ch = ShmChannel(capacity=4, shm_size='8MB')
  ch.pin_memory()
  x = torch.randn(1000, 16, device='cuda')
  ch.send({'x': x, 'ids': torch.arange(7, device='cuda')})
  out = ch.recv(timeout_ms=1000)
  assert out['x'].device.type == 'cpu' and torch.equal(out['x'], x.cpu())
  assert torch.equal(out['ids'], torch.arange(7))
