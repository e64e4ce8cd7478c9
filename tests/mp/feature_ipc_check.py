"""Feature / UnifiedTensor / Graph handed to spawned reader processes (one per GPU when two are
visible, else two readers on GPU 0).  Run as a plain script:  python tests/mp/feature_ipc_check.py
Reference test strategy: test/python/test_feature.py spawns a reader per GPU; test_graph.py checks
CUDA / ZERO_COPY IPC."""
import os
import sys

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _reader(rank, feat, graph, ut_handle, devices, ids, expect, q):
  import graphlearn_for_pytorch_b200 as glt
  dev = devices[rank % len(devices)]
  torch.cuda.set_device(dev)
  try:
    feat.device = dev
    got = feat[ids].cpu()
    ok_feat = bool(torch.equal(got, expect))
    ut = glt.data.UnifiedTensor.new_from_ipc(ut_handle, dev, torch.float32)
    ok_ut = bool(torch.equal(ut[ids].cpu(), expect))
    graph.device = dev
    sampler = glt.sampler.NeighborSampler(graph, [3, 2], device=torch.device('cuda', dev), seed=7)
    out = sampler.sample_from_nodes(torch.arange(16))
    indptr, indices = graph.topo.indptr, graph.topo.indices
    node = out.node.cpu()
    src, dst = node[out.row.cpu()], node[out.col.cpu()]
    ok_graph = all(int(s) in indices[indptr[int(d)]:indptr[int(d) + 1]].tolist()
                   for s, d in zip(src.tolist(), dst.tolist())) and out.row.numel() > 0
    q.put((rank, ok_feat, ok_ut, ok_graph, ''))
  except Exception as e:  # surface the failure in the parent
    import traceback
    q.put((rank, False, False, False, traceback.format_exc()))


def main():
  import graphlearn_for_pytorch_b200 as glt
  ndev = min(2, torch.cuda.device_count())
  devices = list(range(ndev))
  n, d = 4096, 64
  x = torch.randn(n, d)
  perm = torch.randperm(n)
  id2index = torch.empty(n, dtype=torch.int64)
  id2index[perm] = torch.arange(n)
  feat = glt.data.Feature(x[perm], id2index, split_ratio=0.5,
                          device_group_list=[glt.data.DeviceGroup(0, devices)], device=0)
  ids = torch.randint(0, n, (1000,))
  assert torch.equal(feat[ids].cpu(), x[ids])

  ut = glt.data.UnifiedTensor(0, torch.float32)
  ut.init_from([x[:1024], x[1024:2048] if ndev > 1 else None, x[2048:] if ndev > 1 else x[1024:]],
               [0, 1 if ndev > 1 else 0, -1])
  assert torch.equal(ut[ids].cpu(), x[ids])
  ut_handle = ut.share_ipc()
  assert torch.equal(ut[ids].cpu(), x[ids])  # re-homed parts still serve the owner

  row = torch.randint(0, n, (n * 8,))
  col = torch.randint(0, n, (n * 8,))
  topo = glt.data.Topology(torch.stack([row, col]), layout='CSC')
  graph = glt.data.Graph(topo, 'ZERO_COPY', 0)

  mode = os.environ.get('IPC_MODE', 'conc')          # conc | seq   (debug knobs)
  objs = os.environ.get('IPC_OBJS', 'all')           # all | feat | graph | ut | plain
  if os.environ.get('IPC_PRESHARE', '0') == '1':
    feat.share_ipc(); graph.share_ipc()
  if objs != 'all':
    small = glt.data.Feature(x[perm][:64].clone(), None, split_ratio=0.5, device=0)
    small[torch.arange(4)]
    if objs != 'feat':
      feat = small
    if objs != 'graph':
      graph = glt.data.Graph(glt.data.Topology(torch.stack([row[:64] % 64, col[:64] % 64]), layout='CSC'), 'CPU', 0)
    if objs != 'ut':
      u2 = glt.data.UnifiedTensor(0, torch.float32); u2.init_from([x[:8], x[8:16]], [0, -1]); ut_handle = u2.share_ipc()
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_reader, args=(r, feat, graph, ut_handle, devices, ids, x[ids], q))
           for r in range(2)]
  import queue as _queue
  import time

  def wait_results(ps, n):
    out, t0 = [], time.time()
    while len(out) < n and time.time() - t0 < 100:
      try:
        out.append(q.get(timeout=1))
      except _queue.Empty:
        dead = [p_ for p_ in ps if p_.exitcode not in (None, 0)]
        if dead:   # a reader died before reporting (e.g. while unpickling its arguments)
          raise RuntimeError(f'reader process exited with code {dead[0].exitcode} before reporting')
    if len(out) < n:
      raise RuntimeError('timed out waiting for the reader processes')
    return out

  res = []
  if mode == 'seq':
    for p in procs:
      p.start()
      res += wait_results([p], 1)
  else:
    for p in procs:
      p.start()
    res = wait_results(procs, len(procs))
  for p in procs:
    p.join(60)
  if objs != 'all':
    print('variant finished', mode, objs, [r[4][-300:] for r in res], flush=True)
    return
  for rank, ok_feat, ok_ut, ok_graph, err in res:
    assert ok_feat and ok_ut and ok_graph, f'reader {rank}: feat={ok_feat} ut={ok_ut} graph={ok_graph}\n{err}'


if __name__ == '__main__':
  main()
  print('IPC_OK', flush=True)
