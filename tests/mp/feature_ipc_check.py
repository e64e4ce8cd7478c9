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

  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_reader, args=(r, feat, graph, ut_handle, devices, ids, x[ids], q))
           for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(60)
  for rank, ok_feat, ok_ut, ok_graph, err in res:
    assert ok_feat and ok_ut and ok_graph, f'reader {rank}: feat={ok_feat} ut={ok_ut} graph={ok_graph}\n{err}'


if __name__ == '__main__':
  main()
  print('IPC_OK', flush=True)
