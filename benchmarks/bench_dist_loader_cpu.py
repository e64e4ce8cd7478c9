"""Distributed neighbour loader on the RPC plane, CPU mode, one box: this library vs the unmodified reference
(baseline/_ref) on the SAME partition directory (the on-disk format is interchangeable, tools/partition_format_compat.py).

Two trainer processes (one per partition) over localhost RPC, each with `--workers` sampling sub-processes
(`MpDistSamplingWorkerOptions`), graph and features in host memory, fan-out [15,10,5], batch 1024, node features
collected.  The metric is the one of the reference's benchmarks/api/bench_dist_neighbor_loader.py:139-162: sampled
nodes(+feature rows) and edges per second over a whole epoch, summed over the trainers (epoch 0 is warm-up).

  python benchmarks/bench_dist_loader_cpu.py ours | reference [--nodes N --edges E --workers W --epochs K]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FANOUT, BATCH = [15, 10, 5], 1024


def _import(impl):
  if impl == 'reference':
    sys.path.insert(0, os.path.join(ROOT, 'baseline', 'shims'))
    sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
    import graphlearn_torch as glt
  else:
    sys.path.insert(0, ROOT)
    import graphlearn_for_pytorch_b200 as glt
  return glt


def _trainer(rank, world, impl, root, port, workers, epochs, ret):
  glt = _import(impl)
  gd = glt.distributed
  torch.set_num_threads(2)
  gd.init_worker_group(world, rank, 'bench_dist_cpu')
  gd.init_rpc(master_addr='127.0.0.1', master_port=port, num_rpc_threads=2, rpc_timeout=180)   # trainer group
  ds = gd.DistDataset()
  ds.load(root, rank, graph_mode='CPU', feature_with_gpu=False,
          whole_node_label_file=os.path.join(root, 'labels.pt'))
  train = torch.load(os.path.join(root, 'train_idx.pt'))
  train = train[ds.node_pb[train] == rank]
  opts = gd.MpDistSamplingWorkerOptions(num_workers=workers, worker_devices=[torch.device('cpu')] * workers,
                                        worker_concurrency=4, master_addr='127.0.0.1', master_port=port + 1,
                                        channel_size='1GB', pin_memory=False)
  loader = gd.DistNeighborLoader(ds, FANOUT, train, batch_size=BATCH, shuffle=True, drop_last=False,
                                 collect_features=True, to_device=torch.device('cpu'), worker_options=opts)
  stats = []
  for _ in range(epochs):
    gd.barrier()
    t0, nodes, edges, nb = time.time(), 0, 0, 0
    for b in loader:
      nodes += int(b.node.numel())
      edges += int(b.edge_index.shape[1])
      assert b.x.shape[0] == b.node.numel()
      nb += 1
    gd.barrier()
    stats.append((time.time() - t0, nodes, edges, nb))
  ret.put((rank, stats))
  loader.shutdown()
  gd.shutdown_rpc()


def main():
  p = argparse.ArgumentParser()
  p.add_argument('impl', choices=['ours', 'reference'])
  p.add_argument('--nodes', type=int, default=400_000)
  p.add_argument('--edges', type=int, default=8_000_000)
  p.add_argument('--feat-dim', type=int, default=100)
  p.add_argument('--train-frac', type=float, default=0.1)
  p.add_argument('--workers', type=int, default=2)
  p.add_argument('--epochs', type=int, default=3)
  p.add_argument('--root', default=None, help='reuse an existing partition directory')
  args = p.parse_args()
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  root = args.root or tempfile.mkdtemp(prefix='glt_bench_dist_')
  if not os.path.exists(os.path.join(root, 'META')):
    # always partitioned with THIS library (both arms read the same files)
    sys.path.insert(0, ROOT)
    from graphlearn_for_pytorch_b200.partition import RandomPartitioner
    g = torch.Generator().manual_seed(0)
    N, E = args.nodes, args.edges
    src = torch.randint(0, N, (E // 2,), generator=g)
    dst = (src + torch.randint(1, 5000, (E // 2,), generator=g)) % N
    ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
    x = torch.randn(N, args.feat_dim, generator=g)
    torch.save(torch.randint(0, 47, (N,), generator=g), os.path.join(root, 'labels.pt'))
    torch.save(torch.randperm(N, generator=g)[: int(N * args.train_frac)], os.path.join(root, 'train_idx.pt'))
    RandomPartitioner(root, 2, N, ei, node_feat=x).partition()
  sys.path.insert(0, ROOT)
  from graphlearn_for_pytorch_b200.utils import get_free_port
  port = get_free_port()
  os.environ['MASTER_PORT'] = str(port)
  ctx = mp.get_context('spawn')
  ret = ctx.Queue()
  procs = [ctx.Process(target=_trainer, args=(r, 2, args.impl, root, port, args.workers, args.epochs, ret))
           for r in range(2)]
  for pr in procs:
    pr.start()
  out = {}
  while len(out) < len(procs):
    try:
      r, st = ret.get(timeout=5)
      out[r] = st
    except Exception:
      if any(pr.exitcode not in (None, 0) for pr in procs):
        for pr in procs:
          pr.kill()
        raise SystemExit('a trainer process failed')
  for pr in procs:
    pr.join(timeout=120)
  epochs = []
  for e in range(1, args.epochs):       # epoch 0 = warm-up (worker start, first touches)
    dt = max(out[r][e][0] for r in out)
    epochs.append({'s': dt, 'M_nodes_per_s': sum(out[r][e][1] for r in out) / dt / 1e6,
                   'M_edges_per_s': sum(out[r][e][2] for r in out) / dt / 1e6,
                   'batches': sum(out[r][e][3] for r in out)})
  best = max(epochs, key=lambda d: d['M_edges_per_s'])
  print(json.dumps({'impl': args.impl, 'mode': 'cpu / 2 partitions / localhost RPC', 'workers_per_trainer': args.workers,
                    'nodes': args.nodes, 'edges': args.edges, 'feat_dim': args.feat_dim, 'batch': BATCH,
                    'fanout': FANOUT, 'best_epoch': best, 'epochs': epochs}))
  if args.root is None:
    shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
  main()
