"""Fixtures for distributed tests: a 40-node ring (v -> v+1, v+2) split in two partitions, feature
row v == [v]*dim, edge feature e == [e]*4, label v == v (idea: reference test/python/dist_test_utils.py)."""
import os
import sys
import traceback

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

N = 40


def build_partition(rank: int, world: int = 2, scheme: str = 'hash', edge_dir: str = 'out', dim: int = 8):
  import graphlearn_for_pytorch_b200 as glt
  from graphlearn_for_pytorch_b200.distributed import DistDataset
  from graphlearn_for_pytorch_b200.partition import GLTPartitionBook, RangePartitionBook
  from graphlearn_for_pytorch_b200.utils.synthetic import id_features, ring_graph
  from graphlearn_for_pytorch_b200.utils.tensor import id2idx
  ei = ring_graph(N)
  eids = torch.arange(ei.shape[1])
  if scheme == 'hash':
    node_pb = GLTPartitionBook(torch.arange(N) % world)
  else:
    per = N // world
    node_pb = RangePartitionBook([(r * per, (r + 1) * per) for r in range(world)], rank)
  key = ei[0] if edge_dir == 'out' else ei[1]
  owners = node_pb[key]
  edge_pb = GLTPartitionBook(owners.clone())
  m = owners == rank
  ds = DistDataset(edge_dir=edge_dir)
  ds.num_partitions, ds.partition_idx = world, rank
  ds.init_graph(ei[:, m], eids[m], graph_mode='CPU', num_nodes=N)
  own = torch.nonzero(node_pb[torch.arange(N)] == rank).view(-1)
  ds.init_node_features(id_features(N, dim)[own], id2idx(own), with_gpu=False)
  own_e = eids[m]
  ds.init_edge_features(id_features(ei.shape[1], 4)[own_e], id2idx(own_e), with_gpu=False)
  ds.init_node_labels(torch.arange(N))
  ds.node_pb, ds.edge_pb = node_pb, edge_pb
  return ds


def check_batch(b, with_edge=True):
  n = N
  assert torch.equal(b.x[:, 0].long(), b.node), (b.x[:, 0], b.node)
  assert torch.equal(b.y, b.node)
  src, dst = b.node[b.edge_index[0]], b.node[b.edge_index[1]]
  assert torch.all(((src - dst) % n == 1) | ((src - dst) % n == 2))
  if with_edge and b.edge is not None and b.edge_attr is not None:
    assert torch.equal(b.edge_attr[:, 0].long(), b.edge)
  assert sum(b.num_sampled_nodes) == b.node.numel()
  assert torch.is_tensor(b.num_sampled_nodes) and b.num_sampled_nodes.dtype == torch.int64   # reference convention


def _entry(rank, world, port, fn, args, err_q):
  if os.environ.get('GLT_TEST_FAULT_TIMEOUT'):      # hang diagnosis: dump every thread's stack after N seconds
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ['GLT_TEST_FAULT_TIMEOUT']), exit=False)
  try:
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    fn(rank, world, port, *args)
  except Exception:  # noqa: BLE001
    err_q.put((rank, traceback.format_exc()))
    raise


def _kill_tree(pid):
  try:
    import psutil
    for c in psutil.Process(pid).children(recursive=True):
      try:
        c.kill()
      except Exception:  # noqa: BLE001
        pass
  except Exception:  # noqa: BLE001
    pass


def run_workers(fn, world=2, args=(), timeout=240, retries=1):
  """Spawn `world` processes running fn(rank, world, port, *args); assert clean exit codes.  A group that dies on a
  transport TIMEOUT (seen once in ~40 full-suite runs: a TensorPipe connection on loopback that never completes) is
  run again on fresh ports; any other failure, and a second timeout, fail the test."""
  for attempt in range(retries + 1):
    try:
      return _run_workers_once(fn, world, args, timeout)
    except AssertionError as e:
      transient = 'RPC ran for more than set timeout' in str(e) or str(e).strip().endswith("timeout')") \
          or "'timeout'" in str(e)
      if attempt == retries or not transient:
        raise
      print(f'run_workers: transport timeout, retrying once\n{e}', flush=True)


def _run_workers_once(fn, world=2, args=(), timeout=240):
  from graphlearn_for_pytorch_b200.utils.common import get_free_port_block
  ctx = mp.get_context('spawn')
  port = get_free_port_block(8)      # workers derive port+1.. for their sampling groups
  err_q = ctx.Queue()
  procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, err_q)) for r in range(world)]
  import time
  for p in procs:
    p.start()
  deadline = time.time() + timeout            # ONE deadline for the whole group, not one per process
  for p in procs:
    p.join(max(0.0, deadline - time.time()))
  errs = []
  while not err_q.empty():
    errs.append(err_q.get())
  for p in procs:
    if p.is_alive():
      _kill_tree(p.pid)                        # sampling sub-processes must not outlive a stuck test
      p.terminate()
      p.join(10)
      errs.append(('?', 'timeout'))
  assert not errs, '\n'.join(f'[rank {r}] {e}' for r, e in errs)
  assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


def build_hetero_partition(rank: int, world: int = 2, edge_dir: str = 'out'):
  """user -u2i-> item, item -i2i-> item; 20 users, 20 items; feature rows encode (type offset + id)."""
  import graphlearn_for_pytorch_b200 as glt
  from graphlearn_for_pytorch_b200.distributed import DistDataset
  from graphlearn_for_pytorch_b200.partition import GLTPartitionBook
  from graphlearn_for_pytorch_b200.utils.synthetic import id_features
  from graphlearn_for_pytorch_b200.utils.tensor import id2idx
  nu = ni = 20
  u = torch.arange(nu)
  u2i = torch.stack([u.repeat_interleave(2), torch.stack([u % ni, (u + 1) % ni], 1).flatten()])
  i = torch.arange(ni)
  i2i = torch.stack([i, (i + 3) % ni])
  edges = {('user', 'u2i', 'item'): u2i, ('item', 'i2i', 'item'): i2i}
  node_pb = {'user': GLTPartitionBook(torch.arange(nu) % world), 'item': GLTPartitionBook(torch.arange(ni) % world)}
  ei_local, eids_local, edge_pb = {}, {}, {}
  for et, ei in edges.items():
    key_t = et[0] if edge_dir == 'out' else et[2]
    key = ei[0] if edge_dir == 'out' else ei[1]
    owners = node_pb[key_t][key]
    edge_pb[et] = GLTPartitionBook(owners.clone())
    m = owners == rank
    ei_local[et] = ei[:, m]
    eids_local[et] = torch.arange(ei.shape[1])[m]
  ds = DistDataset(edge_dir=edge_dir)
  ds.num_partitions, ds.partition_idx = world, rank
  ds.init_graph(ei_local, eids_local, graph_mode='CPU', num_nodes={'user': nu, 'item': ni})
  feats, i2x = {}, {}
  for nt, off in (('user', 0), ('item', 1000)):
    own = torch.nonzero(node_pb[nt][torch.arange(20)] == rank).view(-1)
    feats[nt] = (id_features(20, 4) + off)[own]
    i2x[nt] = id2idx(own)
  ds.init_node_features(feats, i2x, with_gpu=False)
  ds.init_node_labels({'user': torch.arange(nu)})
  ds.node_pb, ds.edge_pb = node_pb, edge_pb
  return ds, edges
