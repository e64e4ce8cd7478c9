"""temporary: find which object fails to travel to a spawned process on the GPU box"""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.multiprocessing as mp


def child(name, obj, q):
  q.put((name, 'ok'))


def try_send(ctx, name, obj):
  q = ctx.Queue()
  p = ctx.Process(target=child, args=(name, obj, q))
  p.start()
  try:
    print(q.get(timeout=60), flush=True)
  except Exception as e:
    print((name, 'FAILED', repr(e)), flush=True)
  p.join(30)


if __name__ == '__main__':
  import graphlearn_for_pytorch_b200 as glt
  os.system('df -h /dev/shm | tail -1; mount | grep shm')
  n, d = 4096, 64
  x = torch.randn(n, d)
  ctx = mp.get_context('spawn')
  try_send(ctx, 'plain', torch.randn(1000, 64))
  try_send(ctx, 'shared', torch.randn(1000, 64).share_memory_())
  torch.cuda.init()
  y = torch.randn(10, device='cuda')
  try_send(ctx, 'plain_after_cuda', torch.randn(1000, 64))
  try_send(ctx, 'shared_after_cuda', torch.randn(1000, 64).share_memory_())
  pinned = torch.randn(1000, 64).pin_memory()
  try_send(ctx, 'index_result', x[torch.randint(0, n, (1000,))])
  perm = torch.randperm(n)
  id2index = torch.empty(n, dtype=torch.int64); id2index[perm] = torch.arange(n)
  feat = glt.data.Feature(x[perm], id2index, split_ratio=0.5, device=0)
  ids = torch.randint(0, n, (1000,))
  feat[ids]
  h = feat.share_ipc()
  for i, part in enumerate(h):
    try_send(ctx, f'feat_handle[{i}]', part)
  try_send(ctx, 'feat', feat)
  ut = glt.data.UnifiedTensor(0, torch.float32)
  ut.init_from([x[:1024], x[1024:]], [0, -1])
  uh = ut.share_ipc()
  try_send(ctx, 'ut_handle', uh)
  row = torch.randint(0, n, (n * 8,)); col = torch.randint(0, n, (n * 8,))
  topo = glt.data.Topology(torch.stack([row, col]), layout='CSC')
  try_send(ctx, 'topo', topo)
  graph = glt.data.Graph(topo, 'ZERO_COPY', 0)
  try_send(ctx, 'graph', graph)
