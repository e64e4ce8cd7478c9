"""Distributed R-GNN training on partitioned IGBH (MLPerf-GNN style; counterpart of the reference's
examples/igbh/dist_train_rgnn.py): one process per partition, cross-partition hetero neighbour sampling through
`DistNeighborLoader`, DDP model, validation accuracy with early stop at --target_acc, mllog events, checkpoints.

  python examples/igbh/partition.py --src_path D --dst_path P --num_partitions 2
  for r in 0 1; do python examples/igbh/dist_train_rgnn.py --path P --rank $r --world 2 & done
Runs on CPU (gloo) as well as on GPUs (nccl); across machines set --master_addr.
"""
import argparse
import os
import os.path as osp
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, osp.dirname(osp.abspath(__file__)))
from common import glt  # noqa: E402,F401
import graphlearn_for_pytorch_b200.distributed as gd  # noqa: E402
from graphlearn_for_pytorch_b200.models import RGNN  # noqa: E402
from graphlearn_for_pytorch_b200.utils import load_ckpt, save_ckpt  # noqa: E402
from mlperf_logging_utils import get_mlperf_logger, submission_info  # noqa: E402


def evaluate(model, loader, device, with_trim=True):
  model.eval()
  correct = total = 0
  with torch.no_grad():
    for b in loader:
      bs = b['paper'].batch_size
      out = model(b.x_dict, b.edge_index_dict, *((b.num_sampled_nodes, b.num_sampled_edges) if with_trim else ()))[:bs]
      y = b['paper'].y[:bs].to(device)
      correct += int((out.argmax(1) == y).sum())
      total += bs
  t = torch.tensor([correct, total], dtype=torch.float64, device=device)
  dist.all_reduce(t)
  model.train()
  return float(t[0] / t[1].clamp(min=1))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--path', required=True, help='partition root written by partition.py')
  ap.add_argument('--rank', type=int, required=True)
  ap.add_argument('--world', type=int, default=2)
  ap.add_argument('--master_addr', default='127.0.0.1')
  ap.add_argument('--master_port', type=int, default=29800)
  ap.add_argument('--model', default='rgat', choices=['rgat', 'rsage', 'rgcn'])
  ap.add_argument('--fan_out', default='15,10,5')
  ap.add_argument('--batch_size', type=int, default=512)
  ap.add_argument('--hidden_channels', type=int, default=128)
  ap.add_argument('--learning_rate', type=float, default=1e-3)
  ap.add_argument('--epochs', type=int, default=2)
  ap.add_argument('--edge_dir', default='in', choices=['in', 'out'])
  ap.add_argument('--num_workers', type=int, default=0, help='sampling sub-processes per trainer (0 = collocated)')
  ap.add_argument('--target_acc', type=float, default=0.72)
  ap.add_argument('--ckpt_steps', type=int, default=-1)
  ap.add_argument('--ckpt_path', default=None)
  ap.add_argument('--max_steps', type=int, default=-1, help='stop an epoch early (smoke tests)')
  # options of the reference's MLPerf runs (examples/igbh/dist_train_rgnn.py:395-478)
  ap.add_argument('--num_heads', type=int, default=4, help='attention heads (rgat)')
  ap.add_argument('--train_batch_size', type=int, default=None, help='overrides --batch_size for training')
  ap.add_argument('--val_batch_size', type=int, default=None, help='overrides --batch_size for validation')
  ap.add_argument('--validation_acc', type=float, default=None, help='alias of --target_acc')
  ap.add_argument('--validation_frac_within_epoch', type=float, default=1.0,
                  help='validate every this fraction of an epoch (MLPerf: 0.05) and stop once the target is reached')
  ap.add_argument('--evaluate_on_epoch_end', type=int, default=1)
  ap.add_argument('--checkpoint_on_epoch_end', type=int, default=1)
  ap.add_argument('--random_seed', type=int, default=42)
  ap.add_argument('--rpc_timeout', type=float, default=180)
  ap.add_argument('--graph_caching', action='store_true', help='partitions carry the full topology (partition.py)')
  ap.add_argument('--with_trim', type=int, default=1, help='per-layer trimming by num_sampled_nodes/edges')
  ap.add_argument('--use_all2all', action='store_true', help='exchange remote features with collectives, not RPC')
  ap.add_argument('--precision', default='fp32', choices=['fp32', 'bf16'], help='bf16: autocast forward/backward')
  ap.add_argument('--cpu_mode', action='store_true', help='train on the CPU even when GPUs are present')
  a = ap.parse_args()
  if a.validation_acc is not None:
    a.target_acc = a.validation_acc
  torch.manual_seed(a.random_seed)
  glt.utils.RandomSeedManager.set_seed(a.random_seed)

  cuda = torch.cuda.is_available() and not a.cpu_mode
  device = torch.device('cuda', a.rank % max(torch.cuda.device_count(), 1)) if cuda else torch.device('cpu')
  os.environ.setdefault('MASTER_ADDR', a.master_addr)
  os.environ.setdefault('MASTER_PORT', str(a.master_port))
  log = get_mlperf_logger(rank=a.rank)
  submission_info(log)
  log.start('INIT')
  dist.init_process_group('nccl' if cuda else 'gloo', rank=a.rank, world_size=a.world)
  gd.init_worker_group(a.world, a.rank)
  ds = gd.DistDataset(edge_dir=a.edge_dir)
  ds.load(a.path, a.rank, graph_mode='CUDA' if cuda else 'CPU', feature_with_gpu=cuda, graph_caching=a.graph_caching,
          whole_node_label_file={'paper': osp.join(a.path, 'label.pt')}, device=device.index)
  train_idx = torch.load(osp.join(a.path, 'train_idx.pt'))
  val_idx = torch.load(osp.join(a.path, 'val_idx.pt'))
  # equal shares per rank (seeds need not be local: the loader samples across partitions), so every DDP rank
  # runs the same number of steps
  def share(idx):
    per = idx.numel() // a.world
    return idx[a.rank * per:(a.rank + 1) * per]
  train_idx, val_idx = share(train_idx), share(val_idx)
  fan = [int(v) for v in a.fan_out.split(',')]

  def opts(port):
    if a.num_workers > 0:
      return gd.MpDistSamplingWorkerOptions(num_workers=a.num_workers, worker_concurrency=4, master_addr=a.master_addr,
                                            master_port=port, pin_memory=cuda, rpc_timeout=a.rpc_timeout,
                                            use_all2all=a.use_all2all)
    return gd.CollocatedDistSamplingWorkerOptions(master_addr=a.master_addr, master_port=port,
                                                  rpc_timeout=a.rpc_timeout, use_all2all=a.use_all2all)
  train_bs, val_bs = a.train_batch_size or a.batch_size, a.val_batch_size or a.batch_size
  train_loader = gd.DistNeighborLoader(ds, fan, ('paper', train_idx), batch_size=train_bs, shuffle=True,
                                       drop_last=False, collect_features=True, to_device=device, edge_dir=a.edge_dir,
                                       worker_options=opts(a.master_port + 1))
  val_loader = gd.DistNeighborLoader(ds, fan, ('paper', val_idx), batch_size=val_bs, shuffle=False,
                                     collect_features=True, to_device=device, edge_dir=a.edge_dir,
                                     worker_options=opts(a.master_port + 2))
  first = next(iter(train_loader))
  in_dim = next(iter(first.x_dict.values())).shape[1]
  n_cls = int(ds.node_labels['paper'].max()) + 1
  model = RGNN(list(first.edge_index_dict.keys()), in_dim, a.hidden_channels, n_cls, num_layers=len(fan),
               node_type='paper', model=a.model, heads=a.num_heads).to(device)
  model = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
  opt = torch.optim.Adam(model.parameters(), lr=a.learning_rate)
  start_epoch = 0
  if a.ckpt_path:
    start_epoch = max(0, load_ckpt(0, a.ckpt_path, model.module, opt) + 1)
  for k, v in (('GLOBAL_BATCH_SIZE', train_bs * a.world), ('OPT_BASE_LR', a.learning_rate), ('SEED', a.random_seed)):
    log.event(k, v)
  log.end('INIT')
  log.start('RUN')
  step, acc = 0, 0.0
  steps_per_epoch = len(train_loader)
  # MLPerf validates several times per epoch (every 5 % of it) and stops at the first evaluation that reaches the
  # target; every rank takes the same branch because the accuracy is all-reduced
  eval_every = max(1, int(round(steps_per_epoch * a.validation_frac_within_epoch))) \
      if a.validation_frac_within_epoch < 1.0 else 0

  def validate(epoch_num):
    log.start('EVAL', epoch_num=epoch_num)
    v = evaluate(model, val_loader, device, a.with_trim)
    log.end('EVAL', epoch_num=epoch_num)
    log.event('EVAL_ACCURACY', v, {'epoch_num': epoch_num})
    return v

  reached = False
  for epoch in range(start_epoch, a.epochs):
    log.start('EPOCH', epoch_num=epoch)
    t0 = time.time()
    for i, b in enumerate(train_loader):
      if reached or 0 <= a.max_steps <= i:
        continue                      # keep draining so that every rank sees the same number of batches
      bs = b['paper'].batch_size
      with torch.autocast(device.type, dtype=torch.bfloat16, enabled=a.precision == 'bf16'):
        out = model(b.x_dict, b.edge_index_dict, *((b.num_sampled_nodes, b.num_sampled_edges) if a.with_trim else ()))
        loss = F.cross_entropy(out[:bs].float(), b['paper'].y[:bs].to(device))
      opt.zero_grad(); loss.backward(); opt.step()
      step += 1
      if a.ckpt_steps > 0 and step % a.ckpt_steps == 0 and a.rank == 0 and a.ckpt_path:
        save_ckpt(step, a.ckpt_path, model.module, opt, epoch)
      if eval_every and (i + 1) % eval_every == 0 and i + 1 < steps_per_epoch:
        acc = validate(epoch + (i + 1) / steps_per_epoch)
        if a.rank == 0:
          print(f'epoch {epoch} step {i + 1}/{steps_per_epoch}: val-acc {acc:.4f}', flush=True)
        reached = acc >= a.target_acc
    log.end('EPOCH', epoch_num=epoch)
    if a.evaluate_on_epoch_end and not reached:
      acc = validate(epoch + 1)
    if a.rank == 0:
      print(f'epoch {epoch}: loss {float(loss.detach()):.4f} val-acc {acc:.4f} ({time.time() - t0:.1f}s)', flush=True)
      if a.ckpt_path and a.checkpoint_on_epoch_end:
        save_ckpt(0, a.ckpt_path, model.module, opt, epoch)
    dist.barrier()
    if acc >= a.target_acc:
      break
  log.end('RUN', status='success' if acc >= a.target_acc else 'aborted')
  train_loader.shutdown(); val_loader.shutdown()
  dist.barrier()
  if gd.rpc_is_initialized():
    gd.shutdown_rpc()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
