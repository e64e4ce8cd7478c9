"""Smoke-run the CPU-capable example scripts (tiny settings) so they cannot rot."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
  env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', ''))
  out = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
  return out.stdout


def test_llm_on_subgraphs_stub_backend():
  out = _run(['examples/gpt/arxiv_llm.py', '--backend', 'stub', '--batches', '8'])
  assert 'link-prediction accuracy' in out


def test_igbh_pipeline_single_and_distributed(tmp_path):
  """examples/igbh: synthetic IGBH on disk -> split seeds -> compress (CSC, fp16) -> single-process training,
  then two-stage partitioning (topology, per-partition features) -> 2-process distributed R-GNN training."""
  import threading
  sys.path.insert(0, os.path.join(ROOT, 'examples', 'igbh'))
  sys.path.insert(0, os.path.join(ROOT, 'examples'))
  from dataset import make_synthetic_igbh
  d, parts = str(tmp_path / 'igbh'), str(tmp_path / 'parts')
  make_synthetic_igbh(d, papers=1500)
  _run(['examples/igbh/split_seeds.py', '--path', d, '--validation_frac', '0.1'])
  _run(['examples/igbh/compress_graph.py', '--path', d, '--layout', 'CSC', '--use_fp16'])
  out = _run(['examples/igbh/train_rgnn.py', '--path', d, '--layout', 'CSC', '--use_fp16', '--fan_out', '4,4',
              '--epochs', '1', '--max_steps', '5', '--batch_size', '128'])
  assert 'val-acc' in out and 'EVAL_ACCURACY' in out
  _run(['examples/igbh/partition.py', '--src_path', d, '--dst_path', parts, '--num_partitions', '2',
        '--with_feature', '0', '--edge_assign_strategy', 'by_dst'])
  for r in (0, 1):
    _run(['examples/igbh/build_partition_feature.py', '--src_path', d, '--dst_path', parts, '--partition_idx', str(r)])
  from graphlearn_for_pytorch_b200.utils.common import get_free_port_block as get_free_port
  port = get_free_port()
  outs = {}

  def rank(r):
    outs[r] = _run(['examples/igbh/dist_train_rgnn.py', '--path', parts, '--rank', str(r), '--world', '2',
                    '--fan_out', '4,4', '--epochs', '1', '--batch_size', '128', '--master_port', str(port),
                    # MLPerf-run options of the reference script: in-epoch validation, bf16 autocast, seed
                    '--validation_frac_within_epoch', '0.5', '--precision', 'bf16', '--random_seed', '3',
                    '--num_heads', '2', '--validation_acc', '0.99'], timeout=500)
  th = [threading.Thread(target=rank, args=(r,)) for r in (0, 1)]
  [t.start() for t in th]
  [t.join() for t in th]
  assert 'val-acc' in outs.get(0, '') and 'RUN_STOP' in outs.get(0, '')
  assert 'epoch 0 step' in outs[0]                      # the in-epoch validation ran


def test_table_examples_single_and_distributed(tmp_path):
  """examples/table (the reference's PAI/ODPS examples): tables on disk -> TableDataset training, and
  table slices -> DistTableDataset online partitioning -> 2-process training."""
  import threading
  tables = str(tmp_path / 'tables')
  _run(['examples/table/data_preprocess.py', '--out', tables, '--nodes', '3000', '--edges', '30000'])
  out = _run(['examples/table/train_products_sage.py', '--tables', tables, '--epochs', '1', '--max_steps', '5'])
  assert 'train-acc' in out
  from graphlearn_for_pytorch_b200.utils.common import get_free_port_block as get_free_port
  port = get_free_port()
  outs = {}

  def rank(r):
    outs[r] = _run(['examples/table/dist_train_products_sage.py', '--tables', tables, '--rank', str(r), '--world', '2',
                    '--max_steps', '3', '--master_port', str(port), '--out', str(tmp_path / 'parts')], timeout=500)
  th = [threading.Thread(target=rank, args=(r,)) for r in (0, 1)]
  [t.start() for t in th]
  [t.join() for t in th]
  assert 'epoch 0 loss' in outs.get(0, '') and 'epoch 0 loss' in outs.get(1, '')


def test_hierarchical_hetero_sage_example():
  out = _run(['examples/hetero/hierarchical_sage.py', '--papers', '2000', '--epochs', '1', '--max_steps', '4'])
  assert 'trimmed' in out and 'train-acc' in out


def _run_ranks(cmds, timeout=600):
  """Run several example processes concurrently (one per rank); return their stdouts."""
  import threading
  outs = {}

  def go(i, args):
    outs[i] = _run(args, timeout=timeout)
  th = [threading.Thread(target=go, args=(i, c)) for i, c in enumerate(cmds)]
  [t.start() for t in th]
  [t.join() for t in th]
  assert len(outs) == len(cmds), 'a rank failed (see the assertion of its thread above)'
  return outs


def test_distributed_examples_and_benchmark(tmp_path):
  """examples/distributed/* and benchmarks/bench_dist_neighbor_loader.py on a freshly partitioned graph:
  worker mode with sampling SUB-PROCESSES (needs the __main__ guards), server-client mode, ZeRO + Join."""
  from graphlearn_for_pytorch_b200.utils.common import get_free_port_block as get_free_port
  parts = str(tmp_path / 'parts')
  _run(['examples/distributed/partition_dataset.py', '--out', parts, '--parts', '2', '--nodes', '4000', '--edges', '40000'])
  port = get_free_port()
  outs = _run_ranks([['examples/distributed/dist_train_sage.py', '--root', parts, '--rank', str(r), '--world', '2',
                      '--epochs', '1', '--workers', '1', '--master-port', str(port)] for r in (0, 1)])
  assert all('epoch 0 loss' in o for o in outs.values())
  assert 'epoch 0 test acc' in outs[0]                  # second (test) loader, hits all-reduced over the ranks
  port = get_free_port()
  outs = _run_ranks([['benchmarks/bench_dist_neighbor_loader.py', '--root', parts, '--rank', str(r), '--world', '2',
                      '--epochs', '1', '--workers', '1', '--master-port', str(port)] for r in (0, 1)])
  assert all('edges/s' in o for o in outs.values())
  port = get_free_port()
  srv = [['examples/distributed/server_client_mode/sage_server.py', '--root', parts, '--rank', str(r), '--servers', '2',
          '--clients', '1', '--master-port', str(port)] for r in (0, 1)]
  cli = [['examples/distributed/server_client_mode/sage_client.py', '--rank', '0', '--servers', '2', '--clients', '1',
          '--master-port', str(port)]]
  outs = _run_ranks(srv + cli)
  assert 'epoch 1 loss' in outs[2]
  port = get_free_port()
  outs = _run_ranks([['examples/distributed/dist_sage_unsup_zero.py', '--root', parts, '--rank', str(r), '--world', '2',
                      '--master-port', str(port)] for r in (0, 1)])
  assert all('epoch 1 loss' in o for o in outs.values())


@pytest.mark.parametrize('args', [
    ['examples/train_sage_products.py', '--nodes', '4000', '--edges', '40000', '--epochs', '1', '--batch', '256'],
    ['examples/train_sage_products.py', '--mode', 'engine', '--nodes', '4000', '--edges', '40000', '--epochs', '1',
     '--batch', '200'],
    ['examples/graph_sage_unsup.py'],
    ['examples/hetero/train_rgnn_igbh.py', '--papers', '2000', '--fanout', '4,4', '--epochs', '1', '--batch', '256'],
    ['examples/hetero/train_rgnn_igbh.py', '--papers', '2000', '--fanout', '4,4', '--epochs', '1', '--batch', '256',
     '--model', 'hgt'],
    ['examples/hetero/bipartite_sage_unsup.py'],
    ['examples/hetero/train_hgt_mag.py', '--papers', '2500', '--epochs', '1', '--batch', '256'],
    ['examples/hetero/train_hgt_mag_mp.py', '--papers', '2500', '--epochs', '1', '--batch', '256'],
    ['examples/feature_mp.py'],
    ['examples/seal_link_pred.py', '--links', '400', '--epochs', '2'],
], ids=lambda a: os.path.basename(a[0]) + ('-hgt' if 'hgt' in a else '') + ('-engine' if 'engine' in a else ''))
def test_single_process_examples(args):
  out = _run(args)
  assert 'loss' in out.lower() or 'first column' in out


def test_cluster_launcher_local_plan(tmp_path):
  """examples/distributed/launch.py: YAML plan -> one process per rank (local here, ssh for remote hosts)."""
  import yaml
  parts = str(tmp_path / 'parts')
  _run(['examples/distributed/partition_dataset.py', '--out', parts, '--parts', '2', '--nodes', '3000', '--edges', '30000'])
  from graphlearn_for_pytorch_b200.utils.common import get_free_port_block as get_free_port
  cfg = yaml.safe_load(open(os.path.join(ROOT, 'examples', 'distributed', 'dist_train_sage_config.yml')))
  cfg['master_port'] = get_free_port()
  cfg['args'] = [parts if a == '/tmp/glt_parts' else a for a in cfg['args']]
  path = str(tmp_path / 'plan.yml')
  yaml.safe_dump(cfg, open(path, 'w'))
  dry = _run(['examples/distributed/launch.py', '--config', path, '--dry-run'])
  assert dry.count('dist_train_sage.py') == 2 and '--rank 1 --world 2' in dry
  out = _run(['examples/distributed/launch.py', '--config', path])
  assert '2/2 ranks finished cleanly' in out


def test_igbh_multi_gpu_trainer_runs_on_two_cpu_ranks(tmp_path):
  """examples/igbh/train_rgnn_multi_gpu.py (loader back-end): dataset shared over IPC with two spawned trainers,
  DDP over gloo, in-epoch validation, step checkpoints and resume from a checkpoint."""
  sys.path.insert(0, os.path.join(ROOT, 'examples', 'igbh'))
  from dataset import make_synthetic_igbh
  d = str(tmp_path / 'igbh')
  make_synthetic_igbh(d, papers=1500)
  _run(['examples/igbh/split_seeds.py', '--path', d, '--validation_frac', '0.1'])
  ck = str(tmp_path / 'ck')
  out = _run(['examples/igbh/train_rgnn_multi_gpu.py', '--path', d, '--model', 'rsage', '--fan_out', '4,4',
              '--train_batch_size', '128', '--val_batch_size', '128', '--hidden_channels', '32', '--epochs', '1',
              '--max_steps', '4', '--val_batches', '2', '--ckpt_steps', '2', '--ckpt_dir', ck, '--world_size', '2'],
             timeout=600)
  assert 'val-acc' in out and os.path.exists(os.path.join(ck, 'model_step_2.ckpt'))
  out = _run(['examples/igbh/train_rgnn_multi_gpu.py', '--path', d, '--model', 'rsage', '--fan_out', '4,4',
              '--train_batch_size', '128', '--val_batch_size', '128', '--hidden_channels', '32', '--epochs', '1',
              '--max_steps', '2', '--val_batches', '2', '--ckpt_path', os.path.join(ck, 'model_step_2.ckpt'),
              '--world_size', '2'], timeout=600)
  assert 'val-acc' in out


def test_igbh_large_schema_with_venue_types(tmp_path):
  """IGBH-large / -full carry journal and conference nodes: the loader picks them up from the directory layout and
  the single-process trainer runs on the 6-type / 11-relation graph; download.py lists the full manifest."""
  sys.path.insert(0, os.path.join(ROOT, 'examples', 'igbh'))
  from dataset import IGBHeteroDataset, make_synthetic_igbh
  d = str(tmp_path / 'igbh')
  make_synthetic_igbh(d, papers=1200, with_venues=True)
  ds = IGBHeteroDataset(d, 'tiny')
  assert set(ds.ntypes) == {'paper', 'author', 'institute', 'fos', 'conference', 'journal'}
  assert ('journal', 'rev_published', 'paper') in ds.edge_dict and ('paper', 'venue', 'conference') in ds.edge_dict
  assert len(ds.etypes) == 11
  _run(['examples/igbh/split_seeds.py', '--path', d, '--validation_frac', '0.1'])
  out = _run(['examples/igbh/train_rgnn.py', '--path', d, '--fan_out', '3,3', '--epochs', '1', '--max_steps', '3',
              '--batch_size', '64'])
  assert 'val-acc' in out
  out = _run(['examples/igbh/download.py', '--path', d, '--size', 'full', '--dry-run'])
  assert out.count('would fetch') >= 16 and 'paper__venue__conference/edge_index.npy' in out


def test_cpu_mode_distributed_loader_benchmark_runs():
  """benchmarks/bench_dist_loader_cpu.py (2 trainers x 1 sampling worker over localhost RPC) on a small graph."""
  import json
  out = _run(['benchmarks/bench_dist_loader_cpu.py', 'ours', '--nodes', '20000', '--edges', '200000', '--workers', '1',
              '--epochs', '2', '--train-frac', '0.2'], timeout=400)
  line = json.loads([ln for ln in out.splitlines() if ln.startswith('{')][-1])
  # 4000 training seeds over 2 randomly drawn partitions, batch 1024: 2 batches per trainer, 3 for a trainer whose
  # partition happens to own more than 2048 of them
  assert line['impl'] == 'ours' and line['best_epoch']['batches'] in (4, 5) and line['best_epoch']['M_edges_per_s'] > 0
