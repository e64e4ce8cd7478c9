"""Cluster launcher for the distributed examples / benchmarks (counterpart of the reference's
examples/distributed/run_dist_train_sage_sup.py and benchmarks/api/run_dist_bench.py: YAML config + ssh).

  python examples/distributed/launch.py --config examples/distributed/dist_train_sage_config.yml [--dry-run]

The YAML lists the nodes (host, optional ssh user / port / python / working dir, number of ranks on the node) and
the script to run with its arguments; `{rank}`, `{world}`, `{master_addr}`, `{master_port}`, `{local_rank}` are
substituted per process.  Ranks on `localhost` / `127.0.0.1` are started as local sub-processes, every other
host through ssh (paramiko).  The launcher streams the logs with a `[node:rank]` prefix, waits for every rank and
exits non-zero if any rank failed (the reference's tmux-based launcher does not collect exit codes).
"""
import argparse
import os
import shlex
import subprocess
import sys
import threading

import yaml

LOCAL = ('localhost', '127.0.0.1', '::1')


def expand(cfg):
  """-> list of (node dict, rank, local_rank, argv string)."""
  nodes = cfg['nodes']
  world = sum(int(n.get('ranks', 1)) for n in nodes)
  master_addr = cfg.get('master_addr', nodes[0]['host'] if nodes[0]['host'] not in LOCAL else '127.0.0.1')
  master_port = int(cfg.get('master_port', 29500))
  out, rank = [], 0
  for n in nodes:
    for lr in range(int(n.get('ranks', 1))):
      sub = dict(rank=rank, world=world, master_addr=master_addr, master_port=master_port, local_rank=lr)
      args = ' '.join(str(a).format(**sub) for a in cfg.get('args', []))
      py = n.get('python', cfg.get('python', sys.executable))
      cmd = f"{py} {cfg['script']} {args}"
      out.append((n, rank, lr, cmd))
      rank += 1
  return out


def stream(prefix, pipe, sink):
  for line in iter(pipe.readline, ''):
    sink.write(f'{prefix} {line}')
    sink.flush()


def run_local(node, rank, cmd, env_extra, results):
  env = dict(os.environ, **{k: str(v) for k, v in env_extra.items()})
  p = subprocess.Popen(shlex.split(cmd), cwd=node.get('workdir', os.getcwd()), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
  stream(f"[{node['host']}:{rank}]", p.stdout, sys.stdout)
  results[rank] = p.wait()


def run_ssh(node, rank, cmd, env_extra, results):
  import paramiko
  cli = paramiko.SSHClient()
  cli.set_missing_host_key_policy(paramiko.AutoAddPolicy())
  cli.connect(node['host'], port=int(node.get('port', 22)), username=node.get('user'),
              key_filename=node.get('key_filename'))
  exports = ' '.join(f'{k}={shlex.quote(str(v))}' for k, v in env_extra.items())
  full = f"cd {shlex.quote(node.get('workdir', '.'))} && {exports} {cmd}"
  _, out, _ = cli.exec_command(full, get_pty=True)
  stream(f"[{node['host']}:{rank}]", out, sys.stdout)
  results[rank] = out.channel.recv_exit_status()
  cli.close()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--config', required=True)
  ap.add_argument('--dry-run', action='store_true', help='print the per-rank command lines and exit')
  a = ap.parse_args()
  with open(a.config) as f:
    cfg = yaml.safe_load(f)
  plan = expand(cfg)
  if a.dry_run:
    for node, rank, lr, cmd in plan:
      print(f"[{node['host']}:{rank}] {cmd}")
    return 0
  results, threads = {}, []
  for node, rank, lr, cmd in plan:
    env_extra = dict(cfg.get('env', {}), **node.get('env', {}))
    fn = run_local if node['host'] in LOCAL else run_ssh
    t = threading.Thread(target=fn, args=(node, rank, cmd, env_extra, results), daemon=True)
    t.start()
    threads.append(t)
  for t in threads:
    t.join()
  bad = {r: c for r, c in results.items() if c != 0}
  print(f'launch: {len(results) - len(bad)}/{len(plan)} ranks finished cleanly' + (f', failed: {bad}' if bad else ''))
  return 1 if bad or len(results) != len(plan) else 0


if __name__ == '__main__':
  sys.exit(main())
