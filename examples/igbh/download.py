"""Fetch the IGB-heterogeneous dataset into the layout `dataset.py` reads.

  python download.py --path /data/igbh --size tiny|small|medium      one tarball, unpacked in place
  python download.py --path /data/igbh --size full [--dry-run]       IGBH-full (~4 TB): file by file, resumable

The public bucket is the one the IGB authors publish (the reference's examples/igbh/download.py and
download_igbh_full.sh point at the same objects).  Transfers stream to disk in 8 MB pieces and resume from a partial
file with an HTTP Range request, so an interrupted multi-terabyte download continues where it stopped.  The build
environment of this repository has no network: `dataset.make_synthetic_igbh` writes the same layout for testing.
"""
import argparse
import os
import os.path as osp
import tarfile
import urllib.request

from dataset import BASE_ETYPES, NTYPES, VENUE_ETYPES, VENUE_NTYPES, etype_dir

BUCKET = 'https://igb-public.s3.us-east-2.amazonaws.com'
FULL_NTYPES = NTYPES + VENUE_NTYPES          # IGBH-full / large carry the two venue types
FULL_ETYPES = BASE_ETYPES + VENUE_ETYPES


def full_manifest():
  """Relative paths (under <path>/full/processed) of IGBH-full and their object keys."""
  files = []
  for nt in FULL_NTYPES:
    files.append(f'{nt}/node_feat.npy')
    files.append(f'{nt}/{nt}_id_index_mapping.npy')
  files += ['paper/node_label_19.npy', 'paper/node_label_2K.npy']
  files += [f'{etype_dir(et)}/edge_index.npy' for et in FULL_ETYPES]
  return [(f, f'{BUCKET}/IGBH/processed/{f}') for f in files]


def fetch(url: str, dst: str, chunk: int = 8 << 20):
  """Stream url -> dst, resuming from an existing partial file."""
  os.makedirs(osp.dirname(dst), exist_ok=True)
  part = dst + '.part'
  have = osp.getsize(part) if osp.exists(part) else 0
  req = urllib.request.Request(url, headers={'Range': f'bytes={have}-'} if have else {})
  with urllib.request.urlopen(req) as r:
    if have and r.status != 206:        # server ignored the range: start over
      have = 0
    total = have + int(r.headers.get('Content-Length', 0))
    with open(part, 'ab' if have else 'wb') as f:
      done = have
      while True:
        buf = r.read(chunk)
        if not buf:
          break
        f.write(buf)
        done += len(buf)
        print(f'\r{osp.basename(dst)}: {done / 2**20:.0f} / {total / 2**20:.0f} MiB', end='', flush=True)
  print()
  os.replace(part, dst)


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--path', required=True)
  p.add_argument('--size', default='tiny', choices=['tiny', 'small', 'medium', 'full'])
  p.add_argument('--dry-run', action='store_true', help='print what would be fetched')
  a = p.parse_args()
  if a.size == 'full':
    base = osp.join(a.path, 'full', 'processed')
    for rel, url in full_manifest():
      dst = osp.join(base, rel)
      if osp.exists(dst):
        continue
      print(('would fetch ' if a.dry_run else 'fetching ') + url)
      if not a.dry_run:
        fetch(url, dst)
    return
  url = f'{BUCKET}/igb-heterogeneous/igb_heterogeneous_{a.size}.tar.gz'
  tar = osp.join(a.path, osp.basename(url))
  print(('would fetch ' if a.dry_run else 'fetching ') + url)
  if a.dry_run:
    return
  if not osp.exists(tar):
    fetch(url, tar)
  with tarfile.open(tar) as t:
    t.extractall(a.path)
  print('unpacked into', a.path, '- next: split_seeds.py, compress_graph.py (see README.md)')


if __name__ == '__main__':
  main()
