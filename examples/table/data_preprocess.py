"""Write a products-shaped graph as *tables* (counterpart of the reference's
examples/pai/ogbn_products/data_preprocess.py, which uploads ODPS tables): a node table
(id, feature 'f0:f1:...', label) and an edge table (src_id, dst_id, weight) as parquet files.

  python examples/table/data_preprocess.py --out /tmp/products_tables [--nodes 20000]
"""
import argparse
import os
import sys

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import synthetic_homo  # noqa: E402

if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--out', required=True)
  ap.add_argument('--nodes', type=int, default=20000)
  ap.add_argument('--edges', type=int, default=200000)
  ap.add_argument('--dim', type=int, default=32)
  ap.add_argument('--classes', type=int, default=10)
  a = ap.parse_args()
  os.makedirs(a.out, exist_ok=True)
  ei, x, y = synthetic_homo(a.nodes, a.edges, a.dim, a.classes)
  feat = [':'.join(f'{v:.4f}' for v in row) for row in x.numpy()]            # the ODPS string convention
  pq.write_table(pa.table({'id': np.arange(a.nodes), 'feature': feat, 'label': y.numpy()}),
                 os.path.join(a.out, 'node.parquet'))
  pq.write_table(pa.table({'src_id': ei[0].numpy(), 'dst_id': ei[1].numpy(),
                           'weight': np.random.rand(ei.shape[1]).astype(np.float32)}),
                 os.path.join(a.out, 'edge.parquet'))
  print(f'wrote node.parquet ({a.nodes} rows) and edge.parquet ({ei.shape[1]} rows) to {a.out}')
