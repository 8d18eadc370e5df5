"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): traversal (3 metrics, filtered, SQ8),
exact path, SQL operator, shard merge.  Results are also compared with the oracle."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction, visibility_bitmap
from turdb_b200.sql_operator import VectorOp, VectorScanBatch

x = ds.gaussian_latent(3000, 100, seed=1)  # ragged dim: tail path + padded rows
q = ds.gaussian_latent(64, 100, seed=2)
g = ob.OracleGraph.build(x, seed=42)
arrays = g.export()
idx = CudaHnswIndex.from_graph(arrays)
for metric in (0, 1, 2):
    r = idx.search_batch(q, 10, 48, DistanceFunction(metric))
    c = g.search(q, 10, 48, metric)
    assert np.array_equal(r[1], c[1]) and np.array_equal(r[2].view(np.uint32), c[2].view(np.uint32)), metric
vis = np.random.default_rng(3).random(3000) < 0.3
r = idx.search_batch(q, 10, 48, DistanceFunction.L2, visible=visibility_bitmap(vis))
e = idx.bruteforce_topk(q, 10, DistanceFunction.Cosine)
s = VectorScanBatch(idx, VectorOp.L2Distance, 7, 2).execute(q)
o = ob.sql_topk(x, q, 7, op=ob.L2, offset=2)
assert np.array_equal(s[0], o[0]) and np.array_equal(s[1].view(np.uint64), o[1].view(np.uint64))
import torch
idx.enable_sq8()
dev = torch.device("cuda:0")
dq = torch.from_numpy(q).to(dev)
rows = torch.empty((64, 10), dtype=torch.int64, device=dev); dd = torch.empty((64, 10), dtype=torch.float32, device=dev)
nodes = torch.empty((64, 10), dtype=torch.int32, device=dev); cnt = torch.empty(64, dtype=torch.int32, device=dev)
idx.search_batch_sq8_device(dq.data_ptr(), 64, 10, 48, 1, rows.data_ptr(), dd.data_ptr(), cnt.data_ptr(), nodes.data_ptr(), 0,
                            torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
assert int(cnt.min()) == 10
idx.close()
print("sanitize target ok")
