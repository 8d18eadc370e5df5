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
# both traversal forms on tight clusters (many empty hops), filtered too; the device insert path; SQL replay + fallbacks
xc = ds.clustered(4000, 128, seed=1, sigma=0.1, corpus_n=4000)
qc = ds.clustered(64, 128, seed=2, sigma=0.1, corpus_n=4000)
gc = ob.OracleGraph.build(xc, seed=7)
idx = CudaHnswIndex.from_graph(gc.export())
visc = visibility_bitmap(np.random.default_rng(4).random(4000) < 0.5)
for form in (1, 2):
    idx.set_traversal_form(form)
    for ef in (16, 64, 200):
        r = idx.search_batch(qc, 10, ef, DistanceFunction.L2)
        c = gc.search(qc, 10, ef, 0)
        assert np.array_equal(r[2].view(np.uint32), c[2].view(np.uint32)), (form, ef)
    idx.search_batch(qc, 10, 64, DistanceFunction.Cosine, visible=visc)
idx.set_tuning(hash_bits=8)  # force the global-bitset fallback pass
idx.search_batch(qc, 10, 64, DistanceFunction.L2)
idx.close()
rnd = ob.level_randoms(1500, 3)
og = ob.OracleGraph.new(100, 16, 100, ob.BUILD_INTENT)
og.insert_batch(np.arange(1500, dtype=np.uint64), x[:1500], rnd)
b1 = CudaHnswIndex.build(x[:1500], None, rnd, max_batch=1)
ge, oe = b1.export_graph(), og.export()
print("sequential build equals oracle:", all(np.array_equal(ge[k], oe[k]) for k in ("l0_adj", "l0_cnt", "up_adj", "up_cnt")), ge["entry"] == oe["entry"])
b1.close()
b2 = CudaHnswIndex.build(x, None, None, max_batch=256)
b2.search_batch(q, 10, 48, DistanceFunction.L2)
xd = x.copy()
xd[100:2600] = xd[99]
t = CudaHnswIndex.from_graph(dict(arrays, vectors=xd))
sb = VectorScanBatch(t, VectorOp.L2Distance, 5, 1, project=VectorOp.InnerProduct)
s = sb.execute(np.concatenate([xd[99:100], q[:7]]))
o = ob.sql_topk(xd, np.concatenate([xd[99:100], q[:7]]), 5, op=ob.L2, offset=1)
print("sql replay/fallback rows equal:", np.array_equal(s[0], o[0]), "keys equal:", np.array_equal(s[1].view(np.uint64), o[1].view(np.uint64)))
e2 = t.bruteforce_topk(np.concatenate([xd[99:100], q[:7]]), 20, DistanceFunction.L2)
t.close(); b2.close()
print("sanitize target ok")
