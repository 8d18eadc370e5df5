"""Generates tests/golden/hnsw_small.npz with the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The fixture pins (a) the oracle against regressions and (b) the CUDA
path on the GPU box without rebuilding the graph there.  The reference itself holds no golden vectors for
HNSW search (parity unpinned); these come from the oracle port."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from turdb_b200 import datasets as ds  # noqa: E402

n, dim, nq, k, ef = 1500, 24, 48, 10, 40
# full-mantissa floats on purpose: quantised data produces exact distance ties, and ties are the one place
# where the reference's two heaps and the kernel's sorted list may legitimately differ
x = ds.gaussian_latent(n, dim, seed=101, latent=6)
q = ds.gaussian_latent(nq, dim, seed=102, latent=6)
g = ob.OracleGraph.build(x.astype(np.float32), m=16, ef_construction=100, mode=ob.BUILD_INTENT, seed=103,
                         row_ids=np.arange(n, dtype=np.uint64) * 2 + 1)
a = g.export()
out = {k_: v for k_, v in a.items() if isinstance(v, np.ndarray)}
out["entry"] = np.uint32(a["entry"])
out["max_level"] = np.uint8(a["max_level"])
out["queries"] = q.astype(np.float32)
vis = (np.arange(n) % 3 != 0)
out["visible_mask"] = vis
for metric, name in ((ob.L2, "l2"), (ob.COSINE, "cosine"), (ob.IP, "ip")):
    rows, nodes, dist, cnt, st = g.search(q, k, ef, metric)
    out[f"{name}_nodes"], out[f"{name}_dist"], out[f"{name}_counts"], out[f"{name}_stats"] = nodes, dist, cnt, st
    out[f"{name}_rows"] = rows
words = np.packbits(np.pad(vis, (0, (-n) % 64)).reshape(-1, 64), axis=1, bitorder="little").view(np.uint64).ravel()
rows, nodes, dist, cnt, st = g.search(q, k, ef, ob.L2, visible=words)
out["filtered_nodes"], out["filtered_dist"], out["filtered_counts"] = nodes, dist, cnt
erows, edist, ecnt = ob.sql_topk(x.astype(np.float32), q, k, op=ob.L2)
out["sql_l2_rows"], out["sql_l2_dist"] = erows, edist
erows, edist, ecnt = ob.sql_topk(x.astype(np.float32), q, k, op=ob.COSINE)
out["sql_cos_rows"], out["sql_cos_dist"] = erows, edist
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hnsw_small.npz"), **out)
print("wrote", os.path.getsize(os.path.join(ROOT, "tests", "golden", "hnsw_small.npz")), "bytes")
