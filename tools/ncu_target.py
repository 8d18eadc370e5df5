"""Short target for ncu: build one graph, run the traversal kernel a few times (TURDB_CUDA_LIB selects an A/B library)."""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from turdb_b200 import datasets as ds
from turdb_b200.graph_build import build_graph
from turdb_b200.hnsw import CudaHnswIndex

ap = argparse.ArgumentParser()
ap.add_argument("n", nargs="?", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--ef", type=int, default=128)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--metric", type=int, default=1)
ap.add_argument("--gen", default="gaussian_latent")
ap.add_argument("--genkw", default="")
ap.add_argument("--form", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--debug", action="store_true")
args = ap.parse_args()
kw = json.loads(args.genkw) if args.genkw else (dict(latent=16, normalise=args.metric == 1) if args.gen == "gaussian_latent" else {})
if args.gen == "clustered":
    kw.setdefault("corpus_n", args.n)
n, dim, nq, k, ef = args.n, args.dim, args.nq, args.k, args.ef
x = ds.make(args.gen, n, dim, seed=1, **kw)
q = ds.make(args.gen, nq, dim, seed=2, **kw)
arrays = build_graph(x, seed=42)
idx = CudaHnswIndex.from_graph(arrays)
idx.set_traversal_form(args.form)
dev = torch.device("cuda:0")
dq = torch.from_numpy(q).to(dev)
rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev)
stats = torch.empty((nq, 4), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(args.reps):
    idx.search_batch_device(dq.data_ptr(), nq, k, ef, args.metric, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), 0,
                            stats.data_ptr(), 0, stream)
torch.cuda.synchronize()
idx.profile_begin(3)
for _ in range(3):
    idx.search_batch_device(dq.data_ptr(), nq, k, ef, args.metric, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), 0,
                            stats.data_ptr(), 0, stream)
torch.cuda.synchronize()
km, om = idx.profile_read(3)
st = stats.cpu().numpy().astype(np.int64)
print("kernel ms", km.tolist(), "overflow ms", om.tolist())
print("n_dist", st[:, 0].mean(), "n_expanded", st[:, 2].mean(), "upper hops", st[:, 3].mean())
print("algorithmic bytes per launch", int((st[:, 0] * dim * 4 + st[:, 2] * 129 + st[:, 3] * 65 + dim * 4 + k * 12).sum()))
if args.debug:
    idx.debug_counters(True)
    idx.search_batch_device(dq.data_ptr(), nq, k, ef, args.metric, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), 0,
                            stats.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    c = idx.debug_counters(False).astype(np.float64)
    h = max(c[0], 1.0)
    print("dbg per-hop cycles: select %.0f adj+hash %.0f request %.0f merge %.0f | visited-insert %.0f | spec-hit %.2f | per-query total %.0f upper %.0f hops %.1f"
          % (c[1] / h, c[2] / h, c[3] / h, c[4] / h, c[12] / h, c[8] / h, c[9] / max(c[11], 1), c[10] / max(c[11], 1), c[0] / max(c[11], 1)))
