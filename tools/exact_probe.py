"""Dev probe: exact path (tcgen05 GEMM-filter + rerank) throughput and recall on BASELINE config 3 / 2 shapes."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--metric", type=int, default=0)
ap.add_argument("--gen", default="sift_like")
ap.add_argument("--rerank", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--debug", action="store_true")
ap.add_argument("--out", default="gpurun_out/exact_probe.json")
args = ap.parse_args()

if args.gen == "sift_like":
    x = ds.sift_like(args.n, args.dim, seed=1)
    q = ds.sift_like(args.nq, args.dim, seed=2)
else:
    x = ds.gaussian_latent(args.n, args.dim, seed=1, normalise=args.metric == 1)
    q = ds.gaussian_latent(args.nq, args.dim, seed=2, normalise=args.metric == 1)
n = args.n
g = dict(vectors=x, row_ids=np.arange(n, dtype=np.uint64), levels=np.zeros(n, np.uint8),
         l0_adj=np.full((n, 32), 0xFFFFFFFF, np.uint32), l0_cnt=np.zeros(n, np.uint8),
         up_base=np.full(n, 0xFFFFFFFF, np.uint32), up_adj=np.zeros((0, 16), np.uint32), up_cnt=np.zeros(0, np.uint8),
         entry=0, max_level=0)
idx = CudaHnswIndex.from_graph(g)
dev = torch.device("cuda:0")
dq = torch.from_numpy(q).to(dev)
rows = torch.empty((args.nq, args.k), dtype=torch.int64, device=dev)
dist = torch.empty((args.nq, args.k), dtype=torch.float32, device=dev)
nodes = torch.empty((args.nq, args.k), dtype=torch.int32, device=dev)
cnt = torch.empty(args.nq, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream

def run():
    idx.bruteforce_topk_device(dq.data_ptr(), args.nq, args.k, args.metric, args.rerank, rows.data_ptr(), dist.data_ptr(),
                               cnt.data_ptr(), nodes.data_ptr(), stream)
run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
if args.debug:
    idx.debug_counters(True)
    run(); torch.cuda.synchronize()
    c = idx.debug_counters(False).astype(np.float64)
    tiles = max(c[5], 1)
    print("dbg cycles per tile: producer wait a_empty %.0f  b_empty %.0f | mma wait a_full %.0f  t_empty %.0f  b_full %.0f | "
          "epilogue wait t_full %.0f  work %.0f | tiles/CTA %.0f  kernel cycles/CTA(sum over passes) %.0f" %
          (c[0] / tiles, c[1] / tiles, c[2] / tiles, c[3] / tiles, c[4] / tiles, c[6] / tiles, c[7] / tiles, tiles / 148, c[8] / 148), flush=True)
flops = 2.0 * args.nq * n * args.dim
# ground truth: FP32 (no TF32) matmul on a subset
torch.backends.cuda.matmul.allow_tf32 = False
xd = torch.from_numpy(x).to(dev)
sub = 500
qs = dq[:sub]
sc = qs @ xd.T
if args.metric == 0:
    key = (xd * xd).sum(1)[None, :] - 2 * sc
elif args.metric == 1:
    key = -sc / (xd.norm(dim=1)[None, :] * qs.norm(dim=1)[:, None])
else:
    key = -sc
gt = torch.topk(key, args.k, dim=1, largest=False).indices.cpu().numpy()
nd = nodes.cpu().numpy()
rec = float(np.mean([len(set(nd[i].tolist()) & set(gt[i].tolist())) / args.k for i in range(sub)]))
bad = int((cnt.cpu().numpy() != args.k).sum())
r = dict(args=vars(args), ms=ms, qps=args.nq / ms * 1e3, tflops=flops / ms / 1e9, recall_vs_fp32=rec, bad_counts=bad)
print(json.dumps(r), flush=True)
json.dump(r, open(args.out, "w"), indent=1)
