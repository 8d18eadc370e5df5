"""Dev probe: time the traversal kernel on an oracle-built graph for a few tunings (GPU box)."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--ef", type=int, default=64)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--metric", type=int, default=0)
ap.add_argument("--normalise", action="store_true")
ap.add_argument("--out", default="gpurun_out/probe.json")
args = ap.parse_args()

x = ds.gaussian_latent(args.n, args.dim, seed=1, normalise=args.normalise)
q = ds.gaussian_latent(args.nq, args.dim, seed=2, normalise=args.normalise)
t = time.time()
g = ob.OracleGraph.build(x, seed=42)
t_build = time.time() - t
print(f"oracle build {t_build:.1f}s", flush=True)
arrays = g.export()
idx = CudaHnswIndex.from_graph(arrays)
dev = torch.device("cuda:0")
dq = torch.from_numpy(q).to(dev)
rows = torch.empty((args.nq, args.k), dtype=torch.int64, device=dev)
dist = torch.empty((args.nq, args.k), dtype=torch.float32, device=dev)
nodes = torch.empty((args.nq, args.k), dtype=torch.int32, device=dev)
cnt = torch.empty(args.nq, dtype=torch.int32, device=dev)
stats = torch.empty((args.nq, 4), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream

t = time.time()
cpu = g.search(q[:2000], args.k, args.ef, args.metric, n_threads=8)
t_cpu = time.time() - t
cpu_qps = 2000 / t_cpu
print(f"cpu 8 threads: {cpu_qps:.0f} qps", flush=True)

res = []
for warps, slots, hb in [(0, 0, 0), (1, 8, 0), (1, 16, 0), (1, 32, 0), (2, 16, 0), (4, 16, 0), (4, 32, 0), (1, 32, 11), (1, 16, 11)]:
    try:
        idx.set_tuning(warps, slots, hb)
        def run():
            idx.search_batch_device(dq.data_ptr(), args.nq, args.k, args.ef, args.metric, rows.data_ptr(), dist.data_ptr(),
                                    cnt.data_ptr(), nodes.data_ptr(), stats.data_ptr(), 0, stream)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        st = stats.cpu().numpy().astype(np.int64)
        nbytes = (st[:, 0] * args.dim * 4 + st[:, 2] * 129 + st[:, 3] * 65 + args.dim * 4 + args.k * 12).sum()
        ok = bool(np.array_equal(nodes.cpu().numpy()[:2000].view(np.uint32), cpu[1]))
        r = dict(warps=warps, slots=slots, hash_bits=hb, ms=ms, qps=args.nq / ms * 1e3, gbs=nbytes / ms / 1e6,
                 n_dist=float(st[:, 0].mean()), n_exp=float(st[:, 2].mean()), parity=ok)
        print(json.dumps(r), flush=True)
        res.append(r)
    except Exception as ex:  # noqa
        print("tuning", warps, slots, hb, "failed:", ex, flush=True)
json.dump(dict(args=vars(args), build_s=t_build, cpu_qps_8thr=cpu_qps, runs=res), open(args.out, "w"), indent=1)
