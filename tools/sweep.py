"""Dev tool: build one graph on the GPU, then sweep traversal-kernel tunings (GPU box)."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from turdb_b200 import datasets as ds
from turdb_b200.graph_build import build_graph
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--ef", type=int, default=128)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--metric", type=int, default=1)
ap.add_argument("--latent", type=int, default=16)
ap.add_argument("--tunings", default="0,0,0;1,8,0;1,16,0;1,24,0;1,32,0;2,16,0;1,16,12;1,24,12")
ap.add_argument("--out", default="gpurun_out/sweep.json")
ap.add_argument("--debug", action="store_true")
args = ap.parse_args()

norm = args.metric == 1
x = ds.gaussian_latent(args.n, args.dim, seed=1, latent=args.latent, normalise=norm)
q = ds.gaussian_latent(args.nq, args.dim, seed=2, latent=args.latent, normalise=norm)
t = time.time()
arrays = build_graph(x, seed=42)
torch.cuda.synchronize()
print(f"graph build {time.time() - t:.1f}s", flush=True)
idx = CudaHnswIndex.from_graph(arrays)
dev = torch.device("cuda:0")
dq = torch.from_numpy(q).to(dev)
rows = torch.empty((args.nq, args.k), dtype=torch.int64, device=dev)
dist = torch.empty((args.nq, args.k), dtype=torch.float32, device=dev)
nodes = torch.empty((args.nq, args.k), dtype=torch.int32, device=dev)
cnt = torch.empty(args.nq, dtype=torch.int32, device=dev)
stats = torch.empty((args.nq, 4), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
xd = torch.from_numpy(x).to(dev)
gt = torch.topk(dq[:1000] @ xd.T, args.k, dim=1).indices.cpu().numpy() if norm else None
del xd
res = []
ref_nodes = None
for tun in args.tunings.split(";"):
    warps, slots, hb, segs = ([int(v) for v in tun.split(",")] + [0])[:4]
    try:
        idx.set_tuning(warps, slots, hb, segs)
        def run():
            idx.search_batch_device(dq.data_ptr(), args.nq, args.k, args.ef, args.metric, rows.data_ptr(), dist.data_ptr(),
                                    cnt.data_ptr(), nodes.data_ptr(), stats.data_ptr(), 0, stream)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        reps = 5
        idx.profile_begin(reps)
        for _ in range(reps):
            run()
        torch.cuda.synchronize()
        km, om = idx.profile_read(reps)
        if args.debug:
            idx.debug_counters(True)
            run(); torch.cuda.synchronize()
            c = idx.debug_counters(False).astype(np.float64)
            h = max(c[0], 1.0)
            print("  dbg per-hop cycles: select %.0f adj+hash %.0f request %.0f (w0: issue %.0f wait %.0f comp %.0f) merge %.0f | "
                  "spec-hit %.2f | per-query: total %.0f upper %.0f hops %.1f" % (c[1]/h, c[2]/h, c[3]/h, c[6]/h, c[5]/h, c[7]/h, c[4]/h,
                   c[8]/h, c[9]/max(c[11],1), c[10]/max(c[11],1), c[0]/max(c[11],1)), flush=True)
        ms = float(km.mean())
        st = stats.cpu().numpy().astype(np.int64)
        nbytes = (st[:, 0] * args.dim * 4 + st[:, 2] * 129 + st[:, 3] * 65 + args.dim * 4 + args.k * 12).sum()
        nd = nodes.cpu().numpy()
        if ref_nodes is None:
            ref_nodes = nd.copy()
        rec = None
        if gt is not None:
            rec = float(np.mean([len(set(nd[i].tolist()) & set(gt[i].tolist())) / args.k for i in range(1000)]))
        r = dict(warps=warps, slots=slots, hash_bits=hb, segs=segs, kernel_ms=ms, overflow_ms=float(om.mean()), qps=args.nq / ms * 1e3,
                 gbs=nbytes / ms / 1e6, n_dist=float(st[:, 0].mean()), n_exp=float(st[:, 2].mean()), recall=rec,
                 same_as_first=bool(np.array_equal(nd, ref_nodes)))
        print(json.dumps(r), flush=True)
        res.append(r)
    except Exception as ex:  # noqa
        print("tuning", tun, "failed:", ex, flush=True)
json.dump(dict(args=vars(args), runs=res), open(args.out, "w"), indent=1)
