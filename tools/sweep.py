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
ap.add_argument("--gen", default="gaussian_latent", choices=["gaussian_latent", "sift_like", "clustered"])
ap.add_argument("--libs", default="", help="A/B: comma-separated library files; the graph is built once and each "
                "library is measured in a child process (TURDB_CUDA_LIB)")
ap.add_argument("--graph", default="", help="(internal) npz with a prebuilt graph")
ap.add_argument("--sq8", action="store_true", help="traverse the SQ8 arena (enable_sq8 + search_batch_sq8_device)")
ap.add_argument("--genkw", default="{}", help="JSON keyword arguments of the generator (clustered: corpus_n defaults to --n)")
ap.add_argument("--builder", default="knn", choices=["knn", "insert"], help="graph source: exact-kNN stand-in or the device insert path")
ap.add_argument("--probe", default="", help="gather-ceiling probe settings: ctas_per_sm,slots,cta_smem_bytes;...")
args = ap.parse_args()

norm = args.metric == 1
if args.gen == "gaussian_latent":
    x = ds.gaussian_latent(args.n, args.dim, seed=1, latent=args.latent, normalise=norm)
    q = ds.gaussian_latent(args.nq, args.dim, seed=2, latent=args.latent, normalise=norm)
else:
    kw = json.loads(args.genkw)
    if args.gen == "clustered":
        kw.setdefault("corpus_n", args.n)
    x = ds.make(args.gen, args.n, args.dim, seed=1, **kw)
    q = ds.make(args.gen, args.nq, args.dim, seed=2, **kw)
t = time.time()
if args.graph:
    z = np.load(args.graph)
    arrays = {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}
    arrays["vectors"] = x
elif args.builder == "insert":
    b = CudaHnswIndex.build(x, None, None, max_batch=8192, seed=42)
    arrays = b.export_graph(with_vectors=False)
    arrays["vectors"] = x
    b.close()
else:
    arrays = build_graph(x, seed=42)
    torch.cuda.synchronize()
print(f"graph build/load {time.time() - t:.1f}s", flush=True)
if args.libs:
    import os, subprocess
    gpath = "/tmp/sweep_graph.npz"
    np.savez(gpath, **{k: v for k, v in arrays.items() if k != "vectors"})
    del arrays
    torch.cuda.empty_cache()
    allres = {}
    for lib in args.libs.split(","):
        print(f"=== {lib}", flush=True)
        out = args.out + "." + os.path.basename(lib) + ".json"
        cmd = [sys.executable, __file__, "--graph", gpath, "--out", out] + [a for a in sys.argv[1:] if not a.startswith("--libs")]
        # drop the value that followed --libs / --out in the parent's argv
        clean, skip = [], False
        for a in sys.argv[1:]:
            if skip:
                skip = False
                continue
            if a in ("--libs", "--out"):
                skip = True
                continue
            clean.append(a)
        cmd = [sys.executable, __file__, "--graph", gpath, "--out", out] + clean
        subprocess.run(cmd, env=dict(os.environ, TURDB_CUDA_LIB=os.path.abspath(lib)))
        try:
            allres[lib] = json.load(open(out))["runs"]
        except Exception as ex:  # noqa
            allres[lib] = str(ex)
    json.dump(dict(args=vars(args), libs=allres), open(args.out, "w"), indent=1)
    sys.exit(0)
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
probes = []
for pr in [p for p in args.probe.split(";") if p]:
    c, sl, sm = [int(v) for v in pr.split(",")]
    gbs, ms = idx.gather_probe(c, sl, sm, 64)
    probes.append(dict(ctas_per_sm=c, slots=sl, cta_smem_bytes=sm, gbs=gbs, ms=ms))
    print("gather probe", probes[-1], flush=True)
for tun in args.tunings.split(";"):
    warps, slots, hb, segs, form = ([int(v) for v in tun.split(",")] + [0, 0])[:5]
    try:
        idx.set_tuning(warps, slots, hb, segs)
        idx.set_traversal_form(form)
        if args.sq8:
            idx.enable_sq8()
        def run():
            if args.sq8:
                idx.search_batch_sq8_device(dq.data_ptr(), args.nq, args.k, args.ef, args.metric, rows.data_ptr(), dist.data_ptr(),
                                            cnt.data_ptr(), nodes.data_ptr(), stats.data_ptr(), stream)
                return
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
            print("  dbg per-hop cycles: visited-insert (non-speculative) %.0f | helper warp 1: issue %.0f wait %.0f comp(+round-2 issue) %.0f"
                  % (c[12]/h, c[13]/h, c[14]/h, c[15]/h))
            print("  dbg per-hop cycles: select %.0f adj+hash %.0f request %.0f (w0: issue %.0f wait %.0f comp %.0f) merge %.0f | "
                  "spec-hit %.2f | per-query: total %.0f upper %.0f hops %.1f" % (c[1]/h, c[2]/h, c[3]/h, c[6]/h, c[5]/h, c[7]/h, c[4]/h,
                   c[8]/h, c[9]/max(c[11],1), c[10]/max(c[11],1), c[0]/max(c[11],1)), flush=True)
        ms = float(km.mean())
        st = stats.cpu().numpy().astype(np.int64)
        row_b = (args.dim + 8) if args.sq8 else args.dim * 4
        nbytes = (st[:, 0] * row_b + st[:, 2] * 129 + st[:, 3] * 65 + args.dim * 4 + args.k * 12).sum()
        nd = nodes.cpu().numpy()
        if ref_nodes is None:
            ref_nodes = nd.copy()
        rec = None
        if gt is not None:
            rec = float(np.mean([len(set(nd[i].tolist()) & set(gt[i].tolist())) / args.k for i in range(1000)]))
        r = dict(warps=warps, slots=slots, hash_bits=hb, segs=segs, form=form, kernel_ms=ms, overflow_ms=float(om.mean()), qps=args.nq / ms * 1e3,
                 gbs=nbytes / ms / 1e6, n_dist=float(st[:, 0].mean()), n_exp=float(st[:, 2].mean()), recall=rec,
                 same_as_first=bool(np.array_equal(nd, ref_nodes)))
        print(json.dumps(r), flush=True)
        res.append(r)
    except Exception as ex:  # noqa
        print("tuning", tun, "failed:", ex, flush=True)
json.dump(dict(args=vars(args), runs=res, gather_probe=probes), open(args.out, "w"), indent=1)
