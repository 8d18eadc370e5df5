"""Dev tool: recall of the bulk-built graph on clustered corpora as a function of sigma / builder knobs (GPU box)."""
import argparse, json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from turdb_b200 import datasets as ds
from turdb_b200.graph_build import build_graph
from turdb_b200.hnsw import CudaHnswIndex

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--sigmas", default="0.1,0.2,0.3")
ap.add_argument("--knn", default="64")
ap.add_argument("--efs", default="64,128,256")
ap.add_argument("--ivf", type=int, default=1024)
ap.add_argument("--variants", default="10:131072", help="ivf_probe:exact_prefix;... builder variants")
ap.add_argument("--out", default="gpurun_out/cluster_probe.json")
a = ap.parse_args()
dev = torch.device("cuda:0"); stream = torch.cuda.current_stream().cuda_stream
nq, k = 10000, 10
res = []
for sg in [float(v) for v in a.sigmas.split(",")]:
    x = ds.clustered(a.n, a.dim, seed=1, centre_latent=16, corpus_n=a.n, sigma=sg)
    q = ds.clustered(nq, a.dim, seed=2, centre_latent=16, corpus_n=a.n, sigma=sg)
    for kk, var in [(int(v), w) for v in a.knn.split(",") for w in a.variants.split(";")]:
        pr, px = [int(z) for z in var.split(":")]
        t = time.time()
        arrays = build_graph(x, seed=42, knn_k=kk, ivf_cells=a.ivf, ivf_probe=pr, ivf_exact_prefix=px)
        torch.cuda.synchronize(); tb = time.time() - t
        idx = CudaHnswIndex.from_graph(arrays)
        dq = torch.from_numpy(q).to(dev)
        rows = torch.empty((nq, k), dtype=torch.int64, device=dev); dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        nodes = torch.empty((nq, k), dtype=torch.int32, device=dev); cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        stats = torch.empty((nq, 4), dtype=torch.int32, device=dev)
        e_nodes = torch.empty((nq, k), dtype=torch.int32, device=dev)
        idx.bruteforce_topk_device(dq.data_ptr(), nq, k, 0, 4, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), e_nodes.data_ptr(), stream)
        torch.cuda.synchronize()
        gt = e_nodes.cpu().numpy()
        for ef in [int(v) for v in a.efs.split(",")]:
            for _ in range(2):
                idx.search_batch_device(dq.data_ptr(), nq, k, ef, 0, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), nodes.data_ptr(), stats.data_ptr(), 0, stream)
            torch.cuda.synchronize()
            idx.profile_begin(3)
            for _ in range(3):
                idx.search_batch_device(dq.data_ptr(), nq, k, ef, 0, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), nodes.data_ptr(), stats.data_ptr(), 0, stream)
            torch.cuda.synchronize()
            km, _ = idx.profile_read(3)
            nd = nodes.cpu().numpy()
            rec = float(np.mean([len(set(nd[i].tolist()) & set(gt[i].tolist())) / k for i in range(2000)]))
            r = dict(sigma=sg, knn_k=kk, ivf_probe=pr, exact_prefix=px, ef=ef, recall=rec, ms=float(km.mean()), n_dist=float(stats[:, 0].float().mean()), build_s=round(tb, 1))
            print(json.dumps(r), flush=True); res.append(r)
        idx.close(); del arrays
json.dump(res, open(a.out, "w"), indent=1)
