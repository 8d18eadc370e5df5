"""Runs the BASELINE.json configs as parity + throughput cases on one GPU and writes one JSON record per config.
Configs 3-5 name corpora larger than one build fits in minutes; rows per GPU are stated in each record."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.graph_build import build_graph
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/configs.json")
ap.add_argument("--only", default="")
ap.add_argument("--heavy", action="store_true", help="also run the configs marked heavy (minutes of build time)")
args = ap.parse_args()
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
cores = os.cpu_count() or 1

CONFIGS = [
    dict(name="config1_hnsw_integration_scale", n=10_000, dim=128, metric=0, gen="gaussian_latent", kw=dict(latent=16),
         m=16, ef=64, k=10, nq=1000, builder="oracle-reference-intent"),
    dict(name="config2_1Mx384_cosine", n=1_000_000, dim=384, metric=1, gen="gaussian_latent", kw=dict(latent=16, normalise=True),
         m=16, ef=128, k=10, nq=10_000, builder="knn-heuristic"),
    dict(name="config3_1Mx128_sift_like_L2", n=1_000_000, dim=128, metric=0, gen="sift_like", kw={},
         m=16, ef=128, k=10, nq=10_000, builder="knn-heuristic"),
    dict(name="config4_768_inner_product_M32_ef256_k100 (1.25M rows per GPU = 10M over 8 GPUs)", n=1_250_000, dim=768, metric=2,
         gen="gaussian_latent", kw=dict(latent=16, normalise=True), m=32, ef=256, k=100, nq=2000, builder="knn-heuristic"),
    dict(name="config5_128_L2_clustered_latent_centres (2M rows per GPU of the 12.5M named)", n=2_000_000, dim=128, metric=0,
         gen="clustered", kw=dict(centre_latent=16, corpus_n=2_000_000), m=16, ef=64, k=10, nq=10_000, builder="knn-heuristic",
         sql=True, ef_sweep=[64, 128, 256, 512]),
    dict(name="config4_full_10Mx768_inner_product_M32_ef256_k100 (the whole corpus on ONE GPU: the N=1 point of the sharded config)",
         n=10_000_000, dim=768, metric=2, gen="gaussian_latent", kw=dict(latent=16, normalise=True), m=32, ef=256, k=100, nq=2000,
         builder="knn-heuristic", heavy=True, no_cpu=True, min_host_gb=120),
    dict(name="config5_full_shard_12.5Mx128_L2_clustered_latent_centres (one of the 8 sub-indexes of 100M)", n=12_500_000, dim=128,
         metric=0, gen="clustered", kw=dict(centre_latent=16, corpus_n=12_500_000), m=16, ef=64, k=10, nq=10_000,
         builder="knn-heuristic", sql=True, heavy=True, ef_sweep=[64, 128, 256, 512]),
    dict(name="config5_full_shard_12.5Mx128_L2_clustered_sigma0.3 (one of the 8 sub-indexes of 100M; clusters as wide as their spacing)",
         n=12_500_000, dim=128, metric=0, gen="clustered", kw=dict(centre_latent=16, corpus_n=12_500_000, sigma=0.3), m=16, ef=128,
         k=10, nq=10_000, builder="knn-heuristic", sql=True, heavy=True, ef_sweep=[64, 128, 256, 512]),
    dict(name="config5_128_L2_clustered_iid_centres (2M rows per GPU; i.i.d. centres, recall ceiling documented)", n=2_000_000,
         dim=128, metric=0, gen="clustered", kw=dict(corpus_n=2_000_000), m=16, ef=64, k=10, nq=10_000, builder="knn-heuristic"),
]

out = []
for cfg in CONFIGS:
    if args.only and args.only not in cfg["name"]:
        continue
    if cfg.get("heavy") and not (args.heavy or args.only):
        continue
    if cfg.get("min_host_gb"):
        avail = 0
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) / 1e6
        if avail < cfg["min_host_gb"]:
            print(json.dumps(dict(name=cfg["name"], skipped=f"host MemAvailable {avail:.0f} GB < {cfg['min_host_gb']} GB")), flush=True)
            continue
    t0 = time.time()
    x = ds.make(cfg["gen"], cfg["n"], cfg["dim"], seed=1, **cfg["kw"])
    q = ds.make(cfg["gen"], cfg["nq"], cfg["dim"], seed=2, **cfg["kw"])
    if cfg["builder"].startswith("oracle"):
        g = ob.OracleGraph.build(x, m=cfg["m"], seed=42)
        arrays = g.export()
    else:
        arrays = build_graph(x, m=cfg["m"], seed=42)
        g = None
    torch.cuda.synchronize()
    t_build = time.time() - t0
    idx = CudaHnswIndex.from_graph(arrays)
    nq, k, ef, metric, dim = cfg["nq"], cfg["k"], cfg["ef"], cfg["metric"], cfg["dim"]
    dq = torch.from_numpy(q).to(dev)
    rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    nodes = torch.empty((nq, k), dtype=torch.int32, device=dev)
    cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    stats = torch.empty((nq, 4), dtype=torch.int32, device=dev)

    def run():
        idx.search_batch_device(dq.data_ptr(), nq, k, ef, metric, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(),
                                nodes.data_ptr(), stats.data_ptr(), 0, stream)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    reps = 5
    idx.profile_begin(reps)
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    km, om = idx.profile_read(reps)
    st = stats.cpu().numpy().astype(np.int64)
    nbytes = int((st[:, 0] * dim * 4 + st[:, 2] * 129 + st[:, 3] * 65 + dim * 4 + k * 12).sum())
    g_nodes = nodes.cpu().numpy().view(np.uint32)
    g_dist = dist.cpu().numpy()
    rec = dict(name=cfg["name"], rows_per_gpu=cfg["n"], dim=dim, metric=["l2", "cosine", "ip"][metric], M=cfg["m"], ef=ef, k=k,
               batch=nq, generator=cfg["gen"], graph=arrays.get("provenance", cfg["builder"]), build_s=round(t_build, 1),
               kernel_ms=float(km.mean()), overflow_pass_ms=float(om.mean()), qps=nq / float(km.mean()) * 1e3,
               algorithmic_gb_per_launch=nbytes / 1e9, achieved_gbs=nbytes / float(km.mean()) / 1e6,
               frac_of_measured_hbm_peak=nbytes / float(km.mean()) / 1e6 / 6524.9,
               n_dist=float(st[:, 0].mean()), n_expanded=float(st[:, 2].mean()), n_upper_hops=float(st[:, 3].mean()),
               arena_mb=cfg["n"] * dim * 4 / 1e6)
    # exact path: ground truth for recall + its own throughput
    if dim <= 2048 and cfg["n"] <= 4_000_000:
        e_rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
        e_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        e_nodes = torch.empty((nq, k), dtype=torch.int32, device=dev)
        e_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        def exact():
            idx.bruteforce_topk_device(dq.data_ptr(), nq, k, metric, 4, e_rows.data_ptr(), e_dist.data_ptr(), e_cnt.data_ptr(),
                                       e_nodes.data_ptr(), stream)
        exact()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            exact()
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1) / 3
        gt = e_nodes.cpu().numpy().view(np.uint32)
        rec["exact_ms"] = ems
        rec["exact_tflops"] = 2.0 * nq * cfg["n"] * dim / ems / 1e9
        rec["exact_qps"] = nq / ems * 1e3
    else:  # the largest corpora: FP32 matmul ground truth on a subset (keeps the BF16 copy of the arena out of HBM)
        xd = torch.from_numpy(x).to(dev)
        sub = min(nq, 500)
        sc = dq[:sub] @ xd.T
        gt = torch.topk(sc, k, dim=1).indices.cpu().numpy().astype(np.uint32)
        del xd
    ng = gt.shape[0]
    rec["recall_at_k"] = float(np.mean([len(set(g_nodes[i].tolist()) & set(gt[i].tolist())) / k for i in range(ng)]))
    rec["recall_at_10"] = float(np.mean([len(set(g_nodes[i, :10].tolist()) & set(gt[i, :10].tolist())) / 10 for i in range(ng)]))
    if cfg.get("sql"):
        # the SQL `ORDER BY vec <-> q LIMIT k` batch path (index-backed TopK operator, f64 keys): one launch per batch
        from turdb_b200.sql_operator import VectorOp, VectorScanBatch
        sb = VectorScanBatch(idx, VectorOp.L2Distance, k, use_index=True, ef_search=ef)
        s_rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
        s_keys = torch.empty((nq, k), dtype=torch.float64, device=dev)
        s_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        for _ in range(2):
            sb.execute_device(dq.data_ptr(), nq, s_rows.data_ptr(), s_keys.data_ptr(), s_cnt.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            sb.execute_device(dq.data_ptr(), nq, s_rows.data_ptr(), s_keys.data_ptr(), s_cnt.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        sms = e0.elapsed_time(e1) / 5
        sr = s_rows.cpu().numpy().astype(np.uint64)
        rid = np.asarray(arrays["row_ids"], np.uint64)
        rec["sql_batch_ms"] = sms
        rec["sql_batch_statements_per_s"] = nq / sms * 1e3
        rec["sql_batch_recall_at_10"] = float(np.mean([len(set(sr[i].tolist()) & set(rid[gt[i, :10]].tolist())) / 10 for i in range(ng)]))
    if cfg.get("ef_sweep"):
        # QPS at the first ef_search whose recall@10 reaches 0.95 (BASELINE.json metric); same graph, same queries
        sweep = []
        for ef2 in cfg["ef_sweep"]:
            def run2():
                idx.search_batch_device(dq.data_ptr(), nq, k, ef2, metric, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(),
                                        nodes.data_ptr(), stats.data_ptr(), 0, stream)
            run2()
            torch.cuda.synchronize()
            idx.profile_begin(3)
            for _ in range(3):
                run2()
            torch.cuda.synchronize()
            km2, _ = idx.profile_read(3)
            st2 = stats.cpu().numpy().astype(np.int64)
            nb2 = int((st2[:, 0] * dim * 4 + st2[:, 2] * 129 + st2[:, 3] * 65 + dim * 4 + k * 12).sum())
            n2 = nodes.cpu().numpy().view(np.uint32)
            r10 = float(np.mean([len(set(n2[i, :10].tolist()) & set(gt[i, :10].tolist())) / 10 for i in range(ng)]))
            sweep.append(dict(ef=ef2, kernel_ms=float(km2.mean()), qps=nq / float(km2.mean()) * 1e3, recall_at_10=r10,
                              achieved_gbs=nb2 / float(km2.mean()) / 1e6, n_dist=float(st2[:, 0].mean())))
        rec["ef_sweep"] = sweep
        ok = [s_ for s_ in sweep if s_["recall_at_10"] >= 0.95]
        rec["qps_at_recall_0.95"] = ok[0] if ok else None
        run()  # restore the configured ef's outputs for the parity block below
        torch.cuda.synchronize()
        st = stats.cpu().numpy().astype(np.int64)
        g_nodes = nodes.cpu().numpy().view(np.uint32)
        g_dist = dist.cpu().numpy()
    if cfg.get("no_cpu"):  # the oracle would copy the whole arena on the host: skipped for the largest corpus
        rec["parity"] = None
        print(json.dumps(rec), flush=True)
        out.append(rec)
        idx.close()
        del arrays, x
        continue
    # parity vs the CPU oracle on a sample of the same graph
    if g is None:
        g = ob.OracleGraph.from_arrays(arrays)
    sample = min(nq, 1000)
    t = time.perf_counter()
    c_rows, c_nodes, c_dist, c_cnt, c_st = g.search(q[:sample], k, ef, metric, n_threads=cores)
    t_cpu = time.perf_counter() - t
    same_ids = np.array([np.array_equal(g_nodes[i], c_nodes[i]) for i in range(sample)])
    same_bits = np.array([np.array_equal(g_dist[i].view(np.uint32), c_dist[i].view(np.uint32)) for i in range(sample)])
    rec["parity"] = dict(queries=sample, id_match=float(same_ids.mean()), distance_bits_match=float(same_bits.mean()),
                         n_dist_match=float((st[:sample, 0] == c_st["n_dist"]).mean()))
    rec["cpu_qps_all_threads"] = sample / t_cpu
    rec["cpu_threads"] = cores
    print(json.dumps(rec), flush=True)
    out.append(rec)
    idx.close()
    del g, arrays, x
json.dump(out, open(args.out, "w"), indent=1)
