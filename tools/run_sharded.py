"""BASELINE.json configs 4 and 5 on N GPUs (one process per GPU, torchrun): one sub-index per rank built by the device
insert path, queries replicated, per-shard top-k exchanged with ONE packed all-gather and merged on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_sharded.py --config 5 [--rows-per-rank 12500000] --out gpurun_out/x.json

config 5: `rows-per-rank` x 128 L2 clustered per rank (weak scaling: the corpus grows with N), ef sweep, QPS at the first
          ef with recall@10 >= 0.95 (recall of the MERGED result against the merged exact top-k), SQL batch path.
config 4: 10M x 768 inner product split over N ranks (strong scaling), ef 256, k 100.
Reports, per ef: global QPS (device events, max over ranks), per-rank traversal kernel ms, all-gather + merge ms, HBM GB/s.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction, merge_topk_packed_device
from turdb_b200.sharding import ShardedSearch, pack_layout

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True, choices=[4, 5])
ap.add_argument("--rows-per-rank", type=int, default=0)
ap.add_argument("--total-rows", type=int, default=0)
ap.add_argument("--nq", type=int, default=0)
ap.add_argument("--sigma", type=float, default=0.3)
ap.add_argument("--centre-latent", type=int, default=16)
ap.add_argument("--build-batch", type=int, default=8192)
ap.add_argument("--efs", default="")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--parity", type=int, default=300, help="queries checked against the CPU oracle on rank 0 (0 = skip)")
ap.add_argument("--out", default="gpurun_out/sharded.json")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))

if args.config == 5:
    dim, metric, k, m = 128, 0, 10, 16
    n = args.rows_per_rank or 12_500_000
    total = n * world
    nq = args.nq or 10_000
    efs = [int(e) for e in (args.efs or "16,32,64,128").split(",")]
    gen_kw = dict(sigma=args.sigma, centre_latent=args.centre_latent, corpus_n=total)
    # the corpus is one clustered set of `total` rows; rank r holds rows [r n, (r+1) n): same centres, disjoint draws
    x = ds.clustered(n, dim, seed=1000 + rank, **gen_kw)
    q = ds.clustered(nq, dim, seed=2, **gen_kw)
    name = f"config5 {total}x128 L2 clustered(sigma={args.sigma}, centre_latent={args.centre_latent}), {world} sub-indexes of {n}"
    scaling = "weak"
else:
    dim, metric, k, m = 768, 2, 100, 32
    total = args.total_rows or 10_000_000
    n = total // world
    nq = args.nq or 2000
    efs = [int(e) for e in (args.efs or "256").split(",")]
    x = ds.gaussian_latent(n, dim, seed=1000 + rank, latent=16, normalise=True)
    q = ds.gaussian_latent(nq, dim, seed=2, latent=16, normalise=True)
    name = f"config4 {total}x768 inner product, M=32 (caps 32/16), k=100, split over {world} GPUs ({n} rows each)"
    scaling = "strong"

t0 = time.time()
row_ids = np.arange(n, dtype=np.uint64) + np.uint64(rank) * np.uint64(n)
idx = CudaHnswIndex.build(x, row_ids, None, m=m, ef_construction=100, mode=1, max_batch=args.build_batch, device=local_rank,
                          metric=DistanceFunction(metric), seed=7 + rank)
torch.cuda.synchronize()
t_build = time.time() - t0

dq = torch.from_numpy(q).to(dev)
nodes = torch.empty((nq, k), dtype=torch.int32, device=dev)
stats = torch.empty((nq, 4), dtype=torch.int32, device=dev)
m_rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
m_dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
m_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
cur_ef = [efs[0]]


def local_search(dq_batch, o_rows, o_dd, o_cnt):
    idx.search_batch_device(dq_batch.data_ptr(), nq, k, cur_ef[0], metric, o_rows.data_ptr(), o_dd.data_ptr(), o_cnt.data_ptr(),
                            nodes.data_ptr(), stats.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)


def merge(gathered, block_bytes):
    merge_topk_packed_device(local_rank, gathered.data_ptr(), block_bytes, world, nq, k, m_rows.data_ptr(), m_dd.data_ptr(),
                             m_cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return m_rows, m_dd, m_cnt


sharded = ShardedSearch(dist, world, local_search, merge, nq, k, dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# exact ground truth of the MERGED problem: per-shard exact top-k (certified tensor-core filter + FP32 rerank), merged
n_gt = min(nq, 1000)
gt_sh = ShardedSearch(dist, world, None, merge, nq, k, dev)
e_nodes = torch.empty((nq, k), dtype=torch.int32, device=dev)
have_gt = dim <= 2048
if have_gt:
    idx.bruteforce_topk_device(dq.data_ptr(), nq, k, metric, 4, gt_sh.rows.data_ptr(), gt_sh.dd.data_ptr(), gt_sh.cnt.data_ptr(),
                               e_nodes.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if world > 1:
        dist.all_gather_into_tensor(gt_sh.gathered, gt_sh.block)
        merge(gt_sh.gathered, gt_sh.block_bytes)
        gt_rows = m_rows.cpu().numpy().copy()
    else:
        gt_rows = gt_sh.rows.cpu().numpy().copy()
    torch.cuda.synchronize()

sweep = []
for ef in efs:
    cur_ef[0] = ef
    for _ in range(3):
        sharded.search_batch(dq)
    barrier()
    idx.profile_begin(args.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sharded.search_batch(dq)
    e1.record()
    barrier()
    step_ms = allmax(e0.elapsed_time(e1) / args.steps)
    km, om = idx.profile_read(args.steps)
    kern_ms = float(km.mean())
    st = stats.cpu().numpy().astype(np.int64)
    nbytes = int((st[:, 0] * dim * 4 + st[:, 2] * 129 + st[:, 3] * 65 + dim * 4 + k * 12).sum())
    res_rows = (m_rows if world > 1 else sharded.rows).cpu().numpy()
    rec = None
    if have_gt:
        kk = min(k, 10)
        rec = float(np.mean([len(set(res_rows[i, :kk].tolist()) & set(gt_rows[i, :kk].tolist())) / kk for i in range(n_gt)]))
    all_kern = [None] * world
    if world > 1:
        dist.all_gather_object(all_kern, kern_ms)
    else:
        all_kern = [kern_ms]
    sweep.append(dict(ef=ef, step_ms=step_ms, global_qps=nq / step_ms * 1e3, kernel_ms_per_rank=all_kern,
                      exchange_and_merge_ms=step_ms - max(all_kern), recall_at_10=rec,
                      rank0_algorithmic_gb_per_launch=nbytes / 1e9, rank0_achieved_gbs=nbytes / kern_ms / 1e6,
                      rank0_frac_of_measured_hbm_peak=nbytes / kern_ms / 1e6 / 6524.9, rank0_n_dist=float(st[:, 0].mean()),
                      rank0_n_expanded=float(st[:, 2].mean())))
    if rank == 0:
        print(json.dumps(sweep[-1]), flush=True)

sql = None
if args.config == 5:
    # the SQL `ORDER BY vec <-> q LIMIT 10` batch path, index-backed: one launch per batch on every rank; the per-rank rows
    # + f64 keys are merged by key on the host side of this tool (global statement = merge of per-shard statements)
    from turdb_b200.sql_operator import VectorOp, VectorScanBatch
    ok = [s_ for s_ in sweep if s_["recall_at_10"] is not None and s_["recall_at_10"] >= 0.95]
    ef_sql = ok[0]["ef"] if ok else efs[-1]
    sb = VectorScanBatch(idx, VectorOp.L2Distance, k, use_index=True, ef_search=ef_sql)
    s_rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
    s_keys = torch.empty((nq, k), dtype=torch.float64, device=dev)
    s_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        sb.execute_device(dq.data_ptr(), nq, s_rows.data_ptr(), s_keys.data_ptr(), s_cnt.data_ptr(), stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sb.execute_device(dq.data_ptr(), nq, s_rows.data_ptr(), s_keys.data_ptr(), s_cnt.data_ptr(), stream)
    e1.record()
    barrier()
    sms = allmax(e0.elapsed_time(e1) / 5)
    sql = dict(ef=ef_sql, per_rank_ms=sms, statements_per_s_per_rank=nq / sms * 1e3)

def host_mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


parity = None
if rank == 0 and args.parity and host_mem_available_gb() < 2.5 * x.nbytes / 1e9 + 8:
    parity = dict(skipped=f"host MemAvailable {host_mem_available_gb():.0f} GB: the oracle copies the {x.nbytes / 1e9:.0f} GB arena")
elif rank == 0 and args.parity:
    from oracle import binding as ob
    cur_ef[0] = efs[-1]
    local_search(dq, sharded.rows, sharded.dd, sharded.cnt)
    torch.cuda.synchronize()
    g_nodes = nodes.cpu().numpy().view(np.uint32)
    g_dist = sharded.dd.cpu().numpy()
    st = stats.cpu().numpy().astype(np.int64)
    arrays = idx.export_graph(with_vectors=False)
    arrays["vectors"] = x
    og = ob.OracleGraph.from_arrays(arrays)
    cores = os.cpu_count() or 1
    t = time.perf_counter()
    c = og.search(q[:args.parity], k, cur_ef[0], metric, n_threads=cores)
    t_cpu = time.perf_counter() - t
    same = np.array([np.array_equal(g_nodes[i], c[1][i]) for i in range(args.parity)])
    same_set = np.array([set(g_nodes[i].tolist()) == set(c[1][i].tolist()) for i in range(args.parity)])
    same_d = np.array([np.array_equal(g_dist[i].view(np.uint32), c[2][i].view(np.uint32)) for i in range(args.parity)])
    parity = dict(ef=cur_ef[0], queries=args.parity, id_list_match=float(same.mean()), id_set_match=float(same_set.mean()),
                  distance_bits_match=float(same_d.mean()), n_dist_match=float((st[:args.parity, 0] == c[4]["n_dist"]).mean()),
                  cpu_qps_all_threads=args.parity / t_cpu, cpu_threads=cores)

if rank == 0:
    ok = [s_ for s_ in sweep if s_["recall_at_10"] is not None and s_["recall_at_10"] >= 0.95]
    rec = dict(name=name, n_gpus=world, scaling=scaling, rows_per_gpu=n, total_rows=total, dim=dim, metric=["l2", "cosine", "ip"][metric],
               k=k, batch=nq, graph=f"device insert path (reference-intent, efC=100, steps <= {args.build_batch})", build_s=round(t_build, 1),
               exchange="one all-gather of the packed per-rank top-k (rows | distances | counts) + merge kernel" if world > 1 else "none",
               packed_block_bytes=pack_layout(nq, k)[3], sweep=sweep, qps_at_recall_0_95=ok[0] if ok else None, sql_batch_path=sql,
               parity_rank0=parity)
    print(json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rec, open(args.out, "w"), indent=1)
idx.close()
if world > 1:
    dist.destroy_process_group()
