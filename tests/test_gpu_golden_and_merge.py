"""CUDA path vs the committed golden fixture, and the shard-merge kernel vs its host specification."""
import os

import numpy as np
import pytest

from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction, merge_topk_device
from turdb_b200.sharding import merge_topk_host

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hnsw_small.npz")


def test_cuda_search_matches_golden(gpu_required):
    z = np.load(GOLD)
    arrays = {k: z[k] for k in ("vectors", "row_ids", "levels", "l0_adj", "l0_cnt", "up_base", "up_adj", "up_cnt")}
    arrays["entry"], arrays["max_level"] = int(z["entry"]), int(z["max_level"])
    idx = CudaHnswIndex.from_graph(arrays)
    for metric, name in ((0, "l2"), (1, "cosine"), (2, "ip")):
        rows, nodes, dist, cnt, st = idx.search_batch(z["queries"], 10, 40, DistanceFunction(metric))
        assert np.array_equal(nodes, z[f"{name}_nodes"])
        assert np.array_equal(dist.view(np.uint32), z[f"{name}_dist"].view(np.uint32))
        assert np.array_equal(rows, z[f"{name}_rows"]) and np.array_equal(cnt, z[f"{name}_counts"])
        gold_st = z[f"{name}_stats"]
        for f in ("n_dist", "n_dist_upper", "n_expanded", "n_upper_hops"):
            bad = np.where(st[f] != gold_st[f])[0]
            assert bad.size == 0, (name, f, bad[:5], st[f][bad[:5]], gold_st[f][bad[:5]])
    idx.close()


def test_merge_kernel_matches_host_spec(gpu_required):
    import torch
    rng = np.random.default_rng(3)
    for n_shards, nq, k in [(1, 7, 5), (2, 64, 10), (8, 33, 100), (4, 5, 1)]:
        dist = np.sort(rng.integers(0, 50, (n_shards, nq, k)).astype(np.float32) / 4, axis=2)  # many ties
        rows = rng.integers(0, 1 << 40, (n_shards, nq, k)).astype(np.uint64)
        cnt = rng.integers(0, k + 1, (n_shards, nq)).astype(np.uint32)
        dev = torch.device("cuda:0")
        d_rows = torch.from_numpy(rows.view(np.int64)).to(dev)
        d_dist = torch.from_numpy(dist).to(dev)
        d_cnt = torch.from_numpy(cnt.view(np.int32)).to(dev)
        o_rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
        o_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        o_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        merge_topk_device(0, d_rows.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), n_shards, nq, k,
                          o_rows.data_ptr(), o_dist.data_ptr(), o_cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        r, d, c = merge_topk_host(rows, dist, cnt, k)
        assert np.array_equal(o_cnt.cpu().numpy().view(np.uint32), c)
        assert np.array_equal(o_dist.cpu().numpy(), d)
        assert np.array_equal(o_rows.cpu().numpy().view(np.uint64), r)
        # the packed form: one block per shard (rows | distances | counts), as a single all-gather delivers them
        from turdb_b200.hnsw import merge_topk_packed_device
        from turdb_b200.sharding import pack_layout
        o_r, o_d, o_c, size = pack_layout(nq, k)
        blocks = np.zeros((n_shards, size), np.uint8)
        for s_ in range(n_shards):
            blocks[s_, o_r:o_d] = rows[s_].view(np.uint8).ravel()
            blocks[s_, o_d:o_c] = dist[s_].view(np.uint8).ravel()
            blocks[s_, o_c:o_c + nq * 4] = cnt[s_].view(np.uint8).ravel()
        d_blocks = torch.from_numpy(blocks).to(dev)
        o_rows.zero_(); o_dist.zero_(); o_cnt.zero_()
        merge_topk_packed_device(0, d_blocks.data_ptr(), size, n_shards, nq, k, o_rows.data_ptr(), o_dist.data_ptr(),
                                 o_cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(o_cnt.cpu().numpy().view(np.uint32), c)
        assert np.array_equal(o_dist.cpu().numpy(), d)
        assert np.array_equal(o_rows.cpu().numpy().view(np.uint64), r)


def test_single_process_sharded_search_matches_per_shard_oracle_merge(gpu_required):
    """turdb_cuda_shards_search_batch: 3 sub-indexes (here on one device), device-to-device gather + merge, against
    the per-shard oracle searches merged by the host specification of the merge (sharding.merge_topk_host)."""
    import numpy as np
    from oracle import binding as ob
    from turdb_b200 import datasets as ds
    from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction, shards_search_batch
    from turdb_b200.sharding import merge_topk_host, shard_bounds
    x = ds.gaussian_latent(6000, 48, seed=21)
    q = ds.gaussian_latent(100, 48, seed=22)
    shards, o_rows, o_dist, o_cnt = [], [], [], []
    for r in range(3):
        lo, hi = shard_bounds(6000, 3, r)
        g = ob.OracleGraph.build(x[lo:hi], seed=5 + r, row_ids=np.arange(lo, hi, dtype=np.uint64))
        shards.append(CudaHnswIndex.from_graph(g.export()))
        c = g.search(q, 10, 64, ob.COSINE, n_threads=4)
        o_rows.append(c[0]); o_dist.append(c[2]); o_cnt.append(c[3])
    rows, dist, counts = shards_search_batch(shards, q, 10, 64, DistanceFunction.Cosine)
    e_rows, e_dist, e_cnt = merge_topk_host(np.stack(o_rows), np.stack(o_dist), np.stack(o_cnt), 10)
    assert np.array_equal(counts, e_cnt) and np.array_equal(rows, e_rows)
    assert np.array_equal(dist.view(np.uint32), e_dist.view(np.uint32))
    for s in shards:
        s.close()
