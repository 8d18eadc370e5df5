"""One sub-index per GPU (BASELINE.json north_star).

Rows are partitioned contiguously (`shard_bounds`); every rank builds/holds an independent sub-index over
its rows, all ranks search the SAME replicated query batch, and the per-shard top-k lists
(row_id u64, distance f32, count u32) are exchanged with one all-gather per array and merged on every rank
by `turdb_cuda_merge_topk_device` (ties ordered by (distance, row_id)).  torch.distributed is plumbing:
NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`: shard g = rows [g*N/G, (g+1)*N/G)."""
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def merge_topk_host(rows: np.ndarray, dist: np.ndarray, counts: np.ndarray, k: int):
    """Host mirror (and specification) of merge_topk_kernel for [n_shards, nq, k] arrays: a k-way merge of the
    per-shard lists, each consumed in its given (ascending-distance) order; at every step the head with the
    smallest (distance, row_id, shard) is taken."""
    n_shards, nq, kk = rows.shape
    out_rows = np.full((nq, k), np.uint64(2**64 - 1), np.uint64)
    out_dist = np.full((nq, k), np.inf, np.float32)
    out_cnt = np.zeros(nq, np.uint32)
    for q in range(nq):
        head = [0] * n_shards
        lim = [min(int(counts[s, q]), kk) for s in range(n_shards)]
        n_out = 0
        while n_out < k:
            best = None
            for s in range(n_shards):
                if head[s] < lim[s]:
                    key = (float(dist[s, q, head[s]]), int(rows[s, q, head[s]]), s)
                    if best is None or key < best:
                        best = key
            if best is None:
                break
            out_dist[q, n_out] = best[0]
            out_rows[q, n_out] = best[1]
            head[best[2]] += 1
            n_out += 1
        out_cnt[q] = n_out
    return out_rows, out_dist, out_cnt


class ShardedSearch:
    """search_batch over `world` sub-indexes: local search, all-gather, merge.

    `local_search(queries) -> (rows[nq,k] int64/uint64, dist[nq,k] f32, counts[nq] int32)` tensors on `device`;
    `merge(g_rows, g_dist, g_counts) -> (rows, dist, counts)` on the gathered [world, nq, k] tensors."""

    def __init__(self, dist_module, world: int, local_search, merge):
        self.dist = dist_module
        self.world = world
        self.local_search = local_search
        self.merge = merge
        self._bufs = None

    def search_batch(self, queries):
        import torch
        rows, dd, cnt = self.local_search(queries)
        if self.world == 1:
            return rows, dd, cnt
        w = self.world
        # outputs are the concatenation along dim 0 (the layout both NCCL and gloo accept), viewed [world, ...]
        if self._bufs is None or self._bufs[0].shape[0] != w * rows.shape[0]:
            self._bufs = (torch.empty((w * rows.shape[0],) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device),
                          torch.empty((w * dd.shape[0],) + tuple(dd.shape[1:]), dtype=dd.dtype, device=dd.device),
                          torch.empty((w * cnt.shape[0],), dtype=cnt.dtype, device=cnt.device))
        g_rows, g_dd, g_cnt = self._bufs
        self.dist.all_gather_into_tensor(g_rows, rows.contiguous())
        self.dist.all_gather_into_tensor(g_dd, dd.contiguous())
        self.dist.all_gather_into_tensor(g_cnt, cnt.contiguous())
        return self.merge(g_rows.view((w,) + tuple(rows.shape)), g_dd.view((w,) + tuple(dd.shape)),
                          g_cnt.view((w,) + tuple(cnt.shape)))
