"""One sub-index per GPU (BASELINE.json north_star).

Rows are partitioned contiguously (`shard_bounds`); every rank builds/holds an independent sub-index over
its rows, all ranks search the SAME replicated query batch, and the per-shard top-k lists
(row_id u64, distance f32, count u32) travel as ONE packed block per rank in a single all-gather and are merged on every rank
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


def pack_layout(nq: int, k: int) -> tuple[int, int, int, int]:
    """Byte offsets (rows, dist, counts) and padded size of one rank's packed result block:
    row ids [nq][k] u64 | distances [nq][k] f32 | counts [nq] u32, padded to 8 bytes."""
    o_rows, o_dist, o_cnt = 0, nq * k * 8, nq * k * 12
    return o_rows, o_dist, o_cnt, (o_cnt + nq * 4 + 7) & ~7


class ShardedSearch:
    """search_batch over `world` sub-indexes: local search, ONE all-gather, merge.

    `local_search(queries, rows, dist, counts)` writes the rank's top-k into the given tensors, which are views into this
    rank's packed block (rows int64 [nq,k], dist f32 [nq,k], counts int32 [nq]); the block is exchanged with a single
    `all_gather_into_tensor` ([world, block]) and `merge(gathered, block_bytes) -> (rows, dist, counts)` merges the
    per-shard lists (turdb_cuda_merge_topk_packed_device on GPUs, `merge_topk_host` over `unpack` in the CPU tests)."""

    def __init__(self, dist_module, world: int, local_search, merge, nq: int, k: int, device):
        import torch
        self.dist = dist_module
        self.world = world
        self.local_search = local_search
        self.merge = merge
        self.nq, self.k = nq, k
        o_rows, o_dist, o_cnt, size = pack_layout(nq, k)
        self.block_bytes = size
        self.block = torch.zeros(size, dtype=torch.uint8, device=device)
        self.rows = self.block[o_rows:o_dist].view(torch.int64).view(nq, k)
        self.dd = self.block[o_dist:o_cnt].view(torch.float32).view(nq, k)
        self.cnt = self.block[o_cnt:o_cnt + nq * 4].view(torch.int32)
        self.gathered = torch.zeros(world * size, dtype=torch.uint8, device=device) if world > 1 else None

    def unpack(self, gathered):
        """[world, block] bytes -> (rows [world,nq,k] int64, dist [world,nq,k] f32, counts [world,nq] int32) copies."""
        import torch
        o_rows, o_dist, o_cnt, size = pack_layout(self.nq, self.k)
        g = gathered.view(self.world, size)
        rows = torch.stack([g[s, o_rows:o_dist].contiguous().view(torch.int64).view(self.nq, self.k) for s in range(self.world)])
        dd = torch.stack([g[s, o_dist:o_cnt].contiguous().view(torch.float32).view(self.nq, self.k) for s in range(self.world)])
        cnt = torch.stack([g[s, o_cnt:o_cnt + self.nq * 4].contiguous().view(torch.int32) for s in range(self.world)])
        return rows, dd, cnt

    def search_batch(self, queries):
        self.local_search(queries, self.rows, self.dd, self.cnt)
        if self.world == 1:
            return self.rows, self.dd, self.cnt
        self.dist.all_gather_into_tensor(self.gathered, self.block)
        return self.merge(self.gathered, self.block_bytes)
