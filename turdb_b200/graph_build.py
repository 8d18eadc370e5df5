"""Bulk HNSW graph construction on the GPU — a BUILD-TIME UTILITY, not the search hot path.

The reference builds its graph one insert at a time (`insert_with_callback`, src/hnsw/mod.rs:999-1084):
~1.2 ms per insert single-threaded, i.e. hours for the 1M-100M corpora BASELINE.json names.  Benches
need those graphs in seconds, so this module builds a graph with the SAME structure the reference
stores — levels drawn by `select_level` (operations.rs:76-83), level-0 lists capped at 32, upper lists
at 16 (mod.rs:126-127), entry = first node of the top level — and fills the lists the way insertion
would, minus the sequential search: node i's forward candidates are its exact nearest neighbours AMONG
ITS PREDECESSORS (ids < i; insert_connection_phase searches the graph as it stood before i,
operations.rs:135-171 — this is what gives early nodes the long-range links that keep clusters connected),
selected with the reference's own `select_neighbors_heuristic` (operations.rs:181-233), then every
forward edge is offered back (append if room, re-select on overflow: the "reference-intent" mode of
SURVEY.md §7).  Distances are squared L2 on the raw vectors, as in the reference's insert path
(mod.rs:1031,1046).

Result files label this provenance "knn-heuristic"; graphs from the oracle's sequential restatement of
the insert path are labelled "reference-intent" / "verbatim".  Search parity (GPU kernel vs CPU
oracle) is always measured on the same arrays, whichever builder made them.

torch is used here for dense algebra (matmul/topk/sort); none of it runs inside a timed region.
"""
from __future__ import annotations

import math

import numpy as np
import torch

MAX_L0, MAX_UP, INVALID = 32, 16, 0xFFFFFFFF


def select_levels(randoms: np.ndarray, m: int = 16) -> np.ndarray:
    """select_level(random_value, 1/ln(m)) for a stream of random values in (0, 1]."""
    ml = 1.0 / math.log(float(m))
    lv = np.floor(-np.log(randoms.astype(np.float64)) * ml)
    return np.clip(np.nan_to_num(lv, nan=0.0), 0, 15).astype(np.uint8)


@torch.no_grad()
def _knn_level(xf: torch.Tensor, norms: torch.Tensor, ids: torch.Tensor, k: int, margin: int, chunk: int):
    """Exact k-NN (squared L2) of every node in `ids` among the nodes that precede it in `ids` (insertion order).
    Returns local positions [n_l,k] (-1 padded) and distances [n_l,k] (inf padded), ascending."""
    n_l = ids.numel()
    whole = n_l == xf.shape[0]
    cols = xf if whole else xf[ids]
    cn = norms if whole else norms[ids]
    out_pos = torch.full((n_l, k), -1, dtype=torch.int64, device=xf.device)
    out_d = torch.full((n_l, k), float("inf"), dtype=torch.float32, device=xf.device)
    ar = torch.arange(chunk, device=xf.device)
    for s in range(0, n_l, chunk):
        e = min(n_l, s + chunk)
        if e <= 1:
            continue
        rows = cols[s:e]
        pred = cols[: e - 1]  # the last row of the chunk has e-1 predecessors
        score = rows @ pred.T  # TF32 tensor-core pass (candidates only; re-evaluated in FP32 below)
        score.mul_(-2.0).add_(cn[None, : e - 1])
        # row i (global position s+i) may only see columns < s+i
        score.masked_fill_(torch.arange(e - 1, device=xf.device)[None, :] >= (ar[: e - s, None] + s), float("inf"))
        kc = min(k + margin, e - 1)
        cs, cand = torch.topk(score, kc, dim=1, largest=False, sorted=False)
        del score
        valid = torch.isfinite(cs)
        xc = cols[cand]
        de = ((xc - rows[:, None, :]) ** 2).sum(-1)
        de = torch.where(valid, de, torch.full_like(de, float("inf")))
        de, o = torch.sort(de, dim=1)
        cand = torch.gather(cand, 1, o)
        cand = torch.where(torch.isfinite(de), cand, torch.full_like(cand, -1))
        kk = min(k, kc)
        out_pos[s:e, :kk] = cand[:, :kk]
        out_d[s:e, :kk] = de[:, :kk]
    return out_pos, out_d


@torch.no_grad()
def _knn_level0_ivf(xf: torch.Tensor, norms: torch.Tensor, k: int, margin: int, n_cells: int, n_probe: int,
                    exact_prefix: int, chunk: int, seed: int):
    """Predecessor k-NN for corpora too large for the all-pairs pass (O(n^2) scores; 12.5M rows = 8e13 of them).
    Coarse partition: `n_cells` corpus points as centroids, every node in its nearest cell; a node's candidates are
    its predecessors inside the `n_probe` cells nearest to ITS CELL's centroid (one dense score block per cell).
    The first `exact_prefix` nodes still take the exact pass: they are the ones whose few predecessors are far away
    and give the graph its long-range links.  Same outputs as _knn_level (positions == node ids at level 0)."""
    n, dev = xf.shape[0], xf.device
    out_pos = torch.full((n, k), -1, dtype=torch.int64, device=dev)
    out_d = torch.full((n, k), float("inf"), dtype=torch.float32, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    cent_ids = torch.randperm(n, generator=gen)[:n_cells].to(dev)
    cent = xf[cent_ids]
    cn = norms[cent_ids]
    cell = torch.empty(n, dtype=torch.int64, device=dev)
    for s in range(0, n, 1 << 18):
        e = min(n, s + (1 << 18))
        cell[s:e] = torch.argmin(cn[None, :] - 2.0 * (xf[s:e] @ cent.T), dim=1)
    cc = cn[:, None] + cn[None, :] - 2.0 * (cent @ cent.T)
    probe = torch.topk(cc, min(n_probe, n_cells), dim=1, largest=False).indices  # [n_cells, n_probe], own cell first
    order = torch.argsort(cell, stable=True)  # ids ascending inside a cell
    counts = torch.bincount(cell, minlength=n_cells)
    starts = torch.cumsum(counts, 0) - counts
    starts_h, counts_h, probe_h = starts.cpu().tolist(), counts.cpu().tolist(), probe.cpu().tolist()
    for c in range(n_cells):
        if counts_h[c] == 0:
            continue
        col_ids = torch.cat([order[starts_h[p]:starts_h[p] + counts_h[p]] for p in probe_h[c]])
        cols, coln = xf[col_ids], norms[col_ids]
        for rs in range(0, counts_h[c], chunk):
            re = min(counts_h[c], rs + chunk)
            row_ids = order[starts_h[c] + rs:starts_h[c] + re]
            rows = xf[row_ids]
            score = rows @ cols.T
            score.mul_(-2.0).add_(coln[None, :])
            score.masked_fill_(col_ids[None, :] >= row_ids[:, None], float("inf"))  # predecessors only
            kc = min(k + margin, col_ids.numel())
            cs, cand = torch.topk(score, kc, dim=1, largest=False, sorted=False)
            del score
            valid = torch.isfinite(cs)
            cid = col_ids[cand]
            de = ((xf[cid] - rows[:, None, :]) ** 2).sum(-1)
            de = torch.where(valid, de, torch.full_like(de, float("inf")))
            de, o = torch.sort(de, dim=1)
            cid = torch.gather(cid, 1, o)
            cid = torch.where(torch.isfinite(de), cid, torch.full_like(cid, -1))
            kk = min(k, kc)
            out_pos[row_ids, :kk] = cid[:, :kk]
            out_d[row_ids, :kk] = de[:, :kk]
    p = min(exact_prefix, n)
    if p > 1:
        ids = torch.arange(p, device=dev)
        pp, pd = _knn_level(xf[:p], norms[:p], ids, min(k, p - 1), margin, chunk)
        out_pos[:p] = -1
        out_d[:p] = float("inf")
        out_pos[:p, :pp.shape[1]] = pp
        out_d[:p, :pd.shape[1]] = pd
    return out_pos, out_d


@torch.no_grad()
def _heuristic(cols: torch.Tensor, cand: torch.Tensor, cand_d: torch.Tensor, cap: int, chunk: int):
    """select_neighbors_heuristic over rows of ascending candidates (local positions, -1 / inf padded).
    Returns ids [B,cap] (-1 padded), d [B,cap], counts [B]; kept candidates first, then the back-fill."""
    dev = cand.device
    if cand.shape[1] < cap:  # fewer candidates than list slots (tiny top levels): pad
        pad = cap - cand.shape[1]
        cand = torch.cat([cand, torch.full((cand.shape[0], pad), -1, dtype=cand.dtype, device=dev)], 1)
        cand_d = torch.cat([cand_d, torch.full((cand_d.shape[0], pad), float("inf"), dtype=cand_d.dtype, device=dev)], 1)
    b_all, k = cand.shape
    out = torch.full((b_all, cap), -1, dtype=torch.int64, device=dev)
    out_d = torch.full((b_all, cap), float("inf"), dtype=torch.float32, device=dev)
    cnts = torch.zeros(b_all, dtype=torch.int64, device=dev)
    jj = torch.arange(k, device=dev)[None, :]
    for s in range(0, b_all, chunk):
        e = min(b_all, s + chunk)
        c = cand[s:e]
        d = cand_d[s:e]
        valid = c >= 0
        xc = cols[c.clamp(min=0)]
        nc = (xc * xc).sum(-1)
        dcc = nc[:, :, None] + nc[:, None, :] - 2.0 * torch.bmm(xc, xc.transpose(1, 2))
        del xc
        kept = torch.zeros_like(valid)
        cnt = torch.zeros(e - s, dtype=torch.int64, device=dev)
        for j in range(k):
            bad = ((dcc[:, j, :] < d[:, j, None]) & kept).any(1)
            ok = valid[:, j] & ~bad & (cnt < cap)
            kept[:, j] = ok
            cnt += ok
        notkept = valid & ~kept
        fill = notkept & ((torch.cumsum(notkept, 1) - 1) < (cap - cnt)[:, None])
        key = torch.where(kept, jj, torch.where(fill, jj + k, jj + 3 * k))
        o = torch.argsort(key, dim=1)[:, :cap]
        take = torch.gather(kept | fill, 1, o)
        out[s:e] = torch.where(take, torch.gather(c, 1, o), torch.full_like(o, -1))
        out_d[s:e] = torch.where(take, torch.gather(d, 1, o), torch.full_like(o, float("inf"), dtype=torch.float32))
        cnts[s:e] = take.sum(1)
    return out, out_d, cnts


@torch.no_grad()
def _level_lists(xf, norms, ids, cap, knn_k, rev_cap, chunk_rows, chunk_h, ivf=None):
    """Neighbour lists (local positions, -1 padded) for one level.  ivf = (n_cells, n_probe, exact_prefix, seed)
    switches the candidate pass of a whole-corpus level to the partitioned form."""
    dev = xf.device
    n_l = ids.numel()
    if n_l <= 1:
        return torch.full((n_l, cap), -1, dtype=torch.int64, device=dev), torch.zeros(n_l, dtype=torch.int64, device=dev)
    whole = n_l == xf.shape[0]
    cols = xf if whole else xf[ids]
    k = min(knn_k, n_l - 1)
    if ivf is not None and whole:
        pos, d = _knn_level0_ivf(xf, norms, k, 16, ivf[0], ivf[1], ivf[2], chunk_rows, ivf[3])
    else:
        pos, d = _knn_level(xf, norms, ids, k, margin=16, chunk=chunk_rows)  # forward candidates: predecessors only
    f_ids, f_d, f_cnt = _heuristic(cols, pos, d, cap, chunk_h)
    del pos, d
    # ---- back-links: dst <- src for every forward edge src -> dst that dst does not already hold ----
    src = torch.arange(n_l, device=dev)[:, None].expand(n_l, cap).reshape(-1)
    dst = f_ids.reshape(-1)
    dd = f_d.reshape(-1)
    ok = dst >= 0
    src, dst, dd = src[ok], dst[ok], dd[ok]
    keep = torch.empty(src.numel(), dtype=torch.bool, device=dev)
    step = 1 << 22
    for s in range(0, src.numel(), step):
        e = min(src.numel(), s + step)
        keep[s:e] = ~(f_ids[dst[s:e]] == src[s:e, None]).any(1)
    src, dst, dd = src[keep], dst[keep], dd[keep]
    o = torch.argsort(dd)
    src, dst, dd = src[o], dst[o], dd[o]
    o = torch.argsort(dst, stable=True)
    src, dst, dd = src[o], dst[o], dd[o]
    counts = torch.bincount(dst, minlength=n_l)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(dst.numel(), device=dev) - starts[dst]
    sel = rank < rev_cap
    r_ids = torch.full((n_l, rev_cap), -1, dtype=torch.int64, device=dev)
    r_d = torch.full((n_l, rev_cap), float("inf"), dtype=torch.float32, device=dev)
    r_ids[dst[sel], rank[sel]] = src[sel]
    r_d[dst[sel], rank[sel]] = dd[sel]
    r_cnt = torch.clamp(counts, max=rev_cap)
    total = f_cnt + r_cnt
    # room for every back-link: append in arrival (distance) order, as add_neighbor_at_level does
    c_ids = torch.cat([f_ids, r_ids], 1)
    c_d = torch.cat([f_d, r_d], 1)
    key = torch.where(c_ids >= 0, torch.arange(cap + rev_cap, device=dev)[None, :], cap + rev_cap)
    o = torch.argsort(key, dim=1, stable=True)[:, :cap]
    out = torch.gather(c_ids, 1, o)
    cnt = torch.clamp(total, max=cap)
    # overflow: re-select over (list + back-links) sorted by distance to the owner
    over = torch.nonzero(total > cap).squeeze(1)
    if over.numel():
        oc_d, oo = torch.sort(c_d[over], dim=1)
        oc = torch.gather(c_ids[over], 1, oo)
        sel_ids, _, sel_cnt = _heuristic(cols, oc, oc_d, cap, chunk_h)
        out[over] = sel_ids
        cnt[over] = sel_cnt
    return out, cnt


@torch.no_grad()
def build_graph(vectors: np.ndarray, m: int = 16, seed: int = 1234, device: str | torch.device = "cuda:0",
                knn_k: int = 64, row_ids: np.ndarray | None = None, chunk_rows: int = 4096,
                l0_rev: int = MAX_L0, up_rev: int = 64, ivf_cells: int | None = None, ivf_probe: int = 16,
                ivf_exact_prefix: int = 1 << 20) -> dict:
    """vectors [n, dim] f32 -> the flattened graph dict `CudaHnswIndex.from_graph` / the oracle take.
    ivf_cells: None = all-pairs candidate pass up to 3M rows, partitioned pass (~sqrt(n) cells) above; 0 = always
    all-pairs; > 0 = that many cells.  ivf_exact_prefix: the first that many nodes take the exact predecessor pass —
    they carry the long-range links; measured on 12.5M x 128 clustered (sigma 0.3): recall@10 at ef 128 is 0.78 with
    a 131k prefix and 0.96 with a 1M prefix (tools/cluster_probe.py)."""
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    n, dim = vectors.shape
    dev = torch.device(device)
    prev_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        rng = np.random.default_rng(seed)
        levels = select_levels(1.0 - rng.random(n), m)
        max_level = int(levels.max()) if n else 0
        xf = torch.from_numpy(vectors).to(dev)
        norms = (xf * xf).sum(1)
        lv = torch.from_numpy(levels.astype(np.int64)).to(dev)
        chunk_h = max(256, min(8192, (1 << 28) // max(1, (knn_k + max(l0_rev, up_rev)) * dim)))
        if ivf_cells is None:
            ivf_cells = 0 if n <= 3_000_000 else 1 << max(6, int(round(math.log2(math.sqrt(n)))))
        ivf = (min(ivf_cells, n), ivf_probe, ivf_exact_prefix, seed) if ivf_cells else None
        l0, l0_cnt = _level_lists(xf, norms, torch.arange(n, device=dev), MAX_L0, knn_k, l0_rev, chunk_rows, chunk_h, ivf)
        l0_adj = torch.where(l0 >= 0, l0, torch.full_like(l0, INVALID)).to(torch.int64).cpu().numpy().astype(np.uint32)
        n_slots = int(levels.astype(np.int64).sum())
        up_base = np.full(n, INVALID, np.uint32)
        has_up = levels > 0
        base = np.cumsum(levels.astype(np.int64)) - levels
        up_base[has_up] = base[has_up].astype(np.uint32)
        up_adj = np.full((n_slots, MAX_UP), INVALID, np.uint32)
        up_cnt = np.zeros(n_slots, np.uint8)
        for level in range(1, max_level + 1):
            ids = torch.nonzero(lv >= level).squeeze(1)
            lists, cnt = _level_lists(xf, norms, ids, MAX_UP, 2 * m, up_rev, chunk_rows, chunk_h)
            glob = torch.where(lists >= 0, ids[lists.clamp(min=0)], torch.full_like(lists, INVALID))
            ids_np = ids.cpu().numpy()
            slots = base[ids_np] + (level - 1)
            up_adj[slots] = glob.cpu().numpy().astype(np.uint32)
            up_cnt[slots] = cnt.cpu().numpy().astype(np.uint8)
        entry = int(np.flatnonzero(levels == max_level)[0]) if n else INVALID
        return dict(
            vectors=vectors,
            row_ids=np.arange(n, dtype=np.uint64) if row_ids is None else np.ascontiguousarray(row_ids, np.uint64),
            levels=levels, l0_adj=l0_adj, l0_cnt=l0_cnt.cpu().numpy().astype(np.uint8), up_base=up_base,
            up_adj=up_adj, up_cnt=up_cnt, entry=entry, max_level=max_level,
            provenance="predecessor-knn-heuristic" if not ivf else f"predecessor-knn-heuristic(ivf {ivf[0]}x{ivf[1]})")
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        if dev.type == "cuda":
            # hand the builder's cached blocks back: the index is uploaded by libturdb_cuda's own cudaMalloc next
            xf = norms = lv = None  # noqa: F841
            torch.cuda.empty_cache()
