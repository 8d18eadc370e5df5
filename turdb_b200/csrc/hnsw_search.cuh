// hnsw_search.cuh — the traversal kernel: greedy upper-level descent + level-0 beam search,
// one warp per query (PersistentHnswIndex::search, src/hnsw/mod.rs:1092-1174;
// greedy_search src/hnsw/search.rs:259-309; beam_search src/hnsw/search.rs:311-350).
//
// Per hop the warp (1) picks the closest unexpanded entry of its sorted result list, (2) reads that
// node's 32-wide adjacency row with one coalesced 128 B load, (3) filters it through an exact
// visited set in shared memory, (4) gathers ALL unvisited neighbour vectors at once with one TMA
// bulk copy each (cp.async.bulk -> shared memory, mbarrier completion), (5) reduces distances in
// the reference's AVX2 lane order (bit-identical values), and (6) merges the <=32 new (distance,id)
// pairs into the sorted ef-slot list by rank counting.
//
// Single-list equivalence with the reference's two heaps: SURVEY.md Appendix D / DESIGN.md §5.
#pragma once

#include "common.cuh"

namespace turdb {

struct WarpLayout {
  uint32_t off_bar, off_q, off_list, off_tmp, off_hash, off_stage;
  uint32_t warp_bytes;
  uint32_t n_slots;    // staging slots (multiple of 8, <= 32)
  uint32_t stride;     // bytes between staging slots; stride/4 == 8 (mod 32) -> conflict-free quads
  uint32_t vec_bytes;  // ds * 4
  uint32_t hash_bits;  // shared visited table has 1 << hash_bits slots
  // compact table (hash16 != 0): 16-bit entries = (displacement+1) << rem_bits | remainder of a
  // bijective hash of the id, so an entry still identifies exactly one node (exact set, half the bytes)
  uint32_t hash16;
  uint32_t rem_bits;   // key_bits - hash_bits
  uint32_t key_bits;   // ceil(log2(n)), >= hash_bits
};

struct SearchArgs {
  DeviceIndex ix;
  WarpLayout lay;
  const float* queries;  // [nq][dim]
  uint32_t nq, k, ef;
  uint64_t* out_row_ids;  // [nq][k]
  uint32_t* out_node_ids; // [nq][k] or null
  float* out_dist;        // [nq][k]
  uint32_t* out_counts;   // [nq]
  uint32_t* out_stats;    // [nq][4] or null
  uint32_t* work_counter; // zeroed before launch
  uint32_t* overflow_count;  // queries whose shared visited table filled up
  uint32_t* overflow_list;   // [nq]
  uint32_t* global_visited;  // fallback pass: [resident warps][vis_words] bitsets
  uint32_t vis_words;
};

struct WarpCtx {
  uint32_t lane;
  uint32_t bar0;
  const float* q;
  const uint8_t* stage;
  uint32_t stage_u32;
  uint32_t stride, vec_bytes, n_slots;
  uint32_t phases;
  float qnorm;
};

// Distances from the query to the m vectors whose ids sit in lanes 0..m-1 (`cid`).
// Returns d in lane j for id j; +inf in lanes >= m.
//
// The m vectors are gathered in chunks of 8 (one quad of lanes per vector).  Chunk c lands in staging
// group c % G (G = n_slots / 8, one mbarrier per group); the first G chunks are requested up front and
// group g is re-armed with chunk c + G as soon as chunk c has been reduced, so up to n_slots vectors
// stay in flight for the whole hop.
template <int METRIC>
__device__ __forceinline__ float gather_distances(const DeviceIndex& ix, WarpCtx& w, uint32_t cid,
                                                  uint32_t m) {
  const uint32_t lane = w.lane, p = lane & 3, sub = lane & 7, my_chunk = lane >> 3;
  const uint32_t G = w.n_slots >> 3;
  const uint32_t nchunks = (m + 7) >> 3;
  const bool have = lane < m;
  const float* src = ix.arena + (size_t)(have ? cid : 0) * ix.ds;
  float raw = 0.f, nb = 0.f;
  if (METRIC == kCosine && have) nb = __ldg(ix.norm2 + cid);

  auto issue = [&](uint32_t c) {
    const uint32_t g = c % G;
    const uint32_t bar = w.bar0 + 8 * g;
    if (have && my_chunk == c && sub == 0) mbar_expect_tx(bar, min(8u, m - 8 * c) * w.vec_bytes);
    __syncwarp();
    if (have && my_chunk == c) bulk_g2s(w.stage_u32 + (g * 8 + sub) * w.stride, src, w.vec_bytes, bar);
  };
  const uint32_t pro = min(G, nchunks);
  for (uint32_t c = 0; c < pro; ++c) issue(c);
  for (uint32_t c = 0; c < nchunks; ++c) {
    const uint32_t g = c % G;
    mbar_wait(w.bar0 + 8 * g, (w.phases >> g) & 1u);
    w.phases ^= (1u << g);
    const float* b = reinterpret_cast<const float*>(w.stage + (g * 8 + (lane >> 2)) * w.stride);
    const float r = (METRIC == kL2) ? quad_l2sq(w.q, b, ix.dim, p) : quad_dot(w.q, b, ix.dim, p);
    const float t = __shfl_sync(kFullMask, r, sub * 4);
    if (my_chunk == c) raw = t;
    if (c + G < nchunks) issue(c + G);
  }
  __syncwarp();
  if (!have) return INFINITY;
  if (METRIC == kL2) return raw;
  if (METRIC == kIP) return -raw;  // inner_product_avx2, distance.rs:240-242
  return cosine_finish(raw, w.qnorm, nb);
}

// Exact visited set.  Returns 1 = newly inserted, 0 = already present, 2 = table cannot place the key
// (the query is then redone by the global-bitset pass).
//   GLOBAL : one bit per node in global memory.
//   shared, 32-bit entries: open addressing on the full id.
//   shared, 16-bit entries: h = id * odd (mod 2^key_bits) is a bijection; home = top hash_bits of h, the
//     entry stores the remaining rem_bits plus (displacement + 1), which together name h and so the id.
template <bool GLOBAL>
__device__ __forceinline__ uint32_t visited_insert(uint32_t* tab, uint32_t id, const WarpLayout& L) {
  if (GLOBAL) {
    const uint32_t bit = 1u << (id & 31);
    return (atomicOr(tab + (id >> 5), bit) & bit) == 0 ? 1u : 0u;
  }
  const uint32_t mask = (1u << L.hash_bits) - 1;
  if (L.hash16) {
    unsigned short* t16 = reinterpret_cast<unsigned short*>(tab);
    const uint32_t h = (id * 0x9E3779B1u) & ((1u << L.key_bits) - 1);
    const uint32_t home = h >> L.rem_bits, rem = h & ((1u << L.rem_bits) - 1);
    const uint32_t max_disp = (1u << (16 - L.rem_bits)) - 1;  // displacement+1 must fit
    for (uint32_t disp = 0; disp < max_disp; ++disp) {
      const unsigned short want = (unsigned short)(((disp + 1) << L.rem_bits) | rem);
      const unsigned short old = atomicCAS(t16 + ((home + disp) & mask), (unsigned short)0, want);
      if (old == 0) return 1u;
      if (old == want) return 0u;
    }
    return 2u;
  }
  uint32_t slot = (id * 0x9E3779B1u) >> (32 - L.hash_bits);
  for (;;) {
    const uint32_t old = atomicCAS(tab + slot, kInvalid, id);
    if (old == kInvalid) return 1u;
    if (old == id) return 0u;
    slot = (slot + 1) & mask;
  }
}

template <int METRIC, bool GLOBAL_VISITED>
__global__ void __launch_bounds__(256, 1) hnsw_search_kernel(const SearchArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const DeviceIndex& ix = a.ix;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* ws = smem + (size_t)warp * a.lay.warp_bytes;
  float* qs = reinterpret_cast<float*>(ws + a.lay.off_q);
  float* A_d = reinterpret_cast<float*>(ws + a.lay.off_list);
  uint32_t* A_id = reinterpret_cast<uint32_t*>(A_d + a.ef);
  float* B_d = reinterpret_cast<float*>(A_id + a.ef);
  uint32_t* B_id = reinterpret_cast<uint32_t*>(B_d + a.ef);
  uint32_t* tmp_ids = reinterpret_cast<uint32_t*>(ws + a.lay.off_tmp);
  uint32_t* tmp_ub = tmp_ids + 32;
  uint32_t* vis = GLOBAL_VISITED
                      ? a.global_visited + (size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * a.vis_words
                      : reinterpret_cast<uint32_t*>(ws + a.lay.off_hash);
  const uint32_t hash_slots = 1u << a.lay.hash_bits;
  const uint32_t hash_limit = hash_slots - (hash_slots >> 3);  // 7/8 full -> overflow path

  WarpCtx w;
  w.lane = lane;
  w.bar0 = smem_u32(ws + a.lay.off_bar);
  w.q = qs;
  w.stage = ws + a.lay.off_stage;
  w.stage_u32 = smem_u32(w.stage);
  w.stride = a.lay.stride;
  w.vec_bytes = a.lay.vec_bytes;
  w.n_slots = a.lay.n_slots;
  w.phases = 0;
  w.qnorm = 0.f;

  if (lane == 0) {
    for (uint32_t g = 0; g < 4; ++g) mbar_init(w.bar0 + 8 * g, 1);
    mbar_fence_init();
  }
  __syncwarp();

  const uint32_t ef = a.ef;
  const uint32_t n_work = GLOBAL_VISITED ? *a.overflow_count : a.nq;

  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(a.work_counter, 1u);
    wi = __shfl_sync(kFullMask, wi, 0);
    if (wi >= n_work) break;
    const uint32_t qi = GLOBAL_VISITED ? a.overflow_list[wi] : wi;

    // ---- stage the query, clear the visited set ----
    const float* qg = a.queries + (size_t)qi * ix.dim;
    for (uint32_t i = lane; i < ix.ds; i += 32) qs[i] = i < ix.dim ? __ldg(qg + i) : 0.f;
    if (GLOBAL_VISITED) {
      uint4* v4 = reinterpret_cast<uint4*>(vis);
      for (uint32_t i = lane; i < (a.vis_words >> 2); i += 32) v4[i] = make_uint4(0, 0, 0, 0);
    } else if (a.lay.hash16) {
      uint4* v4 = reinterpret_cast<uint4*>(vis);
      for (uint32_t i = lane; i < (hash_slots >> 3); i += 32) v4[i] = make_uint4(0, 0, 0, 0);
    } else {
      uint4* v4 = reinterpret_cast<uint4*>(vis);
      for (uint32_t i = lane; i < (hash_slots >> 2); i += 32)
        v4[i] = make_uint4(kInvalid, kInvalid, kInvalid, kInvalid);
    }
    __syncwarp();
    if (METRIC == kCosine) w.qnorm = quad_dot(qs, qs, ix.dim, lane & 3);

    uint32_t n_dist = 0, n_dist_upper = 0, n_expanded = 0, n_upper_hops = 0;
    uint32_t len = 0;
    bool overflow = false;

    if (ix.entry != kInvalid && ix.n != 0) {
      // ---- entry distance (mod.rs:1129) ----
      uint32_t cur = ix.entry;
      float cur_d = gather_distances<METRIC>(ix, w, lane == 0 ? cur : kInvalid, 1);
      cur_d = __shfl_sync(kFullMask, cur_d, 0);
      n_dist = 1;
      n_dist_upper = 1;

      // ---- greedy descent over levels max_level..1 (mod.rs:1134-1145, search.rs:259-309) ----
      for (uint32_t level = ix.max_level; level >= 1; --level) {
        for (uint32_t it = 0; it < 1000; ++it) {
          const uint32_t lv = ix.levels[cur];
          const uint32_t ub = ix.up_base[cur];
          uint32_t nid = kInvalid;
          if (level <= lv && ub != kInvalid && lane < kUp)
            nid = __ldg(ix.up_adj + ((size_t)ub + level - 1) * kUp + lane);
          n_upper_hops += 1;
          const uint32_t m = __popc(__ballot_sync(kFullMask, nid != kInvalid));
          if (m == 0) break;
          float d = gather_distances<METRIC>(ix, w, nid, m);
          n_dist += m;
          n_dist_upper += m;
          // arg-min, strict `<`, first stored neighbour wins ties (search.rs:272-277)
          float best = d;
          uint32_t bl = lane;
#pragma unroll
          for (uint32_t off = 16; off >= 1; off >>= 1) {
            float od = __shfl_xor_sync(kFullMask, best, off);
            uint32_t ol = __shfl_xor_sync(kFullMask, bl, off);
            if (od < best || (od == best && ol < bl)) {
              best = od;
              bl = ol;
            }
          }
          if (!(best < cur_d)) break;
          cur = __shfl_sync(kFullMask, nid, bl);
          cur_d = best;
        }
      }

      // ---- level-0 beam search (search.rs:311-350) on one sorted list ----
      if (lane == 0) {
        A_d[0] = cur_d;
        A_id[0] = cur;
        (void)visited_insert<GLOBAL_VISITED>(vis, cur, a.lay);
      }
      len = 1;
      uint32_t n_visited = 1;
      uint32_t spec_node = kInvalid, spec_nid = kInvalid;  // speculatively fetched adjacency row
      __syncwarp();

      for (;;) {
        // closest unexpanded entry == the reference's candidates.pop() that passes `d <= worst`;
        // the runner-up is the likely next hop: its adjacency row is fetched while this hop gathers.
        uint32_t i1 = 0xFFFFFFFFu, i2 = 0xFFFFFFFFu;
        for (uint32_t i = lane; i < len; i += 32)
          if (!(A_id[i] & kExpandedBit)) {
            if (i1 == 0xFFFFFFFFu) i1 = i;
            else {
              i2 = i;
              break;
            }
          }
        const uint32_t idx = __reduce_min_sync(kFullMask, i1);
        if (idx >= len) break;
        const uint32_t idx2 = __reduce_min_sync(kFullMask, i1 == idx ? i2 : i1);
        const uint32_t c = A_id[idx];
        const uint32_t c2 = idx2 < len ? A_id[idx2] : kInvalid;
        __syncwarp();
        if (lane == 0) A_id[idx] = c | kExpandedBit;
        __syncwarp();
        n_expanded += 1;

        if (!GLOBAL_VISITED && n_visited + kL0 > hash_limit) {
          overflow = true;
          break;
        }
        const uint32_t nid = (c == spec_node) ? spec_nid : __ldg(ix.l0_adj + (size_t)c * kL0 + lane);
        if (c2 != kInvalid) {
          spec_node = c2;
          spec_nid = __ldg(ix.l0_adj + (size_t)c2 * kL0 + lane);
        } else {
          spec_node = kInvalid;
        }
        const uint32_t ins = (nid != kInvalid) ? visited_insert<GLOBAL_VISITED>(vis, nid, a.lay) : 0u;
        if (__any_sync(kFullMask, ins == 2u)) {
          overflow = true;
          break;
        }
        const bool isnew = ins == 1u;
        const uint32_t newmask = __ballot_sync(kFullMask, isnew);
        const uint32_t m = __popc(newmask);
        if (m == 0) continue;
        n_visited += m;
        // compact to lanes 0..m-1 in stored order
        if (isnew) tmp_ids[__popc(newmask & ((1u << lane) - 1))] = nid;
        __syncwarp();
        const uint32_t cid = lane < m ? tmp_ids[lane] : kInvalid;
        const float d = gather_distances<METRIC>(ix, w, cid, m);
        n_dist += m;

        // admission (search.rs:344): d < worst || results.len() < ef, applied to the batch
        const float worst = (len == ef) ? A_d[len - 1] : INFINITY;
        const bool elig = (lane < m) && (len < ef || d < worst);
        const uint32_t emask = __ballot_sync(kFullMask, elig);
        const uint32_t mp = __popc(emask);
        if (mp == 0) continue;

        // insertion point among old entries: old entries win ties (they were admitted first)
        uint32_t ub = 0;
        if (elig) {
          uint32_t lo = 0, hi = len;
          while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (A_d[mid] <= d) lo = mid + 1;
            else hi = mid;
          }
          ub = lo;
          tmp_ub[__popc(emask & ((1u << lane) - 1))] = ub;
        }
        // rank among the new entries: (distance, stored order)
        uint32_t rank = 0;
        for (uint32_t e = emask; e; e &= e - 1) {
          const uint32_t bsrc = __ffs(e) - 1;
          const float db = __shfl_sync(kFullMask, d, bsrc);
          rank += (db < d || (db == d && bsrc < lane)) ? 1u : 0u;
        }
        __syncwarp();
        for (uint32_t i = lane; i < len; i += 32) {
          uint32_t s = 0;
          for (uint32_t j = 0; j < mp; ++j) s += (tmp_ub[j] <= i) ? 1u : 0u;
          const uint32_t np = i + s;
          if (np < ef) {
            B_d[np] = A_d[i];
            B_id[np] = A_id[i];
          }
        }
        if (elig) {
          const uint32_t np = ub + rank;
          if (np < ef) {
            B_d[np] = d;
            B_id[np] = cid;
          }
        }
        len = min(len + mp, ef);
        float* td = A_d; A_d = B_d; B_d = td;
        uint32_t* ti = A_id; A_id = B_id; B_id = ti;
        __syncwarp();
      }
    }

    if (overflow) {
      // shared visited table filled: hand the query to the global-bitset pass (exact, rare)
      if (lane == 0) a.overflow_list[atomicAdd(a.overflow_count, 1u)] = qi;
      continue;
    }

    // ---- finalize_results(k) (search.rs:245-252) + SearchResult mapping (mod.rs:1159-1171) ----
    const uint32_t count = min(len, a.k);
    for (uint32_t i = lane; i < a.k; i += 32) {
      const size_t o = (size_t)qi * a.k + i;
      if (i < count) {
        const uint32_t id = A_id[i] & ~kExpandedBit;
        a.out_row_ids[o] = ix.row_ids[id];
        if (a.out_node_ids) a.out_node_ids[o] = id;
        a.out_dist[o] = A_d[i];
      } else {
        a.out_row_ids[o] = 0xFFFFFFFFFFFFFFFFull;
        if (a.out_node_ids) a.out_node_ids[o] = kInvalid;
        a.out_dist[o] = INFINITY;
      }
    }
    if (lane == 0) {
      a.out_counts[qi] = count;
      if (a.out_stats) {
        uint32_t* s = a.out_stats + (size_t)qi * 4;
        s[0] = n_dist;
        s[1] = n_dist_upper;
        s[2] = n_expanded;
        s[3] = n_upper_hops;
      }
    }
    __syncwarp();
  }
}

}  // namespace turdb
