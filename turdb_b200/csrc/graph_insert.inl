// graph_insert.inl — the reference's INSERT path on the device: PersistentHnswIndex::insert_with_callback
// (src/hnsw/mod.rs:999-1084) with insert_descent_phase / insert_connection_phase (src/hnsw/operations.rs:111-171),
// the neighbour-side add_neighbor_at_level (mod.rs:275-301) and, in reference-intent mode, select_neighbors_heuristic
// (operations.rs:181-233) for a full back-link list.  Included by turdb_cuda.cu.  SURVEY.md §8(a) "insert path".
//
// Nodes are inserted in id order in STEPS of `batch` consecutive nodes: the searches of a step (greedy descent + one
// ef_construction beam per level, hnsw_insert_search_*kernel: the traversal kernel in INSERT mode) all see the graph as
// it stood when the step began; their links are then applied in id order (own lists, then back-links grouped by
// target list and replayed in id order by one warp per list).  batch == 1 IS the reference's sequential procedure —
// the graph equals the oracle's bit for bit (tests/test_gpu_build.py); larger steps trade that for throughput (nodes
// of one step do not link to one another), the standard batched construction.
//
// Reference quirks kept: the connection loop runs target_level..0 even above the current max_level (a
// one-way link to the old entry); the entry point is NOT refined between levels; selection = the beam's nearest
// m0 = 2M (level 0) / M results, the new node keeps the first 32 / 16, every selected neighbour gets a back-link;
// a back-link to a neighbour that lacks the level is dropped; distances are squared L2 whatever the metric.

#include <cub/device/device_radix_sort.cuh>

namespace turdb {

struct InsertLinkArgs {
  uint32_t first, count;       // nodes [first, first + count) of this step
  uint64_t n_total;            // list index of upper slot s = n_total + s
  const uint8_t* levels;       // [n_total] target levels (all nodes, known up front)
  const uint32_t* up_base;     // [n_total]
  uint32_t* l0_adj;            // [n_total][32]
  uint8_t* l0_cnt;             // [n_total]
  uint32_t* up_adj;            // [slots][16]
  uint8_t* up_cnt;             // [slots]
  const uint32_t* ins_sel;     // [count][kInsLevels][kInsSelMax]
  const uint8_t* ins_cnt;      // [count][kInsLevels]
  unsigned long long* req_keys;  // back-link requests: list_index << 20 | local node << 6 | position
  uint32_t* req_count;
  uint32_t req_cap;
};

// One warp per new node: its own lists (add_neighbor_at_level on the new node: first cap entries, mod.rs:1067-1072) and
// one back-link request per selected neighbour that has the level (mod.rs:1073-1076, :293-301).
__global__ void insert_link_own_kernel(const InsertLinkArgs a) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= a.count) return;
  const uint32_t id = a.first + w;
  const uint32_t target = a.levels[id];
  for (uint32_t l = 0; l <= target; ++l) {
    const uint32_t cnt = a.ins_cnt[(size_t)w * kInsLevels + l];
    const uint32_t* sel = a.ins_sel + ((size_t)w * kInsLevels + l) * kInsSelMax;
    const uint32_t cap = l == 0 ? kL0 : kUp;
    uint32_t* own = l == 0 ? a.l0_adj + (size_t)id * kL0 : a.up_adj + ((size_t)a.up_base[id] + l - 1) * kUp;
    for (uint32_t i = lane; i < cap; i += 32) own[i] = i < cnt ? sel[i] : kInvalid;
    if (lane == 0) {
      if (l == 0) a.l0_cnt[id] = (uint8_t)min(cnt, cap);
      else a.up_cnt[a.up_base[id] + l - 1] = (uint8_t)min(cnt, cap);
    }
    for (uint32_t base = 0; base < cnt; base += 32) {
      const uint32_t i = base + lane;
      uint32_t nb = kInvalid;
      bool ok = false;
      if (i < cnt) {
        nb = sel[i];
        ok = l <= a.levels[nb];  // "neighbour lacks this level": dropped (mod.rs:296)
      }
      const uint32_t mask = __ballot_sync(kFullMask, ok);
      uint32_t pos0 = 0;
      if (lane == 0 && mask) pos0 = atomicAdd(a.req_count, __popc(mask));
      pos0 = __shfl_sync(kFullMask, pos0, 0);
      if (ok) {
        const uint32_t pos = pos0 + __popc(mask & ((1u << lane) - 1));
        const unsigned long long list = l == 0 ? (unsigned long long)nb : a.n_total + a.up_base[nb] + l - 1;
        if (pos < a.req_cap) a.req_keys[pos] = (list << 20) | ((unsigned long long)w << 6) | i;
      }
    }
  }
}

struct InsertBacklinkArgs {
  DeviceIndex ix;              // arena (vectors of ALL nodes are resident from the start)
  uint32_t first;
  uint64_t n_total;
  const unsigned long long* req_keys;  // sorted ascending: by list, then node, then position
  const uint32_t* req_count;
  uint32_t req_cap;
  const uint32_t* up_owner;    // [slots] node that owns upper slot s
  uint32_t* l0_adj;
  uint8_t* l0_cnt;
  uint32_t* up_adj;
  uint8_t* up_cnt;
  int intent;                  // 1: full list -> select_neighbors_heuristic; 0: verbatim (dropped)
};

// One warp per back-link request; only the warp holding the FIRST request of a list works: it replays that list's
// requests in id order — append while there is room (mod.rs:275-280), otherwise (reference-intent) re-select the list
// from (current entries + the new node) with select_neighbors_heuristic (operations.rs:181-233).
__global__ void __launch_bounds__(128) insert_backlink_kernel(const InsertBacklinkArgs a) {
  __shared__ uint32_t s_cid[4][kL0 + 1], s_sid[4][kL0 + 1], s_sel[4][kL0];
  __shared__ float s_cd[4][kL0 + 1], s_sd[4][kL0 + 1];
  const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31, p = lane & 3, quad = lane >> 2;
  const uint32_t r0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_req = min(*a.req_count, a.req_cap);
  if (r0 >= n_req) return;
  const unsigned long long list = a.req_keys[r0] >> 20;
  if (r0 > 0 && (a.req_keys[r0 - 1] >> 20) == list) return;  // not the head of its run
  const bool is_l0 = list < a.n_total;
  const uint32_t owner = is_l0 ? (uint32_t)list : a.up_owner[list - a.n_total];
  const uint32_t cap = is_l0 ? kL0 : kUp;
  uint32_t* row = is_l0 ? a.l0_adj + (size_t)owner * kL0 : a.up_adj + (size_t)(list - a.n_total) * kUp;
  uint8_t* pcnt = is_l0 ? a.l0_cnt + owner : a.up_cnt + (list - a.n_total);
  uint32_t cnt = *pcnt;
  const uint32_t dim = a.ix.dim;
  const float* vo = a.ix.arena + (size_t)owner * a.ix.ds;
  uint32_t* cid = s_cid[wib];
  uint32_t* sid = s_sid[wib];
  uint32_t* sel = s_sel[wib];
  float* cd = s_cd[wib];
  float* sd = s_sd[wib];
  for (uint32_t r = r0; r < n_req && (a.req_keys[r] >> 20) == list; ++r) {
    const uint32_t id = a.first + (uint32_t)((a.req_keys[r] >> 6) & 0x3FFFu);
    if (cnt < cap) {
      if (lane == 0) row[cnt] = id;
      cnt += 1;
      __syncwarp();
      continue;
    }
    if (!a.intent) continue;  // verbatim: silently dropped when full
    // candidates: the current list in stored order, then the new node; distance to the owner (squared L2, AVX2 order)
    const uint32_t nc = cnt + 1;
    for (uint32_t i = lane; i < nc; i += 32) cid[i] = i < cnt ? row[i] : id;
    __syncwarp();
    for (uint32_t base = 0; base < nc; base += 8) {
      const uint32_t j = min(base + quad, nc - 1);
      const float d = quad_l2sq(vo, a.ix.arena + (size_t)cid[j] * a.ix.ds, dim, p);
      if (p == 0 && base + quad < nc) cd[base + quad] = d;
    }
    __syncwarp();
    // stable ascending sort by distance (rank sort; ties keep stored order)
    for (uint32_t j = lane; j < nc; j += 32) {
      const float dj = cd[j];
      uint32_t rank = 0;
      for (uint32_t i = 0; i < nc; ++i) rank += (cd[i] < dj || (cd[i] == dj && i < j)) ? 1u : 0u;
      sid[rank] = cid[j];
      sd[rank] = dj;
    }
    __syncwarp();
    // walk ascending: keep c unless some kept e is closer to c than the owner is (operations.rs:199-217)
    uint32_t ns = 0;
    uint64_t kept_mask = 0;  // bit j: sorted candidate j was kept
    for (uint32_t j = 0; j < nc && ns < cap; ++j) {
      const float* vc = a.ix.arena + (size_t)sid[j] * a.ix.ds;
      const float dc = sd[j];
      bool closer = false;
      for (uint32_t base = 0; base < ns && !closer; base += 8) {
        const uint32_t e = min(base + quad, ns - 1);
        const float de = quad_l2sq(vc, a.ix.arena + (size_t)sel[e] * a.ix.ds, dim, p);
        closer = __any_sync(kFullMask, base + quad < ns && de < dc);
      }
      if (!closer) {
        if (lane == 0) sel[ns] = sid[j];
        kept_mask |= 1ull << j;
        ns += 1;
        __syncwarp();
      }
    }
    // back-fill from the candidates' (sorted) order until the list is full (operations.rs:219-230)
    for (uint32_t j = 0; j < nc && ns < cap; ++j) {
      if (!((kept_mask >> j) & 1ull)) {
        if (lane == 0) sel[ns] = sid[j];
        ns += 1;
      }
    }
    __syncwarp();
    for (uint32_t i = lane; i < cap; i += 32) row[i] = i < ns ? sel[i] : kInvalid;
    cnt = ns;
    __syncwarp();
  }
  if (lane == 0) *pcnt = (uint8_t)cnt;
}

}  // namespace turdb

// select_level + calculate_ml (operations.rs:76-83): floor(-ln(r) * (1 / ln(M))) as u8 (saturating; NaN -> 0), capped at 15
static uint8_t insert_select_level(double r, uint32_t m) {
  const double ml = 1.0 / std::log((double)m);
  const double lv = std::floor(-std::log(r) * ml);
  uint8_t level;
  if (std::isnan(lv) || lv <= 0.0) level = 0;
  else if (lv >= 255.0) level = 255;
  else level = (uint8_t)lv;
  return level < 15 ? level : 15;
}

extern "C" int32_t turdb_cuda_index_build(const turdb_cuda_build_params* bp, uint64_t n, const float* vectors,
                                          const uint64_t* row_ids, const double* random_values, int32_t device,
                                          turdb_cuda_index** out) {
  if (!bp || !out) return fail(TURDB_ERR_INVALID_ARGUMENT, "params/out is null");
  *out = nullptr;
  if (bp->dim == 0 || bp->dim > 65535) return fail(TURDB_ERR_INVALID_ARGUMENT, "dim %u out of range", bp->dim);
  if (bp->m < 2 || bp->m > 32) return fail(TURDB_ERR_UNSUPPORTED, "M = %u: supported range 2..32 (m0 = 2M <= 64)", bp->m);
  if (bp->ef_construction == 0 || bp->ef_construction > 2048) return fail(TURDB_ERR_INVALID_ARGUMENT, "ef_construction %u out of range", bp->ef_construction);
  if (n >= 0x7FFFFFFFull) return fail(TURDB_ERR_UNSUPPORTED, "n %llu exceeds 2^31-2 nodes per index", (unsigned long long)n);
  if (n && (!vectors || !row_ids || !random_values)) return fail(TURDB_ERR_INVALID_ARGUMENT, "null input array");
  const uint32_t max_batch = std::min(std::max(bp->max_batch, 1u), 16384u);
  const uint32_t dim = bp->dim;

  // levels, upper slots and their owners are known before the first insert
  std::vector<uint8_t> levels;
  std::vector<uint32_t> up_base, up_owner;
  try {
    levels.resize(n);
    up_base.resize(n);
    uint64_t slots = 0;
    for (uint64_t i = 0; i < n; ++i) {
      levels[i] = insert_select_level(random_values[i], bp->m);
      up_base[i] = levels[i] ? (uint32_t)slots : kInvalid;
      slots += levels[i];
    }
    up_owner.resize(slots);
    for (uint64_t i = 0; i < n; ++i)
      for (uint32_t l = 0; l < levels[i]; ++l) up_owner[up_base[i] + l] = (uint32_t)i;
  } catch (...) {
    return fail(TURDB_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  const uint64_t slots = up_owner.size();

  // an index with every array at its final size, adjacency empty: index_create uploads vectors / row ids / levels
  std::vector<uint32_t> empty_l0;
  std::vector<uint8_t> zero_cnt;
  try {
    empty_l0.assign((size_t)std::max<uint64_t>(n, 1) * kL0, kInvalid);
    zero_cnt.assign((size_t)std::max<uint64_t>(std::max(n, slots), 1), 0);
  } catch (...) {
    return fail(TURDB_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  std::vector<uint32_t> empty_up;
  try {
    empty_up.assign((size_t)std::max<uint64_t>(slots, 1) * kUp, kInvalid);
  } catch (...) {
    return fail(TURDB_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  turdb_cuda_graph g{};
  g.dim = dim;
  g.max_level = 0;
  g.n = n;
  g.entry = n ? 0u : TURDB_INVALID_NODE;
  g.vectors = vectors;
  g.row_ids = row_ids;
  g.levels = levels.data();
  g.l0_adj = empty_l0.data();
  g.l0_cnt = zero_cnt.data();
  g.up_base = up_base.data();
  g.up_adj = empty_up.data();
  g.up_cnt = zero_cnt.data();
  g.n_up_slots = slots;
  // index_create insists on entry level >= max_level; both are patched below once the build has run
  turdb_cuda_index* idx = nullptr;
  if (int32_t rc = turdb_cuda_index_create(&g, device, &idx); rc != TURDB_OK) return rc;
  if (n == 0) {
    *out = idx;
    return TURDB_OK;
  }
  DeviceGuard guard(device);
  cudaStream_t stream = nullptr;
  uint8_t *d_l0_cnt = nullptr, *d_up_cnt = nullptr, *d_ins_cnt = nullptr;
  uint32_t *d_up_owner = nullptr, *d_ins_sel = nullptr, *d_req_count = nullptr;
  unsigned long long *d_keys = nullptr, *d_keys_alt = nullptr;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  // at most m0 + levels * m requests per node; a step never holds more than max_batch nodes of level <= 15
  const uint32_t m = bp->m, m0 = 2 * bp->m;
  const size_t req_cap = (size_t)max_batch * (m0 + 15 * m);
  auto cleanup = [&]() {
    cudaFree(d_l0_cnt);
    cudaFree(d_up_cnt);
    cudaFree(d_ins_cnt);
    cudaFree(d_up_owner);
    cudaFree(d_ins_sel);
    cudaFree(d_req_count);
    cudaFree(d_keys);
    cudaFree(d_keys_alt);
    cudaFree(d_sort_tmp);
    if (stream) cudaStreamDestroy(stream);
  };
#define BUILD_TRY(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      cleanup();                                                                         \
      turdb_cuda_index_destroy(idx);                                                     \
      return fail(_e == cudaErrorMemoryAllocation ? TURDB_ERR_OUT_OF_MEMORY : TURDB_ERR_CUDA, "%s failed: %s", #expr, \
                  cudaGetErrorString(_e));                                               \
    }                                                                                    \
  } while (0)
  BUILD_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  BUILD_TRY(cudaMalloc(&d_l0_cnt, n));
  BUILD_TRY(cudaMemset(d_l0_cnt, 0, n));
  BUILD_TRY(cudaMalloc(&d_up_cnt, std::max<uint64_t>(slots, 1)));
  BUILD_TRY(cudaMemset(d_up_cnt, 0, std::max<uint64_t>(slots, 1)));
  BUILD_TRY(cudaMalloc(&d_up_owner, std::max<uint64_t>(slots, 1) * 4));
  if (slots) BUILD_TRY(cudaMemcpy(d_up_owner, up_owner.data(), slots * 4, cudaMemcpyHostToDevice));
  BUILD_TRY(cudaMalloc(&d_ins_sel, (size_t)max_batch * kInsLevels * kInsSelMax * 4));
  BUILD_TRY(cudaMalloc(&d_ins_cnt, (size_t)max_batch * kInsLevels));
  BUILD_TRY(cudaMalloc(&d_req_count, 4));
  BUILD_TRY(cudaMalloc(&d_keys, req_cap * 8));
  BUILD_TRY(cudaMalloc(&d_keys_alt, req_cap * 8));
  BUILD_TRY(cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp_bytes, d_keys, d_keys_alt, (int)req_cap, 0, 64, stream));
  BUILD_TRY(cudaMalloc(&d_sort_tmp, std::max<size_t>(sort_tmp_bytes, 16)));

  // node 0: the first insert only sets the entry point (mod.rs:1020-1023)
  uint32_t entry = 0, max_level = levels[0];
  const uint32_t ef = bp->ef_construction;
  uint64_t s = 1;
  while (s < n) {
    // all nodes of a step search the graph of nodes [0, s); steps grow with the graph (at most 1/8 of it)
    const uint32_t B = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(max_batch, std::max<uint64_t>(1, s / 8)), n - s);
    idx->ix.entry = entry;
    idx->ix.max_level = max_level;
    int32_t rc = insert_search_step(idx, (uint32_t)s, B, ef, m, m0, d_ins_sel, d_ins_cnt, stream);
    if (rc != TURDB_OK) {
      cleanup();
      turdb_cuda_index_destroy(idx);
      return rc;
    }
    InsertLinkArgs la{};
    la.first = (uint32_t)s;
    la.count = B;
    la.n_total = n;
    la.levels = idx->d_levels;
    la.up_base = idx->d_up_base;
    la.l0_adj = idx->d_l0_adj;
    la.l0_cnt = d_l0_cnt;
    la.up_adj = idx->d_up_adj;
    la.up_cnt = d_up_cnt;
    la.ins_sel = d_ins_sel;
    la.ins_cnt = d_ins_cnt;
    la.req_keys = d_keys;
    la.req_count = d_req_count;
    la.req_cap = (uint32_t)req_cap;
    size_t bound = 0;  // requests this step can produce
    for (uint64_t i = s; i < s + B; ++i) bound += m0 + (size_t)levels[i] * m;
    BUILD_TRY(cudaMemsetAsync(d_req_count, 0, 4, stream));
    BUILD_TRY(cudaMemsetAsync(d_keys, 0xFF, bound * 8, stream));
    insert_link_own_kernel<<<(B * 32 + 127) / 128, 128, 0, stream>>>(la);
    const unsigned long long* sorted = d_keys;
    if (bound > 1) {
      size_t tmp = sort_tmp_bytes;
      BUILD_TRY(cub::DeviceRadixSort::SortKeys(d_sort_tmp, tmp, d_keys, d_keys_alt, (int)bound, 0, 64, stream));
      sorted = d_keys_alt;
    }
    InsertBacklinkArgs ba{};
    ba.ix = idx->ix;
    ba.first = (uint32_t)s;
    ba.n_total = n;
    ba.req_keys = sorted;
    ba.req_count = d_req_count;
    ba.req_cap = (uint32_t)bound;
    ba.up_owner = d_up_owner;
    ba.l0_adj = idx->d_l0_adj;
    ba.l0_cnt = d_l0_cnt;
    ba.up_adj = idx->d_up_adj;
    ba.up_cnt = d_up_cnt;
    ba.intent = bp->mode != 0;
    insert_backlink_kernel<<<(unsigned)((bound * 32 + 127) / 128), 128, 0, stream>>>(ba);
    BUILD_TRY(cudaGetLastError());
    for (uint64_t i = s; i < s + B; ++i)  // mod.rs:1079-1081, in id order
      if (levels[i] > max_level) {
        max_level = levels[i];
        entry = (uint32_t)i;
      }
    s += B;
  }
  BUILD_TRY(cudaStreamSynchronize(stream));
  idx->ix.entry = entry;
  idx->ix.max_level = max_level;
  cleanup();
#undef BUILD_TRY
  *out = idx;
  return TURDB_OK;
}

// The graph of an index as flat host arrays (the turdb_cuda_graph layout): what the device holds after an upload or a
// build.  Counts are the number of valid ids of a row.  Every pointer is nullable.
extern "C" int32_t turdb_cuda_index_export_graph(turdb_cuda_index* idx, float* vectors, uint64_t* row_ids, uint8_t* levels,
                                                 uint32_t* l0_adj, uint8_t* l0_cnt, uint32_t* up_base, uint32_t* up_adj,
                                                 uint8_t* up_cnt, uint32_t* entry, uint32_t* max_level, uint64_t* n_up_slots) {
  if (!idx) return fail(TURDB_ERR_INVALID_ARGUMENT, "idx is null");
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", idx->device);
  const uint64_t n = idx->ix.n;
  const uint32_t dim = idx->ix.dim, ds = idx->ix.ds;
  if (entry) *entry = idx->ix.entry;
  if (max_level) *max_level = idx->ix.max_level;
  if (n_up_slots) *n_up_slots = idx->n_up_slots;
  if (n == 0) return TURDB_OK;
  CUDA_TRY(cudaDeviceSynchronize());
  if (vectors) CUDA_TRY(cudaMemcpy2D(vectors, (size_t)dim * 4, idx->d_arena, (size_t)ds * 4, (size_t)dim * 4, n, cudaMemcpyDeviceToHost));
  if (row_ids) CUDA_TRY(cudaMemcpy(row_ids, idx->d_row_ids, n * 8, cudaMemcpyDeviceToHost));
  if (levels) CUDA_TRY(cudaMemcpy(levels, idx->d_levels, n, cudaMemcpyDeviceToHost));
  if (up_base) CUDA_TRY(cudaMemcpy(up_base, idx->d_up_base, n * 4, cudaMemcpyDeviceToHost));
  try {
    if (l0_adj || l0_cnt) {
      std::vector<uint32_t> tmp;
      uint32_t* dst = l0_adj;
      if (!dst) {
        tmp.resize(n * kL0);
        dst = tmp.data();
      }
      CUDA_TRY(cudaMemcpy(dst, idx->d_l0_adj, n * kL0 * 4, cudaMemcpyDeviceToHost));
      if (l0_cnt)
        for (uint64_t i = 0; i < n; ++i) {
          uint32_t c = 0;
          while (c < kL0 && dst[i * kL0 + c] != kInvalid) ++c;
          l0_cnt[i] = (uint8_t)c;
        }
    }
    const uint64_t slots = idx->n_up_slots;
    if (slots && (up_adj || up_cnt)) {
      std::vector<uint32_t> tmp;
      uint32_t* dst = up_adj;
      if (!dst) {
        tmp.resize(slots * kUp);
        dst = tmp.data();
      }
      CUDA_TRY(cudaMemcpy(dst, idx->d_up_adj, slots * kUp * 4, cudaMemcpyDeviceToHost));
      if (up_cnt)
        for (uint64_t i = 0; i < slots; ++i) {
          uint32_t c = 0;
          while (c < kUp && dst[i * kUp + c] != kInvalid) ++c;
          up_cnt[i] = (uint8_t)c;
        }
    }
  } catch (...) {
    return fail(TURDB_ERR_OUT_OF_MEMORY, "host allocation failed");
  }
  return TURDB_OK;
}
