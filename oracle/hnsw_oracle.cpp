/*
 * hnsw_oracle.cpp — CPU oracle for the TurDB HNSW vector-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see hnsw_oracle.h).  PARITY STATUS: "parity unpinned" for HNSW
 * search / insert / distance: the reference has no test pinning them and cannot be compiled here
 * (no cargo/rustc).  This file is a line-against-line restatement of the cited ranges over dense
 * u32 node ids (id = insertion order = allocate_node order, src/hnsw/mod.rs:883-904).
 *
 * Deliberate divergence: the 13-bit slot-offset truncation of src/hnsw/storage.rs:338-344 is
 * storage corruption, not search semantics, and is NOT emulated.
 *
 * Build: see oracle/Makefile (g++ -O3 -mavx2 -mfma, no other dependency).
 */
#include "hnsw_oracle.h"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------------------------
// Distances — src/hnsw/distance.rs
// ----------------------------------------------------------------------------------------------

// horizontal_sum_avx2, distance.rs:150-161: (lo128 + hi128) -> movehl add -> lane-1 add.
inline float hsum_avx2(__m256 v) {
  __m128 hi = _mm256_extractf128_ps(v, 1);
  __m128 lo = _mm256_castps256_ps128(v);
  __m128 sum128 = _mm_add_ps(lo, hi);
  __m128 hi64 = _mm_movehl_ps(sum128, sum128);
  __m128 sum64 = _mm_add_ps(sum128, hi64);
  __m128 hi32 = _mm_shuffle_ps(sum64, sum64, 1);
  __m128 sum32 = _mm_add_ss(sum64, hi32);
  return _mm_cvtss_f32(sum32);
}

// euclidean_squared_avx2, distance.rs:105-129.  The scalar tail is `result += diff * diff`, which
// rustc does not contract into an FMA; `volatile` pins the separately rounded product here.
float l2sq_avx2(const float* a, const float* b, size_t n) {
  size_t i = 0;
  __m256 sum = _mm256_setzero_ps();
  while (i + 8 <= n) {
    __m256 va = _mm256_loadu_ps(a + i);
    __m256 vb = _mm256_loadu_ps(b + i);
    __m256 diff = _mm256_sub_ps(va, vb);
    sum = _mm256_fmadd_ps(diff, diff, sum);
    i += 8;
  }
  float result = hsum_avx2(sum);
  while (i < n) {
    float diff = a[i] - b[i];
    volatile float sq = diff * diff;
    result += sq;
    i += 1;
  }
  return result;
}

// dot_product_avx2, distance.rs:210-232
float dot_avx2(const float* a, const float* b, size_t n) {
  size_t i = 0;
  __m256 sum = _mm256_setzero_ps();
  while (i + 8 <= n) {
    __m256 va = _mm256_loadu_ps(a + i);
    __m256 vb = _mm256_loadu_ps(b + i);
    sum = _mm256_fmadd_ps(va, vb, sum);
    i += 8;
  }
  float result = hsum_avx2(sum);
  while (i < n) {
    volatile float p = a[i] * b[i];
    result += p;
    i += 1;
  }
  return result;
}

// cosine_avx2, distance.rs:250-285
float cosine_avx2(const float* a, const float* b, size_t n) {
  size_t i = 0;
  __m256 dot_sum = _mm256_setzero_ps();
  __m256 na_sum = _mm256_setzero_ps();
  __m256 nb_sum = _mm256_setzero_ps();
  while (i + 8 <= n) {
    __m256 va = _mm256_loadu_ps(a + i);
    __m256 vb = _mm256_loadu_ps(b + i);
    dot_sum = _mm256_fmadd_ps(va, vb, dot_sum);
    na_sum = _mm256_fmadd_ps(va, va, na_sum);
    nb_sum = _mm256_fmadd_ps(vb, vb, nb_sum);
    i += 8;
  }
  float dot = hsum_avx2(dot_sum);
  float norm_a = hsum_avx2(na_sum);
  float norm_b = hsum_avx2(nb_sum);
  while (i < n) {
    volatile float p = a[i] * b[i];
    dot += p;
    volatile float pa = a[i] * a[i];
    norm_a += pa;
    volatile float pb = b[i] * b[i];
    norm_b += pb;
    i += 1;
  }
  float norm_product = std::sqrt(norm_a * norm_b);
  if (norm_product == 0.0f) return 1.0f;
  return 1.0f - (dot / norm_product);
}

// Scalar emulation of the 8-lane order above (what the CUDA kernels mirror lane for lane):
// lane j accumulates elements i == j (mod 8) with one fused multiply-add per element, then the
// fixed tree (s0+s4, s1+s5, s2+s6, s3+s7) -> (t0+t2, t1+t3) -> u0+u1, then the unfused tail.
struct Emu8 {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float hsum() const {
    float t0 = acc[0] + acc[4], t1 = acc[1] + acc[5], t2 = acc[2] + acc[6], t3 = acc[3] + acc[7];
    float u0 = t0 + t2, u1 = t1 + t3;
    return u0 + u1;
  }
};

float l2sq_emulated(const float* a, const float* b, size_t n) {
  Emu8 e;
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; ++j) {
      float d = a[i + j] - b[i + j];
      e.acc[j] = std::fmaf(d, d, e.acc[j]);
    }
  float r = e.hsum();
  for (; i < n; ++i) {
    float d = a[i] - b[i];
    volatile float sq = d * d;
    r += sq;
  }
  return r;
}

float dot_emulated(const float* a, const float* b, size_t n) {
  Emu8 e;
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; ++j) e.acc[j] = std::fmaf(a[i + j], b[i + j], e.acc[j]);
  float r = e.hsum();
  for (; i < n; ++i) {
    volatile float p = a[i] * b[i];
    r += p;
  }
  return r;
}

float cosine_emulated(const float* a, const float* b, size_t n) {
  float dot = dot_emulated(a, b, n);
  float na = dot_emulated(a, a, n);
  float nb = dot_emulated(b, b, n);
  float np = std::sqrt(na * nb);
  if (np == 0.0f) return 1.0f;
  return 1.0f - (dot / np);
}

// scalar variants, distance.rs:55-97 (used when AVX2/FMA are not detected)
float l2sq_scalar(const float* a, const float* b, size_t n) {
  float sum = 0.0f;
  for (size_t i = 0; i < n; ++i) {
    float d = a[i] - b[i];
    volatile float sq = d * d;
    sum += sq;
  }
  return sum;
}
float dot_scalar(const float* a, const float* b, size_t n) {
  float sum = 0.0f;
  for (size_t i = 0; i < n; ++i) {
    volatile float p = a[i] * b[i];
    sum += p;
  }
  return sum;
}
float cosine_scalar(const float* a, const float* b, size_t n) {
  float dot = 0, na = 0, nb = 0;
  for (size_t i = 0; i < n; ++i) {
    volatile float p = a[i] * b[i];
    dot += p;
    volatile float pa = a[i] * a[i];
    na += pa;
    volatile float pb = b[i] * b[i];
    nb += pb;
  }
  float np = std::sqrt(na * nb);
  if (np == 0.0f) return 1.0f;
  return 1.0f - (dot / np);
}

// select_squared_distance_fn, distance.rs:438-444 (L2 -> L2^2, Cosine -> 1-cos, IP -> -dot)
inline float metric_distance(int metric, const float* a, const float* b, size_t n) {
  switch (metric) {
    case TDO_COSINE: return cosine_avx2(a, b, n);
    case TDO_IP: return -dot_avx2(a, b, n);
    default: return l2sq_avx2(a, b, n);
  }
}

// ----------------------------------------------------------------------------------------------
// Rust std::collections::BinaryHeap (max-heap) — push/pop order among equal keys follows the
// std algorithm (sift_up with `<=` stop; pop = swap-remove root, sift_down_to_bottom, sift_up).
// Call sites: src/hnsw/search.rs:194-195,228,232-235,247,329.
// ----------------------------------------------------------------------------------------------
struct Cand {
  uint32_t id;
  float d;
};

// `le(a, b)`: a <= b in the heap's Ord.
template <class Le>
struct RustHeap {
  std::vector<Cand> data;
  Le le;
  void clear() { data.clear(); }
  size_t size() const { return data.size(); }
  bool empty() const { return data.empty(); }
  const Cand& peek() const { return data[0]; }
  void sift_up(size_t start, size_t pos) {
    Cand elt = data[pos];
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (le(elt, data[parent])) break;
      data[pos] = data[parent];
      pos = parent;
    }
    data[pos] = elt;
  }
  void push(Cand c) {
    size_t old_len = data.size();
    data.push_back(c);
    sift_up(0, old_len);
  }
  void sift_down_to_bottom(size_t pos) {
    size_t end = data.size();
    size_t start = pos;
    Cand elt = data[pos];
    size_t child = 2 * pos + 1;
    size_t lim = end >= 2 ? end - 2 : 0;
    while (child <= lim && end >= 2) {
      if (le(data[child], data[child + 1])) child += 1;
      data[pos] = data[child];
      pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1) {
      data[pos] = data[child];
      pos = child;
    }
    data[pos] = elt;
    sift_up(start, pos);
  }
  Cand pop() {
    Cand item = data.back();
    data.pop_back();
    if (!data.empty()) {
      std::swap(item, data[0]);
      sift_down_to_bottom(0);
    }
    return item;
  }
};

// Candidate: Ord::cmp = other.distance.partial_cmp(self.distance).unwrap_or(Equal)  (search.rs:108-115)
// a <= b  <=>  cmp(a,b) != Greater  <=>  !(b.d < a.d)... cmp(a,b)=Greater iff other.d > self.d iff b.d > a.d
struct MinLe {
  bool operator()(const Cand& a, const Cand& b) const { return !(b.d > a.d); }
};
// ReverseCandidate: cmp = self.d.partial_cmp(other.d).unwrap_or(Equal)  (search.rs:134-141)
struct MaxLe {
  bool operator()(const Cand& a, const Cand& b) const { return !(a.d > b.d); }
};

// VisitedSet, search.rs:143-191 (generation stamps; dense index = node id)
struct Visited {
  uint32_t gen = 1;
  std::vector<uint32_t> stamp;
  void clear() {
    gen += 1;
    if (gen == 0) {
      std::fill(stamp.begin(), stamp.end(), 0u);
      gen = 1;
    }
  }
  bool insert(size_t idx) {
    if (idx >= stamp.size()) {
      size_t ns = 1;
      while (ns < idx + 1) ns <<= 1;
      stamp.resize(ns, 0u);
    }
    if (stamp[idx] == gen) return false;
    stamp[idx] = gen;
    return true;
  }
};

// HnswSearchContext, search.rs:193-257
struct Ctx {
  RustHeap<MinLe> candidates;
  RustHeap<MaxLe> results;
  Visited visited;
  std::vector<Cand> output;
  size_t ef = 0;
  void reset() {
    candidates.clear();
    results.clear();
    visited.clear();
    output.clear();
  }
  void add_result(Cand c) {  // :231-236
    results.push(c);
    if (results.size() > ef) results.pop();
  }
  float worst() const {  // :238-243
    return results.empty() ? std::numeric_limits<float>::infinity() : results.peek().d;
  }
  void finalize(size_t k) {  // :245-252
    output.clear();
    while (!results.empty()) output.push_back(results.pop());
    std::reverse(output.begin(), output.end());
    if (output.size() > k) output.resize(k);
  }
};

}  // namespace

// ----------------------------------------------------------------------------------------------
// Graph
// ----------------------------------------------------------------------------------------------
struct tdo_graph {
  uint16_t dim = 0, m = 16, efc = 100;
  int mode = TDO_BUILD_INTENT;
  std::vector<float> vec;
  std::vector<uint64_t> row_ids;
  std::vector<uint8_t> levels;
  std::vector<uint32_t> l0_adj;
  std::vector<uint8_t> l0_cnt;
  std::vector<uint32_t> up_base;
  std::vector<uint32_t> up_adj;
  std::vector<uint8_t> up_cnt;
  uint32_t entry = TDO_INVALID;
  uint8_t max_level = 0;
  uint64_t build_dist_evals = 0;
  Ctx build_ctx;

  size_t n() const { return levels.size(); }
  const float* v(uint32_t id) const { return vec.data() + (size_t)id * dim; }

  // HnswNode::neighbors_at_level, mod.rs:283-291
  const uint32_t* nbrs(uint32_t id, uint8_t level, uint32_t* cnt) const {
    if (level == 0) {
      *cnt = l0_cnt[id];
      return l0_adj.data() + (size_t)id * TDO_MAX_L0_NEIGHBORS;
    }
    if (level <= levels[id]) {
      size_t slot = (size_t)up_base[id] + (level - 1);
      *cnt = up_cnt[slot];
      return up_adj.data() + slot * TDO_MAX_LEVEL_NEIGHBORS;
    }
    *cnt = 0;
    return nullptr;
  }
  uint32_t cap(uint8_t level) const {
    return level == 0 ? TDO_MAX_L0_NEIGHBORS : TDO_MAX_LEVEL_NEIGHBORS;
  }
  uint32_t* list(uint32_t id, uint8_t level, uint8_t** cnt) {
    if (level == 0) {
      *cnt = &l0_cnt[id];
      return l0_adj.data() + (size_t)id * TDO_MAX_L0_NEIGHBORS;
    }
    if (level <= levels[id]) {
      size_t slot = (size_t)up_base[id] + (level - 1);
      *cnt = &up_cnt[slot];
      return up_adj.data() + slot * TDO_MAX_LEVEL_NEIGHBORS;
    }
    *cnt = nullptr;
    return nullptr;
  }
};

namespace {

// greedy_search, search.rs:283-309 (+ greedy_search_step :259-281): strict `<`, first wins, <=1000 iters
template <class Dist>
void greedy(const tdo_graph& g, uint8_t level, uint32_t& cur, float& cur_d, Dist&& dist,
            tdo_stats* st) {
  for (int it = 0; it < 1000; ++it) {
    uint32_t cnt;
    const uint32_t* nb = g.nbrs(cur, level, &cnt);
    if (st) st->n_upper_hops += 1;
    uint32_t best = cur;
    float best_d = cur_d;
    for (uint32_t i = 0; i < cnt; ++i) {
      float d = dist(nb[i]);
      if (st) {
        st->n_dist += 1;
        st->n_dist_upper += 1;
      }
      if (d < best_d) {
        best_d = d;
        best = nb[i];
      }
    }
    if (best == cur) break;
    cur = best;
    cur_d = best_d;
  }
}

// beam_search, search.rs:311-350
template <class Dist>
void beam(const tdo_graph& g, Ctx& ctx, uint8_t level, Cand entry, Dist&& dist, tdo_stats* st) {
  ctx.reset();
  if (ctx.visited.insert(entry.id)) {
    ctx.candidates.push(entry);
    ctx.add_result(entry);
  }
  while (!ctx.candidates.empty()) {
    Cand cur = ctx.candidates.pop();
    if (cur.d > ctx.worst()) break;
    uint32_t cnt;
    const uint32_t* nb = g.nbrs(cur.id, level, &cnt);
    if (st) st->n_expanded += 1;
    for (uint32_t i = 0; i < cnt; ++i) {
      uint32_t n = nb[i];
      if (!ctx.visited.insert(n)) continue;
      float d = dist(n);
      if (st) st->n_dist += 1;
      if (d < ctx.worst() || ctx.results.size() < ctx.ef) {
        ctx.candidates.push(Cand{n, d});
        ctx.add_result(Cand{n, d});
      }
    }
  }
}

// beam_search_filtered, search.rs:352-398
template <class Dist, class Vis>
void beam_filtered(const tdo_graph& g, Ctx& ctx, Cand entry, Dist&& dist, Vis&& visible,
                   tdo_stats* st) {
  ctx.reset();
  if (ctx.visited.insert(entry.id)) {
    ctx.candidates.push(entry);
    if (visible(entry.id)) ctx.add_result(entry);
  }
  while (!ctx.candidates.empty()) {
    Cand cur = ctx.candidates.pop();
    if (cur.d > ctx.worst()) break;
    uint32_t cnt;
    const uint32_t* nb = g.nbrs(cur.id, 0, &cnt);
    if (st) st->n_expanded += 1;
    for (uint32_t i = 0; i < cnt; ++i) {
      uint32_t n = nb[i];
      if (!ctx.visited.insert(n)) continue;
      float d = dist(n);
      if (st) st->n_dist += 1;
      ctx.candidates.push(Cand{n, d});
      if (visible(n) && (d < ctx.worst() || ctx.results.size() < ctx.ef)) ctx.add_result(Cand{n, d});
    }
  }
}

// select_neighbors_heuristic, operations.rs:181-233.  `cands` arrive in the caller's order; they
// are stably sorted ascending for the walk, and the back-fill walks the caller's order.
std::vector<uint32_t> select_heuristic(const tdo_graph& g, const std::vector<Cand>& cands,
                                       size_t max_neighbors, uint64_t* evals) {
  std::vector<uint32_t> selected;
  if (cands.empty()) return selected;
  std::vector<const Cand*> remaining;
  for (auto& c : cands) remaining.push_back(&c);
  std::stable_sort(remaining.begin(), remaining.end(),
                   [](const Cand* a, const Cand* b) { return a->d < b->d; });
  for (const Cand* c : remaining) {
    if (selected.size() >= max_neighbors) break;
    bool closer = false;
    for (uint32_t e : selected) {
      float de = l2sq_avx2(g.v(c->id), g.v(e), g.dim);
      *evals += 1;
      if (de < c->d) {
        closer = true;
        break;
      }
    }
    if (!closer) selected.push_back(c->id);
  }
  if (selected.size() < max_neighbors) {
    for (auto& c : cands) {
      if (selected.size() >= max_neighbors) break;
      if (std::find(selected.begin(), selected.end(), c.id) == selected.end())
        selected.push_back(c.id);
    }
  }
  return selected;
}

// add_neighbor_at_level on the neighbour side, mod.rs:293-301 (+ :275-280): append if room.
// verbatim: silently dropped when full.  intent: re-select with select_neighbors_heuristic over
// (current list + new id) sorted by distance to the owner.
void add_backlink(tdo_graph& g, uint32_t owner, uint8_t level, uint32_t id) {
  uint8_t* cnt;
  uint32_t* l = g.list(owner, level, &cnt);
  if (!l) return;  // neighbour lacks this level: dropped (mod.rs:296)
  uint32_t cap = g.cap(level);
  if (*cnt < cap) {
    l[*cnt] = id;
    *cnt += 1;
    return;
  }
  if (g.mode == TDO_BUILD_VERBATIM) return;
  std::vector<Cand> cands;
  cands.reserve(cap + 1);
  for (uint32_t i = 0; i < *cnt; ++i) {
    cands.push_back(Cand{l[i], l2sq_avx2(g.v(owner), g.v(l[i]), g.dim)});
    g.build_dist_evals += 1;
  }
  cands.push_back(Cand{id, l2sq_avx2(g.v(owner), g.v(id), g.dim)});
  g.build_dist_evals += 1;
  std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.d < b.d; });
  std::vector<uint32_t> sel = select_heuristic(g, cands, cap, &g.build_dist_evals);
  for (size_t i = 0; i < sel.size(); ++i) l[i] = sel[i];
  for (size_t i = sel.size(); i < cap; ++i) l[i] = TDO_INVALID;
  *cnt = (uint8_t)sel.size();
}

}  // namespace

extern "C" {

float tdo_distance(int metric, const float* a, const float* b, uint32_t dim, int impl) {
  if (impl == TDO_DIST_AVX2) return metric_distance(metric, a, b, dim);
  if (impl == TDO_DIST_AVX2_EMULATED) {
    switch (metric) {
      case TDO_COSINE: return cosine_emulated(a, b, dim);
      case TDO_IP: return -dot_emulated(a, b, dim);
      default: return l2sq_emulated(a, b, dim);
    }
  }
  switch (metric) {
    case TDO_COSINE: return cosine_scalar(a, b, dim);
    case TDO_IP: return -dot_scalar(a, b, dim);
    default: return l2sq_scalar(a, b, dim);
  }
}

// select_level + calculate_ml, operations.rs:76-83.  Rust `as u8` saturates (NaN -> 0).
uint8_t tdo_select_level(double random_value, uint16_t m) {
  double ml = 1.0 / std::log((double)m);
  double lv = std::floor(-std::log(random_value) * ml);
  uint8_t level;
  if (std::isnan(lv) || lv <= 0.0) level = 0;
  else if (lv >= 255.0) level = 255;
  else level = (uint8_t)lv;
  return level < 15 ? level : 15;
}

tdo_graph* tdo_graph_new(uint16_t dim, uint16_t m, uint16_t ef_construction, int build_mode) {
  tdo_graph* g = new tdo_graph();
  g->dim = dim;
  g->m = m;
  g->efc = ef_construction;
  g->mode = build_mode;
  return g;
}

void tdo_graph_free(tdo_graph* g) { delete g; }

// insert_with_callback, mod.rs:999-1084 (+ operations.rs:111-171)
int tdo_graph_insert(tdo_graph* gp, uint64_t row_id, const float* vector, double random_value) {
  tdo_graph& g = *gp;
  const uint8_t target_level = tdo_select_level(random_value, g.m);

  // HnswNode::new + allocate_node (mod.rs:1016-1018)
  const uint32_t id = (uint32_t)g.n();
  g.vec.insert(g.vec.end(), vector, vector + g.dim);
  g.row_ids.push_back(row_id);
  g.levels.push_back(target_level);
  g.l0_adj.insert(g.l0_adj.end(), TDO_MAX_L0_NEIGHBORS, TDO_INVALID);
  g.l0_cnt.push_back(0);
  if (target_level > 0) {
    g.up_base.push_back((uint32_t)g.up_cnt.size());
    g.up_adj.insert(g.up_adj.end(), (size_t)target_level * TDO_MAX_LEVEL_NEIGHBORS, TDO_INVALID);
    g.up_cnt.insert(g.up_cnt.end(), target_level, 0);
  } else {
    g.up_base.push_back(TDO_INVALID);
  }

  if (g.entry == TDO_INVALID) {  // mod.rs:1020-1023, set_entry_point :700-705
    g.entry = id;
    if (target_level > g.max_level) g.max_level = target_level;
    return 0;
  }

  const float* q = g.v(id);
  auto dist = [&](uint32_t n) {
    g.build_dist_evals += 1;
    return l2sq_avx2(q, g.v(n), g.dim);  // always L2^2, mod.rs:1031,1046
  };

  uint32_t e = g.entry;
  float de = dist(e);

  // insert_descent_phase, operations.rs:111-133
  for (int level = g.max_level; level > (int)target_level; --level)
    greedy(g, (uint8_t)level, e, de, dist, nullptr);

  // insert_connection_phase, operations.rs:135-171.  current_entry is never refined: the second
  // finalize_results(1) drains an already-empty heap (operations.rs:164-169, search.rs:245-252).
  Ctx& ctx = g.build_ctx;
  ctx.ef = g.efc;
  const size_t m = g.m, m0 = (size_t)g.m * 2;  // HnswIndex::new, mod.rs:628-641
  std::vector<std::pair<uint8_t, std::vector<uint32_t>>> to_add;
  for (int level = target_level; level >= 0; --level) {
    beam(g, ctx, (uint8_t)level, Cand{e, de}, dist, nullptr);
    ctx.finalize(level == 0 ? m0 : m);
    std::vector<uint32_t> selected;
    for (auto& c : ctx.output) selected.push_back(c.id);
    to_add.emplace_back((uint8_t)level, std::move(selected));
  }

  // mod.rs:1067-1077
  for (auto& [level, neighbors] : to_add) {
    for (uint32_t nb : neighbors) {
      uint8_t* cnt;
      uint32_t* l = g.list(id, level, &cnt);
      if (l && *cnt < g.cap(level)) {
        l[*cnt] = nb;
        *cnt += 1;
      }
      add_backlink(g, nb, level, id);
    }
  }

  if (target_level > g.max_level) {  // mod.rs:1079-1081
    g.entry = id;
    g.max_level = target_level;
  }
  return 0;
}

int tdo_graph_insert_batch(tdo_graph* g, uint64_t n, const uint64_t* row_ids, const float* vecs,
                           const double* random_values) {
  for (uint64_t i = 0; i < n; ++i) {
    int rc = tdo_graph_insert(g, row_ids[i], vecs + i * g->dim, random_values[i]);
    if (rc) return rc;
  }
  return 0;
}

tdo_graph* tdo_graph_from_arrays(uint16_t dim, uint64_t n, const float* vectors,
                                 const uint64_t* row_ids, const uint8_t* levels,
                                 const uint32_t* l0_adj, const uint8_t* l0_cnt,
                                 const uint32_t* up_base, const uint32_t* up_adj,
                                 const uint8_t* up_cnt, uint64_t n_up_slots, uint32_t entry,
                                 uint8_t max_level) {
  tdo_graph* g = new tdo_graph();
  g->dim = dim;
  g->vec.assign(vectors, vectors + n * dim);
  g->row_ids.assign(row_ids, row_ids + n);
  g->levels.assign(levels, levels + n);
  g->l0_adj.assign(l0_adj, l0_adj + n * TDO_MAX_L0_NEIGHBORS);
  g->l0_cnt.assign(l0_cnt, l0_cnt + n);
  g->up_base.assign(up_base, up_base + n);
  if (n_up_slots) {
    g->up_adj.assign(up_adj, up_adj + n_up_slots * TDO_MAX_LEVEL_NEIGHBORS);
    g->up_cnt.assign(up_cnt, up_cnt + n_up_slots);
  }
  g->entry = n ? entry : TDO_INVALID;
  g->max_level = max_level;
  return g;
}

uint64_t tdo_graph_n(const tdo_graph* g) { return g->n(); }
uint64_t tdo_graph_n_up_slots(const tdo_graph* g) { return g->up_cnt.size(); }
uint32_t tdo_graph_entry(const tdo_graph* g) { return g->entry; }
uint8_t tdo_graph_max_level(const tdo_graph* g) { return g->max_level; }
uint64_t tdo_graph_build_dist_evals(const tdo_graph* g) { return g->build_dist_evals; }

int tdo_graph_export(const tdo_graph* g, float* vectors, uint64_t* row_ids, uint8_t* levels,
                     uint32_t* l0_adj, uint8_t* l0_cnt, uint32_t* up_base, uint32_t* up_adj,
                     uint8_t* up_cnt) {
  auto cp = [](auto* dst, const auto& src) {
    if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(src[0]));
  };
  cp(vectors, g->vec);
  cp(row_ids, g->row_ids);
  cp(levels, g->levels);
  cp(l0_adj, g->l0_adj);
  cp(l0_cnt, g->l0_cnt);
  cp(up_base, g->up_base);
  cp(up_adj, g->up_adj);
  cp(up_cnt, g->up_cnt);
  return 0;
}

// PersistentHnswIndex::search / search_filtered, mod.rs:1092-1273, extended with the metric
// selected per select_squared_distance_fn (distance.rs:438-444).
int tdo_search_batch(const tdo_graph* gp, const float* queries, uint32_t query_dim, uint32_t nq,
                     uint32_t k, uint32_t ef, int metric, const uint64_t* visible,
                     uint64_t* out_row_ids, uint32_t* out_node_ids, float* out_dist,
                     uint32_t* out_counts, tdo_stats* out_stats, int n_threads) {
  const tdo_graph& g = *gp;
  if (query_dim != g.dim) return 2;  // mod.rs:1099-1104
  if (n_threads < 1) n_threads = 1;
  if ((uint32_t)n_threads > nq) n_threads = nq ? (int)nq : 1;
  std::atomic<uint32_t> next{0};
  auto worker = [&]() {
    Ctx ctx;
    ctx.ef = ef;
    ctx.visited.stamp.assign(g.n(), 0u);
    for (;;) {
      uint32_t q0 = next.fetch_add(16);
      if (q0 >= nq) break;
      uint32_t q1 = std::min(nq, q0 + 16);
      for (uint32_t qi = q0; qi < q1; ++qi) {
        const float* q = queries + (size_t)qi * g.dim;
        tdo_stats st{0, 0, 0, 0};
        uint32_t cnt = 0;
        if (g.entry != TDO_INVALID) {  // mod.rs:1106-1109
          // compute_distance, mod.rs:1111-1121: get_vector(row_id) -> None (and an unreadable node) => f32::INFINITY,
          // whatever the metric.  In the flattened arrays an absent vector is a row whose elements are +inf
          // (the uploader's convention, turdb_cuda_hnsw_file_upload).
          auto dist = [&](uint32_t n) {
            const float* v = g.v(n);
            if (g.dim && v[0] == std::numeric_limits<float>::infinity()) return std::numeric_limits<float>::infinity();
            return metric_distance(metric, q, v, g.dim);
          };
          uint32_t cur = g.entry;
          float cur_d = dist(cur);
          st.n_dist += 1;
          st.n_dist_upper += 1;
          for (int level = g.max_level; level >= 1; --level)  // mod.rs:1134-1145
            greedy(g, (uint8_t)level, cur, cur_d, dist, &st);
          if (visible) {
            auto vis = [&](uint32_t n) { return (visible[n >> 6] >> (n & 63)) & 1ull; };
            beam_filtered(g, ctx, Cand{cur, cur_d}, dist, vis, &st);
            ctx.finalize(k);
            for (auto& c : ctx.output) {  // mod.rs:1257-1269
              if (!vis(c.id)) continue;
              out_row_ids[(size_t)qi * k + cnt] = g.row_ids[c.id];
              out_node_ids[(size_t)qi * k + cnt] = c.id;
              out_dist[(size_t)qi * k + cnt] = c.d;
              cnt += 1;
            }
          } else {
            beam(g, ctx, 0, Cand{cur, cur_d}, dist, &st);
            ctx.finalize(k);
            for (auto& c : ctx.output) {  // mod.rs:1159-1171
              out_row_ids[(size_t)qi * k + cnt] = g.row_ids[c.id];
              out_node_ids[(size_t)qi * k + cnt] = c.id;
              out_dist[(size_t)qi * k + cnt] = c.d;
              cnt += 1;
            }
          }
        }
        for (uint32_t j = cnt; j < k; ++j) {
          out_row_ids[(size_t)qi * k + j] = ~0ull;
          out_node_ids[(size_t)qi * k + j] = TDO_INVALID;
          out_dist[(size_t)qi * k + j] = std::numeric_limits<float>::infinity();
        }
        out_counts[qi] = cnt;
        if (out_stats) out_stats[qi] = st;
      }
    }
  };
  if (n_threads == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Exact path — SQL TopK, src/sql/executor.rs:2239-2379 with the ORDER BY arithmetic of :169-212.
// ----------------------------------------------------------------------------------------------
static double sql_order_distance(int op, const float* row, const float* query, uint32_t dim) {
  const double NULLV = std::numeric_limits<double>::quiet_NaN();
  if (op == TDO_L2) {  // executor.rs:177-188
    double s = 0.0;
    for (uint32_t i = 0; i < dim; ++i) {
      float df = row[i] - query[i];
      double d = (double)df;
      s += d * d;
    }
    return std::sqrt(s);
  }
  if (op == TDO_COSINE) {  // executor.rs:190-212
    double dot = 0.0, ml = 0.0, mr = 0.0;
    for (uint32_t i = 0; i < dim; ++i) dot += (double)row[i] * (double)query[i];
    for (uint32_t i = 0; i < dim; ++i) ml += (double)row[i] * (double)row[i];
    for (uint32_t i = 0; i < dim; ++i) mr += (double)query[i] * (double)query[i];
    ml = std::sqrt(ml);
    mr = std::sqrt(mr);
    if (ml > 0.0 && mr > 0.0) return 1.0 - dot / (ml * mr);
    return NULLV;
  }
  return NULLV;  // `<#>` falls to `_ => Value::Null`, executor.rs:241
}

// compare_values_for_sort on floats: partial_cmp, None (NULL/NaN) => Equal
static inline int cmp_sort(double a, double b) { return a < b ? -1 : (a > b ? 1 : 0); }

int tdo_sql_topk(const float* vectors, uint64_t n, uint32_t dim, const float* queries, uint32_t nq,
                 uint32_t limit, uint32_t offset, int op, uint64_t* out_rows, double* out_dist,
                 uint32_t* out_counts, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::atomic<uint32_t> next{0};
  struct Row {
    uint64_t row;
    double d;
  };
  auto worker = [&]() {
    std::vector<Row> heap;
    for (;;) {
      uint32_t qi = next.fetch_add(1);
      if (qi >= nq) break;
      const float* q = queries + (size_t)qi * dim;
      const size_t heap_size = (size_t)limit + offset;
      heap.clear();
      for (uint64_t r = 0; r < n && heap_size > 0; ++r) {
        double d = sql_order_distance(op, vectors + r * dim, q, dim);
        if (heap.size() < heap_size) {
          heap.push_back(Row{r, d});
          if (heap.size() == heap_size)  // :2260-2275: stable sort, worst first
            std::stable_sort(heap.begin(), heap.end(),
                             [](const Row& a, const Row& b) { return cmp_sort(a.d, b.d) > 0; });
        } else {
          if (cmp_sort(d, heap[0].d) < 0) {  // :2277-2290 strictly less replaces the root
            heap[0] = Row{r, d};
            size_t i = 0, len = heap.size();
            for (;;) {  // :2292-2357
              size_t left = 2 * i + 1, right = 2 * i + 2, largest = i;
              if (left < len && cmp_sort(heap[left].d, heap[largest].d) > 0) largest = left;
              if (right < len && cmp_sort(heap[right].d, heap[largest].d) > 0) largest = right;
              if (largest == i) break;
              std::swap(heap[i], heap[largest]);
              i = largest;
            }
          }
        }
      }
      std::stable_sort(heap.begin(), heap.end(),
                       [](const Row& a, const Row& b) { return cmp_sort(a.d, b.d) < 0; });
      size_t start = std::min<size_t>(offset, heap.size());
      size_t end = std::min<size_t>((size_t)offset + limit, heap.size());
      uint32_t cnt = 0;
      for (size_t i = start; i < end; ++i, ++cnt) {
        out_rows[(size_t)qi * limit + cnt] = heap[i].row;
        out_dist[(size_t)qi * limit + cnt] = heap[i].d;
      }
      for (uint32_t j = cnt; j < limit; ++j) {
        out_rows[(size_t)qi * limit + j] = ~0ull;
        out_dist[(size_t)qi * limit + j] = std::numeric_limits<double>::infinity();
      }
      out_counts[qi] = cnt;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  return 0;
}

// src/sql/predicate.rs:1634-1688 (f32, sequential; IP is +dot, cosine NULL on zero norm)
float tdo_sql_projection_distance(int op, const float* a, const float* b, uint32_t dim,
                                  int* is_null) {
  *is_null = 0;
  if (op == TDO_L2) {
    float sum = 0.0f;
    for (uint32_t i = 0; i < dim; ++i) {
      volatile float t = (a[i] - b[i]) * (a[i] - b[i]);
      sum += t;
    }
    return std::sqrt(sum);
  }
  if (op == TDO_COSINE) {
    float dot = 0, n1 = 0, n2 = 0;
    for (uint32_t i = 0; i < dim; ++i) {
      volatile float t = a[i] * b[i];
      dot += t;
    }
    for (uint32_t i = 0; i < dim; ++i) {
      volatile float t = a[i] * a[i];
      n1 += t;
    }
    for (uint32_t i = 0; i < dim; ++i) {
      volatile float t = b[i] * b[i];
      n2 += t;
    }
    n1 = std::sqrt(n1);
    n2 = std::sqrt(n2);
    if (n1 == 0.0f || n2 == 0.0f) {
      *is_null = 1;
      return 0.0f;
    }
    return 1.0f - dot / (n1 * n2);
  }
  float dot = 0;
  for (uint32_t i = 0; i < dim; ++i) {
    volatile float t = a[i] * b[i];
    dot += t;
  }
  return dot;
}

// ----------------------------------------------------------------------------------------------
// Node wire format — HnswNode::write_to / read_from, src/hnsw/mod.rs:333-421:
// row_id u64 LE | max_level u8 | l0_count u8 | l0_count x (page u32 LE, slot u16 LE) |
// per level 1..=max_level: count u8 | count x 6 B.
// ----------------------------------------------------------------------------------------------
static void put_node_id(uint8_t* p, uint32_t page, uint16_t slot) {
  p[0] = page & 0xFF; p[1] = (page >> 8) & 0xFF; p[2] = (page >> 16) & 0xFF; p[3] = (page >> 24) & 0xFF;
  p[4] = slot & 0xFF; p[5] = (slot >> 8) & 0xFF;
}

int64_t tdo_node_write(uint64_t row_id, uint8_t max_level, const uint32_t* l0_pages,
                       const uint16_t* l0_slots, uint8_t l0_count, const uint32_t* up_pages,
                       const uint16_t* up_slots, const uint8_t* up_counts, uint8_t* buf,
                       uint64_t buf_len) {
  uint64_t need = 10 + (uint64_t)l0_count * 6;
  for (uint8_t l = 0; l < max_level; ++l) need += 1 + (uint64_t)up_counts[l] * 6;
  if (need > buf_len) return -1;
  uint64_t off = 0;
  for (int i = 0; i < 8; ++i) buf[off++] = (row_id >> (8 * i)) & 0xFF;
  buf[off++] = max_level;
  buf[off++] = l0_count;
  for (uint8_t i = 0; i < l0_count; ++i, off += 6) put_node_id(buf + off, l0_pages[i], l0_slots[i]);
  for (uint8_t l = 0; l < max_level; ++l) {
    buf[off++] = up_counts[l];
    for (uint8_t i = 0; i < up_counts[l]; ++i, off += 6)
      put_node_id(buf + off, up_pages[l * TDO_MAX_LEVEL_NEIGHBORS + i],
                  up_slots[l * TDO_MAX_LEVEL_NEIGHBORS + i]);
  }
  return (int64_t)off;
}

int tdo_node_read(const uint8_t* buf, uint64_t len, uint64_t* row_id, uint8_t* max_level,
                  uint32_t* l0_pages, uint16_t* l0_slots, uint8_t* l0_count, uint32_t* up_pages,
                  uint16_t* up_slots, uint8_t* up_counts) {
  if (len < 10) return 1;  // "buffer too small for HnswNode header"
  uint64_t rid = 0;
  for (int i = 0; i < 8; ++i) rid |= (uint64_t)buf[i] << (8 * i);
  *row_id = rid;
  *max_level = buf[8];
  *l0_count = buf[9];
  if (*l0_count > TDO_MAX_L0_NEIGHBORS) return 2;  // "l0_count exceeds maximum"
  uint64_t off = 10;
  auto get = [&](uint32_t* page, uint16_t* slot) {
    *page = (uint32_t)buf[off] | ((uint32_t)buf[off + 1] << 8) | ((uint32_t)buf[off + 2] << 16) |
            ((uint32_t)buf[off + 3] << 24);
    *slot = (uint16_t)(buf[off + 4] | (buf[off + 5] << 8));
    off += 6;
  };
  for (uint8_t i = 0; i < *l0_count; ++i) {
    if (off + 6 > len) return 3;
    get(&l0_pages[i], &l0_slots[i]);
  }
  for (uint8_t l = 0; l < *max_level; ++l) {
    if (off >= len) return 4;
    uint8_t c = buf[off++];
    up_counts[l] = c;
    for (uint8_t i = 0; i < c; ++i) {
      if (off + 6 > len) return 5;
      if (i < TDO_MAX_LEVEL_NEIGHBORS)
        get(&up_pages[l * TDO_MAX_LEVEL_NEIGHBORS + i], &up_slots[l * TDO_MAX_LEVEL_NEIGHBORS + i]);
      else
        off += 6;
    }
  }
  return 0;
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------
// `.hnsw` file — the reference's WRITE path restated, so tests can hand the product reader a file laid
// out the way a TurDB process would have left it:
//   PersistentHnswIndex::create (mod.rs:776-809) -> HnswStorage::create (storage.rs:693-722, page 0 =
//   HnswFileHeader::new, storage.rs:121-158), then per node allocate_node (mod.rs:883-904):
//   page_has_space/can_fit (storage.rs:622-625) else allocate_page (storage.rs:746-757, HnswPage::init
//   :545-565 with PageHeader::new, page.rs:130-141), allocate_slot (storage.rs:627-650), write_node_data
//   (storage.rs:652-669, THROUGH the decoded 13-bit offset, :338-356), finally sync (mod.rs:877-881,
//   sync_to_header :662-666).  All header fields are read back from the page bytes like the reference does.
// The records hold each node's FINAL neighbour lists (the reference reaches the same bytes through
// update_node, mod.rs:913-935, as long as records do not overlap).
// mode 0 = verbatim can_fit rule (records overlap once a page holds more than ~39 of them: the 13-bit
//          offset defect); mode 1 = additionally start a new page before a record's truncated offset would
//          reach the slot directory (a file every record of which reads back intact).
// ----------------------------------------------------------------------------------------------
namespace {
constexpr size_t kPage = 16384, kHnswHdr = 64;
inline void w16(uint8_t* p, uint16_t v) { p[0] = v & 0xFF; p[1] = v >> 8; }
inline void w32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (v >> (8 * i)) & 0xFF; }
inline void w64(uint8_t* p, uint64_t v) { for (int i = 0; i < 8; ++i) p[i] = (v >> (8 * i)) & 0xFF; }
inline uint16_t r16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
// HnswNode::max_serialized_size, mod.rs:252-257
inline size_t max_serialized_size(uint8_t max_level) { return 10 + 32 * 6 + (size_t)max_level * (1 + 16 * 6); }
}  // namespace

extern "C" int64_t tdo_hnsw_file_write(const tdo_graph* g, uint64_t index_id, uint64_t table_id, uint16_t ef_search,
                                       int distance_fn, int quantization, int mode, uint8_t* buf, uint64_t buf_len,
                                       uint32_t* out_pages, uint16_t* out_slots) {
  const size_t n = g->n();
  std::vector<uint8_t> file(kPage, 0);
  // HnswFileHeader::new + write_to
  memcpy(file.data(), "TurDB HNSW\0\0\0\0\0\0", 16);
  w64(&file[16], index_id);
  w64(&file[24], table_id);
  w16(&file[32], g->dim);
  w16(&file[34], g->m);
  w16(&file[36], (uint16_t)(g->m * 2));
  w16(&file[38], g->efc);
  w16(&file[40], ef_search);
  file[42] = (uint8_t)distance_fn;
  file[43] = (uint8_t)quantization;
  w32(&file[44], 0xFFFFFFFFu);
  w16(&file[48], 0xFFFF);
  // pass 1: NodeIds.  The page state lives in the page bytes, as in the reference.
  std::vector<uint32_t> pages(n);
  std::vector<uint16_t> slots(n);
  uint32_t current_page = 0;
  auto page_ptr = [&](uint32_t p) { return file.data() + (size_t)p * kPage; };
  std::vector<std::pair<size_t, size_t>> where(n);  // byte position + slot size of every record
  for (size_t i = 0; i < n; ++i) {
    const size_t max_size = max_serialized_size(g->levels[i]);
    bool fits = false;
    if (current_page != 0) {
      uint8_t* pg = page_ptr(current_page);
      if (pg[0] != 0x10) return -2;  // HnswPage::from_bytes: "not an HNSW node page" (header overwritten)
      const uint16_t fs = r16(pg + 16 + 2), fe = r16(pg + 16 + 4);
      const size_t space = fe > fs ? (size_t)(fe - fs) : 0;       // free_space(): saturating_sub
      fits = space >= max_size + 4 + 64;                            // can_fit
      if (fits && mode == 1) {
        const size_t new_end = (size_t)fe - max_size, dir_end = (size_t)fs + 4;
        if (new_end < 8192 + dir_end) fits = false;                 // truncated offset would meet the directory
      }
    }
    if (!fits) {
      current_page = (uint32_t)(file.size() / kPage);
      file.resize(file.size() + kPage, 0);
      uint8_t* pg = page_ptr(current_page);
      pg[0] = 0x10;                    // PageHeader::new(PageType::HnswNode)
      w16(pg + 4, 16);                 // free_start = PAGE_HEADER_SIZE
      w16(pg + 6, (uint16_t)kPage);    // free_end = PAGE_SIZE
      w16(pg + 16 + 0, 0);             // HnswPageHeader::new
      w16(pg + 16 + 2, (uint16_t)kHnswHdr);
      w16(pg + 16 + 4, (uint16_t)kPage);
      w16(pg + 16 + 10, (uint16_t)(kPage - kHnswHdr));
    }
    uint8_t* pg = page_ptr(current_page);
    // allocate_slot
    const uint16_t slot_index = r16(pg + 16 + 0);
    const uint16_t new_fs = (uint16_t)(r16(pg + 16 + 2) + 4);
    const uint16_t new_fe = (uint16_t)(r16(pg + 16 + 4) - (uint16_t)max_size);
    w16(pg + 16 + 0, (uint16_t)(slot_index + 1));
    w16(pg + 16 + 2, new_fs);
    w16(pg + 16 + 4, new_fe);
    w16(pg + 16 + 6, (uint16_t)(r16(pg + 16 + 6) + 1));
    w16(pg + 16 + 10, (uint16_t)(r16(pg + 16 + 10) - (uint16_t)max_size - 4));
    uint8_t* se = pg + kHnswHdr + (size_t)slot_index * 4;
    w16(se, (uint16_t)((new_fe & 0x1FFF) | (1u << 13)));  // SlotEntry::encode, status Active
    w16(se + 2, (uint16_t)max_size);
    pages[i] = current_page;
    slots[i] = slot_index;
    where[i] = {(size_t)current_page * kPage + (new_fe & 0x1FFF), max_size};  // write_node_data: decoded offset
    // a placeholder record of the node's own size is written now so later allocations see the same bytes the
    // reference's insert would have left (row id / level / empty lists); the final record replaces it below
    uint8_t* rec = file.data() + where[i].first;
    if (where[i].first + 10 + g->levels[i] <= file.size()) {
      w64(rec, g->row_ids[i]);
      rec[8] = g->levels[i];
      rec[9] = 0;
      for (uint8_t l = 0; l < g->levels[i]; ++l) rec[10 + l] = 0;
    }
  }
  // pass 2: final records
  for (size_t i = 0; i < n; ++i) {
    uint8_t rec[10 + 32 * 6 + 255 * 97];
    size_t off = 0;
    w64(rec, g->row_ids[i]);
    rec[8] = g->levels[i];
    rec[9] = g->l0_cnt[i];
    off = 10;
    for (uint8_t j = 0; j < g->l0_cnt[i]; ++j, off += 6) {
      const uint32_t nb = g->l0_adj[i * TDO_MAX_L0_NEIGHBORS + j];
      put_node_id(rec + off, pages[nb], slots[nb]);
    }
    for (uint8_t l = 1; l <= g->levels[i]; ++l) {
      const size_t slot = (size_t)g->up_base[i] + (l - 1);
      rec[off++] = g->up_cnt[slot];
      for (uint8_t j = 0; j < g->up_cnt[slot]; ++j, off += 6) {
        const uint32_t nb = g->up_adj[slot * TDO_MAX_LEVEL_NEIGHBORS + j];
        put_node_id(rec + off, pages[nb], slots[nb]);
      }
    }
    if (off > where[i].second) return -3;  // write_node_data: "data size exceeds slot size"
    memcpy(file.data() + where[i].first, rec, off);
  }
  // sync_to_header
  if (g->entry != TDO_INVALID) {
    w32(&file[44], pages[g->entry]);
    w16(&file[48], slots[g->entry]);
  }
  file[50] = g->max_level;
  w64(&file[52], n);
  if (out_pages && n) memcpy(out_pages, pages.data(), n * 4);
  if (out_slots && n) memcpy(out_slots, slots.data(), n * 2);
  if (buf) {
    if (buf_len < file.size()) return -1;
    memcpy(buf, file.data(), file.size());
  }
  return (int64_t)file.size();
}

// ----------------------------------------------------------------------------------------------
// SQ8 — SQ8Vector::from_f32 / decode, src/hnsw/quantization.rs:68-95, 108-113.  (The reference never wires
// SQ8 into the index; an SQ8 traversal is DEFINED here as the FP32 search over the decoded vectors.)
// ----------------------------------------------------------------------------------------------
extern "C" void tdo_sq8_encode(const float* values, uint32_t dim, uint8_t* codes, float* out_min, float* out_scale) {
  if (dim == 0) {
    *out_min = 0.f;
    *out_scale = 0.f;
    return;
  }
  float mn = INFINITY, mx = -INFINITY;
  for (uint32_t i = 0; i < dim; ++i) {  // fold(f32::INFINITY, f32::min) / fold(NEG_INFINITY, f32::max)
    mn = std::fmin(mn, values[i]);
    mx = std::fmax(mx, values[i]);
  }
  const float range = mx - mn;
  const float scale = range > 0.0f ? range / 255.0f : 1.0f;
  for (uint32_t i = 0; i < dim; ++i) {
    if (scale == 0.0f || range == 0.0f) {
      codes[i] = 0;
    } else {
      float q = std::round((values[i] - mn) / scale);  // f32::round: half away from zero
      q = q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q);
      codes[i] = (uint8_t)q;
    }
  }
  *out_min = mn;
  *out_scale = scale;
}

extern "C" void tdo_sq8_decode(const uint8_t* codes, uint32_t dim, float mn, float scale, float* out) {
  for (uint32_t i = 0; i < dim; ++i) {
    volatile float prod = (float)codes[i] * scale;  // two roundings: rustc does not contract `min + q * scale`
    out[i] = mn + prod;
  }
}
