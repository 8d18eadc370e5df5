// C++ host mirror (include/turdb_cuda.hpp) end to end.  `--no-device`: the CPU-only checks (the library has no CPU
// fallback: index creation must fail loudly); default: a 64-point index against a brute force computed here.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "turdb_cuda.hpp"

using namespace turdb_cuda;

#define REQUIRE(c)                                                      \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

struct Corpus {
  uint32_t n = 64, dim = 8;
  std::vector<float> x;
  std::vector<uint64_t> row_ids;
  std::vector<uint8_t> levels, l0_cnt;
  std::vector<uint32_t> l0_adj, up_base;
  turdb_cuda_graph g{};
  float d2(const float* a, const float* b) const {
    float s = 0;
    for (uint32_t i = 0; i < dim; ++i) s += (a[i] - b[i]) * (a[i] - b[i]);
    return s;
  }
  Corpus() {
    uint32_t st = 12345;
    x.resize((size_t)n * dim);
    for (auto& v : x) {
      st = st * 1664525u + 1013904223u;
      v = (float)((st >> 8) & 0xFFFF) / 65536.0f;
    }
    row_ids.resize(n);
    for (uint32_t i = 0; i < n; ++i) row_ids[i] = 1000 + 7 * i;
    levels.assign(n, 0);
    l0_cnt.assign(n, 32);
    up_base.assign(n, TURDB_INVALID_NODE);
    l0_adj.assign((size_t)n * 32, TURDB_INVALID_NODE);
    for (uint32_t i = 0; i < n; ++i) {  // 32 nearest neighbours of every node: a navigable toy graph
      std::vector<uint32_t> o(n);
      std::iota(o.begin(), o.end(), 0u);
      std::sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return d2(&x[i * dim], &x[a * dim]) < d2(&x[i * dim], &x[b * dim]); });
      for (uint32_t j = 0; j < 32; ++j) l0_adj[i * 32 + j] = o[j + 1];
    }
    g.dim = dim;
    g.max_level = 0;
    g.n = n;
    g.entry = 0;
    g.vectors = x.data();
    g.row_ids = row_ids.data();
    g.levels = levels.data();
    g.l0_adj = l0_adj.data();
    g.l0_cnt = l0_cnt.data();
    g.up_base = up_base.data();
    g.up_adj = nullptr;
    g.up_cnt = nullptr;
    g.n_up_slots = 0;
  }
  std::vector<uint32_t> exact(const std::vector<float>& q) const {
    std::vector<uint32_t> o(n);
    std::iota(o.begin(), o.end(), 0u);
    std::sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return d2(q.data(), &x[a * dim]) < d2(q.data(), &x[b * dim]); });
    return o;
  }
};

int main(int argc, char** argv) {
  const bool no_device = argc > 1 && std::strcmp(argv[1], "--no-device") == 0;
  Corpus c;
  REQUIRE(turdb_cuda_abi_version() == TURDB_CUDA_ABI_VERSION);
  try {  // a missing file is an error, not a crash
    CudaHnswIndex::open("/nonexistent/idx.hnsw", [](uint64_t) { return std::optional<std::vector<float>>{}; });
    REQUIRE(false);
  } catch (const Error& e) {
    REQUIRE(e.status() == TURDB_ERR_INVALID_ARGUMENT);
  }
  if (no_device) {
    try {
      CudaHnswIndex::from_graph(c.g);
      REQUIRE(false);  // no CPU fallback
    } catch (const Error& e) {
      REQUIRE(e.status() == TURDB_ERR_NO_DEVICE);
      REQUIRE(std::string(e.what()).find("no CPU fallback") != std::string::npos);
    }
    std::puts("cpp host mirror: no-device checks ok");
    return 0;
  }
  CudaHnswIndex idx = CudaHnswIndex::from_graph(c.g, 0, DistanceFunction::L2);
  REQUIRE(idx.dimensions() == 8 && idx.node_count() == 64);
  std::vector<float> q(c.x.begin() + 5 * 8, c.x.begin() + 6 * 8);
  q[0] += 0.01f;
  const auto truth = c.exact(q);
  HnswSearchContext ctx(64);
  auto r = idx.search(q, 5, ctx);
  REQUIRE(r.size() == 5);
  for (size_t i = 0; i < 5; ++i) {
    REQUIRE(r[i].node_id == truth[i] && r[i].row_id == 1000 + 7 * (uint64_t)truth[i]);
    REQUIRE(std::fabs(r[i].distance - c.d2(q.data(), &c.x[truth[i] * 8])) <= 1e-5f * std::max(1e-3f, r[i].distance));
    REQUIRE(i == 0 || r[i - 1].distance <= r[i].distance);
  }
  // filtered: hide the nearest row, the rest shifts up (invisible nodes are traversed but not returned)
  const uint64_t hidden = r[0].row_id;
  auto rf = idx.search_filtered(q, 5, ctx, c.row_ids, [&](uint64_t row) { return row != hidden; });
  REQUIRE(rf.size() == 5 && rf[0].node_id == truth[1] && rf[4].node_id == truth[5]);
  // batch and exact path agree with the single search
  std::vector<float> two(q);
  two.insert(two.end(), c.x.begin(), c.x.begin() + 8);
  auto rb = idx.search_batch(two, 2, 5, 64);
  REQUIRE(rb.size() == 2 && rb[0].size() == 5 && rb[0][0].node_id == truth[0] && rb[1][0].node_id == 0 && rb[1][0].distance == 0.0f);
  auto re = idx.bruteforce_topk(q, 5);
  REQUIRE(re.size() == 5 && re[0].node_id == truth[0] && re[4].node_id == truth[4]);
  // dimension mismatch: the reference's message (mod.rs:1099-1104)
  try {
    idx.search(std::vector<float>(7, 0.f), 5, ctx);
    REQUIRE(false);
  } catch (const Error& e) {
    REQUIRE(e.status() == TURDB_ERR_DIMENSION_MISMATCH);
    REQUIRE(std::string(e.what()) == "query dimension 7 does not match index dimension 8");
  }
  // SQL operator: open / next / close, f64 key = sqrt(sum_f64((a - b)_f32^2)) (executor.rs:174-183)
  VectorTopKExec ex(idx, VectorOp::L2Distance, q, 3, 1);
  ex.open();
  for (int i = 0; i < 3; ++i) {
    auto row = ex.next();
    REQUIRE(row.has_value() && row->first == 1000 + 7 * (uint64_t)truth[i + 1]);
    double s = 0;
    for (int j = 0; j < 8; ++j) {
      const double d = (double)(c.x[truth[i + 1] * 8 + j] - q[j]);
      s += d * d;
    }
    REQUIRE(row->second == std::sqrt(s));
  }
  REQUIRE(!ex.next().has_value());
  ex.close();
  // empty index: Ok(vec![]) (mod.rs:1106-1109)
  turdb_cuda_graph eg{};
  eg.dim = 8;
  eg.entry = TURDB_INVALID_NODE;
  CudaHnswIndex empty = CudaHnswIndex::from_graph(eg);
  REQUIRE(empty.search(q, 5, ctx).empty());
  std::puts("cpp host mirror: all checks ok");
  return 0;
}
