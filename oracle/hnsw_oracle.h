/*
 * hnsw_oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * The oracle is a CPU restatement of kahflane/TurDB's HNSW search / insert / distance code and of
 * its SQL brute-force TopK, used ONLY as the checker in tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under turdb_b200/ may link or call it.
 *
 * PARITY STATUS: "parity unpinned" for HNSW search/insert/distance — the reference holds no test
 * that pins those and no Rust toolchain exists here to run it.  Pinned parts: the
 * node wire format round-trip (tests/hnsw_integration.rs:120-140) and the three SQL k-NN
 * known answers (tests/hnsw_integration.rs:220-276), see tests/test_oracle_kat.py.
 *
 * All citations are into /root/reference (kahflane/TurDB).
 */
#ifndef TURDB_HNSW_ORACLE_H
#define TURDB_HNSW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDO_MAX_L0_NEIGHBORS 32   /* src/hnsw/mod.rs:126 */
#define TDO_MAX_LEVEL_NEIGHBORS 16 /* src/hnsw/mod.rs:127 */
#define TDO_INVALID 0xFFFFFFFFu

enum { TDO_L2 = 0, TDO_COSINE = 1, TDO_IP = 2 }; /* src/hnsw/mod.rs:129-137 */
enum { TDO_BUILD_VERBATIM = 0, TDO_BUILD_INTENT = 1 };
enum { TDO_DIST_AVX2 = 0, TDO_DIST_AVX2_EMULATED = 1, TDO_DIST_SCALAR = 2 };

typedef struct tdo_graph tdo_graph;

/* per-query traversal counters (the roofline's algorithmic-bytes inputs) */
typedef struct {
  uint32_t n_dist;       /* every distance evaluation (entry + upper levels + level 0) */
  uint32_t n_dist_upper; /* of which: entry + greedy upper-level evaluations */
  uint32_t n_expanded;   /* level-0 nodes whose adjacency row was read */
  uint32_t n_upper_hops; /* greedy steps (one upper-level adjacency row read each) */
} tdo_stats;

/* ---- distance (src/hnsw/distance.rs) ---- */
float tdo_distance(int metric, const float* a, const float* b, uint32_t dim, int impl);
uint8_t tdo_select_level(double random_value, uint16_t m); /* operations.rs:76-83 */

/* ---- graph: build by the reference's insert path (src/hnsw/mod.rs:999-1084) ---- */
tdo_graph* tdo_graph_new(uint16_t dim, uint16_t m, uint16_t ef_construction, int build_mode);
void tdo_graph_free(tdo_graph* g);
int tdo_graph_insert(tdo_graph* g, uint64_t row_id, const float* vec, double random_value);
int tdo_graph_insert_batch(tdo_graph* g, uint64_t n, const uint64_t* row_ids, const float* vecs,
                           const double* random_values);
/* adopt a flattened graph (same arrays the C-ABI index_create takes); vectors are copied */
tdo_graph* tdo_graph_from_arrays(uint16_t dim, uint64_t n, const float* vectors,
                                 const uint64_t* row_ids, const uint8_t* levels,
                                 const uint32_t* l0_adj, const uint8_t* l0_cnt,
                                 const uint32_t* up_base, const uint32_t* up_adj,
                                 const uint8_t* up_cnt, uint64_t n_up_slots, uint32_t entry,
                                 uint8_t max_level);
uint64_t tdo_graph_n(const tdo_graph* g);
uint64_t tdo_graph_n_up_slots(const tdo_graph* g);
uint32_t tdo_graph_entry(const tdo_graph* g); /* TDO_INVALID when empty */
uint8_t tdo_graph_max_level(const tdo_graph* g);
uint64_t tdo_graph_build_dist_evals(const tdo_graph* g);
int tdo_graph_export(const tdo_graph* g, float* vectors, uint64_t* row_ids, uint8_t* levels,
                     uint32_t* l0_adj, uint8_t* l0_cnt, uint32_t* up_base, uint32_t* up_adj,
                     uint8_t* up_cnt);

/* ---- search (src/hnsw/mod.rs:1092-1273, src/hnsw/search.rs:259-398) ----
 * visible: NULL => search(); else one bit per NODE id (bit i of word i/64) => search_filtered().
 * Outputs are [nq][k]; out_counts[q] = number of valid results.  Returns 0, or 2 on dim mismatch. */
int tdo_search_batch(const tdo_graph* g, const float* queries, uint32_t query_dim, uint32_t nq,
                     uint32_t k, uint32_t ef, int metric, const uint64_t* visible,
                     uint64_t* out_row_ids, uint32_t* out_node_ids, float* out_dist,
                     uint32_t* out_counts, tdo_stats* out_stats, int n_threads);

/* ---- exact path: SQL TopK (src/sql/executor.rs:2239-2379 + :169-212) ----
 * op: TDO_L2 => sqrt(sum_f64((a-b)_f32^2)); TDO_COSINE => 1 - dot/(|a||b|) in f64, NULL on zero norm;
 * TDO_IP => NULL for every row (executor.rs:241).  NULL is reported as NaN and compares Equal.
 * Rows are scanned in index order (primary-key order). */
int tdo_sql_topk(const float* vectors, uint64_t n, uint32_t dim, const float* queries, uint32_t nq,
                 uint32_t limit, uint32_t offset, int op, uint64_t* out_rows, double* out_dist,
                 uint32_t* out_counts, int n_threads);
/* projection-flavour distances (src/sql/predicate.rs:1634-1688), f32 sequential */
float tdo_sql_projection_distance(int op, const float* a, const float* b, uint32_t dim, int* is_null);

/* ---- node wire format (src/hnsw/mod.rs:333-421) ---- */
int64_t tdo_node_write(uint64_t row_id, uint8_t max_level, const uint32_t* l0_pages,
                       const uint16_t* l0_slots, uint8_t l0_count, const uint32_t* up_pages,
                       const uint16_t* up_slots, const uint8_t* up_counts, uint8_t* buf,
                       uint64_t buf_len);
int tdo_node_read(const uint8_t* buf, uint64_t len, uint64_t* row_id, uint8_t* max_level,
                  uint32_t* l0_pages, uint16_t* l0_slots, uint8_t* l0_count, uint32_t* up_pages,
                  uint16_t* up_slots, uint8_t* up_counts /* [max_level] , up arrays [max_level][16] */);

/* ---- SQ8 (src/hnsw/quantization.rs:68-95, 108-113) ---- */
void tdo_sq8_encode(const float* values, uint32_t dim, uint8_t* codes, float* out_min, float* out_scale);
void tdo_sq8_decode(const uint8_t* codes, uint32_t dim, float mn, float scale, float* out);

/* ---- .hnsw file: the reference's write path (storage.rs:121-158, 545-669, 693-757; mod.rs:776-904) ----
 * mode 0 = verbatim page-fill rule (records overlap: 13-bit slot offsets); mode 1 = new page before overlap.
 * buf may be NULL (size query).  out_pages/out_slots: NodeId of every dense node.  Returns the file length,
 * -1 buffer too small, -2 a page header was overwritten (the reference's insert would fail there), -3 record
 * larger than its slot. */
int64_t tdo_hnsw_file_write(const tdo_graph* g, uint64_t index_id, uint64_t table_id, uint16_t ef_search,
                            int distance_fn, int quantization, int mode, uint8_t* buf, uint64_t buf_len,
                            uint32_t* out_pages, uint16_t* out_slots);

#ifdef __cplusplus
}
#endif
#endif
