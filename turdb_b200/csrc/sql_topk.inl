// sql_topk.inl — the SQL vector-scan operator: `ORDER BY vec <op> '[...]' LIMIT k OFFSET o` for a batch of
// statements (TopKExec, src/sql/executor.rs:2239-2392; sort-key arithmetic :169-212).  SURVEY.md §8(f) rank 3,
// BASELINE.json config 5's "SQL ORDER BY distance LIMIT 10 batch path".  Included by turdb_cuda.cu.
//
// The reference evaluates an f64 sort key per row (L2: sqrt(sum_f64(((a-b) as f32 -> f64)^2)); cosine:
// 1 - dot/(|a||b|) in f64, NULL when a norm is 0) and keeps the limit+offset smallest.  Here the candidate
// rows come from the HNSW metric contract in FP32 (exact path: tensor-core filter + FP32 rerank; or the graph
// traversal when use_index != 0), a margin wider than limit+offset, and the f64 key is then evaluated for
// those candidates only, in the reference's summation order, and the rows re-sorted by (key, scan position).
// Exact scan: a query is CERTIFIED when its (limit+offset)-th f64 key lies below every row outside the
// candidate set by more than the FP32/f64 discrepancy bound; an uncertified query is reported
// (count 0xFFFFFFFD) and the host retries it with a wider margin.

namespace turdb {

// One warp per statement; lane j evaluates candidates j, j+32, ...  Keys in shared memory, bitonic sort by
// (key, node id) with NULL (NaN) keys last, rows [offset, offset+limit) written out.
template <int OP>
__global__ void __launch_bounds__(128) sql_rekey_kernel(DeviceIndex ix, const float* __restrict__ queries, uint32_t nq,
                                                        uint32_t kq, uint32_t limit, uint32_t offset,
                                                        const uint32_t* __restrict__ cand_nodes,
                                                        const float* __restrict__ cand_dist,
                                                        const uint32_t* __restrict__ cand_counts, int certify,
                                                        uint64_t* out_rows, double* out_keys, uint32_t* out_counts) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const uint32_t q = blockIdx.x * wpb + warp;
  uint32_t n2 = 1;
  while (n2 < kq) n2 <<= 1;
  double* sk = reinterpret_cast<double*>(smem_raw) + (size_t)warp * n2;                 // [wpb][n2]
  uint32_t* si = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)wpb * n2) + (size_t)warp * n2;
  if (q >= nq) return;
  const uint32_t cnt = min(cand_counts[q], kq);
  const float* a = queries + (size_t)q * ix.dim;
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  double mag_q = 0.0;
  if (OP == 1) {  // mag_r of the literal: sum of x.powi(2) in f64, sequential (executor.rs:198)
    double s = 0.0;
    for (uint32_t i = 0; i < ix.dim; ++i) {
      const double x = (double)__ldg(a + i);
      s = __dadd_rn(s, __dmul_rn(x, x));
    }
    mag_q = __dsqrt_rn(s);
  }
  for (uint32_t c = lane; c < n2; c += 32) {
    double key = nan;
    uint32_t id = 0xFFFFFFFFu;
    if (c < cnt) {
      id = cand_nodes[(size_t)q * kq + c];
      const float* b = ix.arena + (size_t)id * ix.ds;
      if (OP == 0) {  // executor.rs:174-183: row - literal in f32, squared and summed in f64, sqrt
        double s = 0.0;
        for (uint32_t i = 0; i < ix.dim; ++i) {
          const double d = (double)__fsub_rn(b[i], __ldg(a + i));
          s = __dadd_rn(s, __dmul_rn(d, d));
        }
        key = __dsqrt_rn(s);
      } else {        // executor.rs:188-206
        double dot = 0.0, sl = 0.0;
        for (uint32_t i = 0; i < ix.dim; ++i) {
          const double x = (double)b[i];
          dot = __dadd_rn(dot, __dmul_rn(x, (double)__ldg(a + i)));
          sl = __dadd_rn(sl, __dmul_rn(x, x));
        }
        const double mag_l = __dsqrt_rn(sl);
        if (mag_l > 0.0 && mag_q > 0.0) key = __dsub_rn(1.0, __ddiv_rn(dot, __dmul_rn(mag_l, mag_q)));
      }
    }
    sk[c] = key;
    si[c] = id;
  }
  __syncwarp();
  auto first = [](double ka, uint32_t ia, double kb, uint32_t ib) {
    const bool na = ka != ka, nb = kb != kb;  // NULL keys after every number; absent entries (id INVALID) last
    if (na != nb) return nb;
    if (!na && ka != kb) return ka < kb;
    return ia < ib;
  };
  for (uint32_t size = 2; size <= n2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = lane; i < n2; i += 32) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool asc = (i & size) == 0;
          const double ki = sk[i], kj = sk[j];
          const uint32_t ii = si[i], ij = si[j];
          if (first(ki, ii, kj, ij) != asc) {
            sk[i] = kj; sk[j] = ki;
            si[i] = ij; si[j] = ii;
          }
        }
      }
      __syncwarp();
    }
  }
  const uint32_t K = limit + offset;
  bool certified = true;
  if (certify && cnt == kq && K > 0 && cnt > 0) {
    // rows outside the candidate set have an FP32 metric >= the largest candidate's (cand_dist is ascending)
    const float worst32 = cand_dist[(size_t)q * kq + cnt - 1];
    const uint32_t kth = min(K, cnt) - 1;
    const double kth_key = sk[kth];
    double bound;
    if (OP == 0) bound = sqrt(fmax(0.0, (double)worst32 * (1.0 - 1e-4)));  // squared L2 in FP32 vs f64: rel. 1e-4 >> dim * 2^-24
    else bound = (double)worst32 - 1e-5;
    certified = !(kth_key != kth_key) && kth_key < bound;
  }
  const uint32_t have = cnt > offset ? min(cnt - offset, limit) : 0u;
  for (uint32_t i = lane; i < limit; i += 32) {
    const size_t o = (size_t)q * limit + i;
    if (i < have && certified) {
      out_rows[o] = ix.row_ids[si[offset + i]];
      out_keys[o] = sk[offset + i];
    } else {
      out_rows[o] = 0xFFFFFFFFFFFFFFFFull;
      out_keys[o] = nan;
    }
  }
  if (lane == 0) out_counts[q] = certified ? have : 0xFFFFFFFDu;
}

}  // namespace turdb

extern "C" int32_t turdb_cuda_sql_topk_batch_device(turdb_cuda_index* idx, const float* d_queries, uint32_t query_dim,
                                                    uint32_t nq, uint32_t limit, uint32_t offset, uint8_t op,
                                                    uint32_t margin, int32_t use_index, uint32_t ef,
                                                    uint64_t* d_out_row_ids, double* d_out_keys, uint32_t* d_out_counts,
                                                    void* stream_) {
  if (!idx) return fail(TURDB_ERR_INVALID_ARGUMENT, "idx is null");
  if (query_dim != idx->ix.dim)
    return fail(TURDB_ERR_DIMENSION_MISMATCH, "query dimension %u does not match index dimension %u", query_dim, idx->ix.dim);
  if (op > 1)
    return fail(TURDB_ERR_UNSUPPORTED, "ORDER BY <#> evaluates to NULL for every row in the reference (executor.rs:241); "
                                       "only <-> (0) and <=> (1) are sort keys");
  if (nq == 0) return TURDB_OK;
  if (!d_queries || !d_out_counts || (limit && (!d_out_row_ids || !d_out_keys)))
    return fail(TURDB_ERR_INVALID_ARGUMENT, "null query/output pointer");
  const uint64_t K = (uint64_t)limit + offset;
  if (K > 1024) return fail(TURDB_ERR_UNSUPPORTED, "limit + offset = %llu > 1024", (unsigned long long)K);
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", idx->device);
  const uint64_t n = idx->ix.n;
  if (!margin) margin = (uint32_t)std::max<uint64_t>(8, K / 4);
  const uint32_t kq = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(K + margin, 1), std::max<uint64_t>(n, 1));
  if (use_index && ef < kq) ef = kq;

  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_rows = take((size_t)nq * kq * 8), o_nodes = take((size_t)nq * kq * 4), o_dist = take((size_t)nq * kq * 4),
               o_cnt = take((size_t)nq * 4);
  uint8_t* scr = nullptr;
  CUDA_TRY(cudaMallocFromPoolAsync(&scr, off, idx->pool, stream));
  uint64_t* c_rows = (uint64_t*)(scr + o_rows);
  uint32_t* c_nodes = (uint32_t*)(scr + o_nodes);
  float* c_dist = (float*)(scr + o_dist);
  uint32_t* c_cnt = (uint32_t*)(scr + o_cnt);
  int32_t rc;
  if (use_index)
    rc = turdb_cuda_search_batch_device(idx, d_queries, query_dim, nq, kq, ef, op, nullptr, c_rows, c_nodes, c_dist, c_cnt,
                                        nullptr, stream);
  else
    rc = turdb_cuda_bruteforce_topk_device(idx, d_queries, query_dim, nq, kq, op, 0, c_rows, c_nodes, c_dist, c_cnt, stream);
  if (rc != TURDB_OK) {
    cudaFreeAsync(scr, stream);
    return rc;
  }
  uint32_t n2 = 1;
  while (n2 < kq) n2 <<= 1;
  const uint32_t wpb = 4;
  const size_t smem = (size_t)wpb * n2 * 12;
  const uint32_t blocks = (nq + wpb - 1) / wpb;
  const int certify = (!use_index && kq < n) ? 1 : 0;
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(sql_rekey_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sql_rekey_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (e == cudaSuccess) {
    if (op == 0)
      sql_rekey_kernel<0><<<blocks, wpb * 32, smem, stream>>>(idx->ix, d_queries, nq, kq, limit, offset, c_nodes, c_dist, c_cnt,
                                                              certify, d_out_row_ids, d_out_keys, d_out_counts);
    else
      sql_rekey_kernel<1><<<blocks, wpb * 32, smem, stream>>>(idx->ix, d_queries, nq, kq, limit, offset, c_nodes, c_dist, c_cnt,
                                                              certify, d_out_row_ids, d_out_keys, d_out_counts);
    e = cudaGetLastError();
  }
  cudaFreeAsync(scr, stream);
  if (e != cudaSuccess) return fail(TURDB_ERR_CUDA, "sql_rekey launch failed: %s", cudaGetErrorString(e));
  return TURDB_OK;
}

extern "C" int32_t turdb_cuda_sql_topk_batch(turdb_cuda_index* idx, const float* queries, uint32_t query_dim, uint32_t nq,
                                             uint32_t limit, uint32_t offset, uint8_t op, int32_t use_index, uint32_t ef,
                                             uint64_t* out_row_ids, double* out_keys, uint32_t* out_counts) {
  if (!idx) return fail(TURDB_ERR_INVALID_ARGUMENT, "idx is null");
  if (query_dim != idx->ix.dim)
    return fail(TURDB_ERR_DIMENSION_MISMATCH, "query dimension %u does not match index dimension %u", query_dim, idx->ix.dim);
  if (nq == 0) return TURDB_OK;
  if (!queries || !out_counts || (limit && (!out_row_ids || !out_keys)))
    return fail(TURDB_ERR_INVALID_ARGUMENT, "null query/output pointer");
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", idx->device);
  cudaStream_t stream;
  CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const size_t ll = std::max(limit, 1u);
  const size_t qbytes = (size_t)nq * query_dim * 4;
  const size_t off_rows = (qbytes + 255) & ~(size_t)255, off_keys = off_rows + nq * ll * 8, off_counts = off_keys + nq * ll * 8,
               total = off_counts + (size_t)nq * 4;
  uint8_t* slab = nullptr;
  cudaError_t e = cudaMallocFromPoolAsync(&slab, total, idx->pool, stream);
  if (e != cudaSuccess) {
    cudaStreamDestroy(stream);
    return fail(TURDB_ERR_OUT_OF_MEMORY, "cudaMallocFromPoolAsync(%zu) failed: %s", total, cudaGetErrorString(e));
  }
  auto cleanup = [&]() {
    cudaFreeAsync(slab, stream);
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
  };
  e = cudaMemcpyAsync(slab, queries, qbytes, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) {
    cleanup();
    return fail(TURDB_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  }
  // a query the exact scan cannot certify at the default margin is retried with 4x, then 16x the margin
  uint32_t margin = 0;
  const uint64_t K = (uint64_t)limit + offset;
  for (int attempt = 0;; ++attempt) {
    int32_t rc = turdb_cuda_sql_topk_batch_device(idx, (const float*)slab, query_dim, nq, limit, offset, op, margin, use_index,
                                                  ef, (uint64_t*)(slab + off_rows), (double*)(slab + off_keys),
                                                  (uint32_t*)(slab + off_counts), stream);
    if (rc != TURDB_OK) {
      cleanup();
      return rc;
    }
    e = cudaMemcpyAsync(out_counts, slab + off_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) break;
    bool uncertified = false;
    for (uint32_t i = 0; i < nq; ++i) uncertified |= out_counts[i] == 0xFFFFFFFDu;
    if (!uncertified) break;
    if (attempt == 2) {
      cleanup();
      return fail(TURDB_ERR_UNSUPPORTED, "sql_topk: result not certifiable (more than %llu rows tie with the limit-th key)",
                  (unsigned long long)(K + margin));
    }
    margin = (uint32_t)std::max<uint64_t>(8, K / 4) * (attempt == 0 ? 4u : 16u);
  }
  if (e == cudaSuccess && limit) {
    e = cudaMemcpyAsync(out_row_ids, slab + off_rows, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_keys, slab + off_keys, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  }
  cleanup();
  if (e != cudaSuccess) return fail(TURDB_ERR_CUDA, "sql_topk failed: %s", cudaGetErrorString(e));
  return TURDB_OK;
}
