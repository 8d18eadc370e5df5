// sql_topk.inl — the SQL vector-scan operator: `ORDER BY vec <op> '[...]' LIMIT k OFFSET o` for a batch of
// statements (TopKExec, src/sql/executor.rs:2239-2392; sort-key arithmetic :169-212; projected value
// src/sql/predicate.rs:1634-1688).  SURVEY.md §8(f) rank 3, BASELINE.json config 5's "SQL ORDER BY distance LIMIT 10
// batch path".  Included by turdb_cuda.cu.
//
// The reference streams the table in primary-key order through a (limit+offset)-row max-heap: the first K rows are
// pushed and stably sorted worst-first, every later row replaces the root iff its key is STRICTLY smaller (sift-down
// preferring the left child on ties), and a final stable ascending sort yields rows [offset, offset+limit).  Which of
// several equal keys survive, and in which order they are returned, follows from that procedure alone — so it is
// REPLAYED here, exactly, over a SUPERSET of the rows the reference would ever have pushed:
//   exact scan : the certified tensor-core filter (exact_search.cuh) — every slice's arrivals, archived;
//   use_index  : the HNSW traversal's ef best rows ("TopKExec over the rows the index returns").
// Rows the reference would not have pushed are ignored by the replay itself (their key is not below the root), so
// feeding a superset in scan order reproduces the reference's heap state step by step.  Keys are its f64 values in its
// summation order.  A query whose candidate buffers overflowed is redone by sql_stream_scan_kernel, which IS the
// reference's loop over all rows.  Order among NULL (NaN) keys is unspecified, as Rust's sort_by is for them.

namespace turdb {

// ORDER BY key (executor.rs:169-212).  OP 0: sqrt(sum_f64(((row - literal) as f32 -> f64)^2)); OP 1: 1 - dot/(|l||r|)
// in f64, NULL (NaN) when a norm is zero.  mag_q = |literal| (OP 1).
template <int OP>
__device__ __forceinline__ double sql_key64(const float* __restrict__ q, const float* __restrict__ row, uint32_t dim,
                                            double mag_q) {
  if (OP == 0) {
    double s = 0.0;
    for (uint32_t i = 0; i < dim; ++i) {
      const double d = (double)__fsub_rn(row[i], q[i]);
      s = __dadd_rn(s, __dmul_rn(d, d));
    }
    return __dsqrt_rn(s);
  }
  double dot = 0.0, sl = 0.0;
  for (uint32_t i = 0; i < dim; ++i) {
    const double x = (double)row[i];
    dot = __dadd_rn(dot, __dmul_rn(x, (double)q[i]));
    sl = __dadd_rn(sl, __dmul_rn(x, x));
  }
  const double mag_l = __dsqrt_rn(sl);
  if (mag_l > 0.0 && mag_q > 0.0) return __dsub_rn(1.0, __ddiv_rn(dot, __dmul_rn(mag_l, mag_q)));
  return __longlong_as_double(0x7FF8000000000000ll);
}
__device__ __forceinline__ double sql_mag64(const float* __restrict__ q, uint32_t dim) {
  double s = 0.0;
  for (uint32_t i = 0; i < dim; ++i) {
    const double x = (double)q[i];
    s = __dadd_rn(s, __dmul_rn(x, x));
  }
  return __dsqrt_rn(s);
}

// Projected value of `vec <op> literal` (predicate.rs:1634-1688): f32, sequential, unfused; returned as f64.
// op 0: sqrt(sum (a-b)*(a-b)); 1: 1 - dot/(|a||b|) with each norm's own sqrt, NULL (NaN) on a zero norm; 2: +dot.
__device__ __forceinline__ double sql_projection(int op, const float* __restrict__ a, const float* __restrict__ b, uint32_t dim) {
  if (op == 0) {
    float s = 0.f;
    for (uint32_t i = 0; i < dim; ++i) {
      const float d = __fsub_rn(a[i], b[i]);
      s = __fadd_rn(s, __fmul_rn(d, d));
    }
    return (double)__fsqrt_rn(s);
  }
  float dot = 0.f;
  for (uint32_t i = 0; i < dim; ++i) dot = __fadd_rn(dot, __fmul_rn(a[i], b[i]));
  if (op == 2) return (double)dot;
  float n1 = 0.f, n2 = 0.f;
  for (uint32_t i = 0; i < dim; ++i) n1 = __fadd_rn(n1, __fmul_rn(a[i], a[i]));
  for (uint32_t i = 0; i < dim; ++i) n2 = __fadd_rn(n2, __fmul_rn(b[i], b[i]));
  n1 = __fsqrt_rn(n1);
  n2 = __fsqrt_rn(n2);
  if (n1 == 0.f || n2 == 0.f) return __longlong_as_double(0x7FF8000000000000ll);
  return (double)__fsub_rn(1.0f, __fdiv_rn(dot, __fmul_rn(n1, n2)));
}

// TopKExec's heap, driven by one warp.  hk/hid: the heap (K entries); tk/tid: scratch of the same size for the two
// stable sorts.  All state lives in shared memory; `len` is warp-uniform.
struct TopKReplay {
  double* hk;
  uint32_t* hid;
  double* tk;
  uint32_t* tid;
  uint32_t K, len;

  // total preorder used by the two sorts: numbers by value, NULL (NaN) keys after every number, equal among themselves
  static __device__ __forceinline__ bool lt(double a, double b) { return a < b || (b != b && a == a); }

  // stable sort of the heap array: DESC (worst first, executor.rs:2260-2275) or ascending (:2360-2370).  Rank sort:
  // position = elements strictly before + equivalent elements with a smaller index.
  __device__ __forceinline__ void stable_sort(bool desc, uint32_t lane) {
    __syncwarp();
    for (uint32_t i = lane; i < len; i += 32) {
      const double ki = hk[i];
      uint32_t pos = 0;
      for (uint32_t j = 0; j < len; ++j) {
        const double kj = hk[j];
        const bool before = desc ? lt(ki, kj) : lt(kj, ki);
        const bool after = desc ? lt(kj, ki) : lt(ki, kj);
        pos += (before || (!after && j < i)) ? 1u : 0u;
      }
      tk[pos] = ki;
      tid[pos] = hid[i];
    }
    __syncwarp();
    for (uint32_t i = lane; i < len; i += 32) {
      hk[i] = tk[i];
      hid[i] = tid[i];
    }
    __syncwarp();
  }

  // rows in scan order (ascending id); keys[i] belongs to ids[i].  n <= whatever the caller staged.
  __device__ __forceinline__ void feed(const uint32_t* ids, const double* keys, uint32_t n, uint32_t lane) {
    uint32_t c = 0;
    if (len < K) {  // executor.rs:2248-2259: push until the heap holds limit + offset rows
      const uint32_t take = min(K - len, n);
      for (uint32_t i = lane; i < take; i += 32) {
        hk[len + i] = keys[i];
        hid[len + i] = ids[i];
      }
      len += take;
      c = take;
      __syncwarp();
      if (len == K) stable_sort(true, lane);
    }
    for (uint32_t base = c; base < n; base += 32) {
      const uint32_t i = base + lane;
      const double kc = i < n ? keys[i] : 0.0;
      // the root only ever decreases: a row that fails now can never pass later
      uint32_t mask = __ballot_sync(kFullMask, i < n && kc < hk[0]);
      while (mask) {
        const uint32_t l = __ffs(mask) - 1;
        mask &= mask - 1;
        if (lane == 0) {
          const double k = keys[base + l];
          if (k < hk[0]) {  // strictly less replaces the root (executor.rs:2277-2290), then sift down (:2292-2357)
            hk[0] = k;
            hid[0] = ids[base + l];
            uint32_t p = 0;
            for (;;) {
              const uint32_t left = 2 * p + 1, right = 2 * p + 2;
              uint32_t largest = p;
              if (left < len && hk[left] > hk[largest]) largest = left;
              if (right < len && hk[right] > hk[largest]) largest = right;
              if (largest == p) break;
              const double tkk = hk[p];
              const uint32_t tii = hid[p];
              hk[p] = hk[largest];
              hid[p] = hid[largest];
              hk[largest] = tkk;
              hid[largest] = tii;
              p = largest;
            }
          }
        }
        __syncwarp();
      }
    }
  }
};

// rows [offset, offset + limit) of the finished heap -> outputs (one warp)
__device__ __forceinline__ void sql_emit(const DeviceIndex& ix, TopKReplay& h, const float* __restrict__ qv, uint32_t q,
                                         uint32_t limit, uint32_t offset, int proj_op, uint64_t* out_rows, double* out_keys,
                                         double* out_proj, uint32_t* out_counts, uint32_t lane) {
  h.stable_sort(false, lane);
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  const uint32_t have = h.len > offset ? min(h.len - offset, limit) : 0u;
  for (uint32_t i = lane; i < limit; i += 32) {
    const size_t o = (size_t)q * limit + i;
    if (i < have) {
      const uint32_t id = h.hid[offset + i];
      out_rows[o] = ix.row_ids[id];
      out_keys[o] = h.hk[offset + i];
      if (out_proj) out_proj[o] = sql_projection(proj_op, ix.arena + (size_t)id * ix.ds, qv, ix.dim);
    } else {
      out_rows[o] = 0xFFFFFFFFFFFFFFFFull;
      out_keys[o] = nan;
      if (out_proj) out_proj[o] = nan;
    }
  }
  if (lane == 0) out_counts[q] = have;
}

// One warp (= one CTA) per statement: candidate ids -> ascending (scan order) -> f64 keys -> replay -> output.
// smem: ids [n2] u32 | keys [n2] f64 | heap 2 x K x (f64 + u32)
template <int OP>
__global__ void __launch_bounds__(32) sql_replay_kernel(DeviceIndex ix, const float* __restrict__ queries, uint32_t nq,
                                                        uint32_t limit, uint32_t offset, const uint32_t* __restrict__ cand_ids,
                                                        const uint32_t* __restrict__ cand_counts, uint32_t cand_stride,
                                                        uint32_t n2max, const uint32_t* __restrict__ qflags, int proj_op,
                                                        uint64_t* out_rows, double* out_keys, double* out_proj,
                                                        uint32_t* out_counts) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t lane = threadIdx.x, q = blockIdx.x;
  if (q >= nq) return;
  if (qflags && qflags[q]) return;  // redone by the streaming scan
  const uint32_t K = limit + offset;
  double* keys = reinterpret_cast<double*>(smem_raw);
  TopKReplay h;
  h.hk = keys + n2max;
  h.tk = h.hk + K;
  uint32_t* ids = reinterpret_cast<uint32_t*>(h.tk + K);
  h.hid = ids + n2max;
  h.tid = h.hid + K;
  h.K = K;
  h.len = 0;
  const uint32_t cnt = min(cand_counts[q], min(cand_stride, n2max));
  uint32_t n2 = 1;
  while (n2 < cnt) n2 <<= 1;
  for (uint32_t i = lane; i < n2; i += 32) ids[i] = i < cnt ? cand_ids[(size_t)q * cand_stride + i] : 0xFFFFFFFFu;
  __syncwarp();
  for (uint32_t size = 2; size <= n2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = lane; i < n2; i += 32) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool asc = (i & size) == 0;
          const uint32_t a = ids[i], b = ids[j];
          if ((a < b) != asc) {
            ids[i] = b;
            ids[j] = a;
          }
        }
      }
      __syncwarp();
    }
  }
  const float* qv = queries + (size_t)q * ix.dim;
  const double mag_q = OP == 1 ? sql_mag64(qv, ix.dim) : 0.0;
  for (uint32_t i = lane; i < cnt; i += 32) keys[i] = sql_key64<OP>(qv, ix.arena + (size_t)ids[i] * ix.ds, ix.dim, mag_q);
  __syncwarp();
  if (K) h.feed(ids, keys, cnt, lane);
  sql_emit(ix, h, qv, q, limit, offset, proj_op, out_rows, out_keys, out_proj, out_counts, lane);
}

// The reference's loop itself, for the statements the filter could not serve (qflags[q] != 0; e.g. thousands of rows
// tying with the limit-th key, or data on which the BF16 error bound admits too many candidates): every row's f64 key
// in primary-key order through the same heap.  CTAs stride over the statements and skip the unflagged ones; always
// enqueued, exits at once when nothing is flagged.  smem: chunk ids [256] | chunk keys [256] | heap.
template <int OP>
__global__ void __launch_bounds__(256) sql_stream_scan_kernel(DeviceIndex ix, const float* __restrict__ queries, uint32_t nq,
                                                              uint32_t limit, uint32_t offset,
                                                              const uint32_t* __restrict__ qflags, int proj_op,
                                                              uint64_t* out_rows, double* out_keys, double* out_proj,
                                                              uint32_t* out_counts) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t K = limit + offset;
  double* ckeys = reinterpret_cast<double*>(smem_raw);
  TopKReplay h;
  h.hk = ckeys + 256;
  h.tk = h.hk + K;
  uint32_t* cids = reinterpret_cast<uint32_t*>(h.tk + K);
  h.hid = cids + 256;
  h.tid = h.hid + K;
  h.K = K;
  for (uint32_t q = blockIdx.x; q < nq; q += gridDim.x) {
    if (!qflags[q]) continue;
    const float* qv = queries + (size_t)q * ix.dim;
    const double mag_q = OP == 1 ? sql_mag64(qv, ix.dim) : 0.0;
    h.len = 0;
    __syncthreads();
    for (uint64_t base = 0; base < ix.n && K; base += 256) {
      const uint64_t r = base + tid;
      if (r < ix.n) {
        ckeys[tid] = sql_key64<OP>(qv, ix.arena + r * ix.ds, ix.dim, mag_q);
        cids[tid] = (uint32_t)r;
      }
      __syncthreads();
      if (warp == 0) h.feed(cids, ckeys, (uint32_t)min((uint64_t)256, (uint64_t)(ix.n - base)), lane);
      __syncthreads();
    }
    if (warp == 0) sql_emit(ix, h, qv, q, limit, offset, proj_op, out_rows, out_keys, out_proj, out_counts, lane);
    __syncthreads();
  }
}

__global__ void fill_empty_sql_kernel(uint64_t* rows, double* keys, double* proj, uint32_t* counts, uint32_t nq, uint32_t limit) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (uint64_t)nq * limit) {
    rows[i] = 0xFFFFFFFFFFFFFFFFull;
    keys[i] = __longlong_as_double(0x7FF8000000000000ll);
    if (proj) proj[i] = __longlong_as_double(0x7FF8000000000000ll);
  }
  if (i < nq) counts[i] = 0;
}

}  // namespace turdb

// defined in exact_abi.inl: the certified filter, archiving every arrival (ids) per query
static int32_t exact_filter_archive(turdb_cuda_index* idx, const float* d_queries, uint32_t nq, uint32_t K, uint8_t metric,
                                    uint32_t arch_cap, uint32_t* d_arch_cnt, uint32_t* d_arch_id, uint32_t* d_qflags,
                                    cudaStream_t stream);

extern "C" int32_t turdb_cuda_sql_topk_batch_device(turdb_cuda_index* idx, const float* d_queries, uint32_t query_dim,
                                                    uint32_t nq, uint32_t limit, uint32_t offset, uint8_t op,
                                                    uint8_t proj_op, int32_t use_index, uint32_t ef,
                                                    uint64_t* d_out_row_ids, double* d_out_keys, double* d_out_proj,
                                                    uint32_t* d_out_counts, void* stream_) {
  if (!idx) return fail(TURDB_ERR_INVALID_ARGUMENT, "idx is null");
  if (query_dim != idx->ix.dim)
    return fail(TURDB_ERR_DIMENSION_MISMATCH, "query dimension %u does not match index dimension %u", query_dim, idx->ix.dim);
  if (op > 1)
    return fail(TURDB_ERR_UNSUPPORTED, "ORDER BY <#> evaluates to NULL for every row in the reference (executor.rs:241); "
                                       "only <-> (0) and <=> (1) are sort keys");
  if (d_out_proj && proj_op > 2) return fail(TURDB_ERR_INVALID_ARGUMENT, "proj_op %u unknown (0 <->, 1 <=>, 2 <#>)", proj_op);
  if (nq == 0) return TURDB_OK;
  if (!d_queries || !d_out_counts || (limit && (!d_out_row_ids || !d_out_keys)))
    return fail(TURDB_ERR_INVALID_ARGUMENT, "null query/output pointer");
  const uint64_t K64 = (uint64_t)limit + offset;
  if (K64 > TURDB_SQL_MAX_LIMIT_PLUS_OFFSET)
    return fail(TURDB_ERR_UNSUPPORTED, "limit + offset = %llu > %u", (unsigned long long)K64, TURDB_SQL_MAX_LIMIT_PLUS_OFFSET);
  const uint32_t K = (uint32_t)K64;
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", idx->device);
  const uint64_t n = idx->ix.n;
  if (n == 0 || limit == 0) {
    uint64_t total = std::max<uint64_t>((uint64_t)nq * std::max(limit, 1u), nq);
    fill_empty_sql_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_out_row_ids, d_out_keys, d_out_proj, d_out_counts, nq, limit);
    CUDA_TRY(cudaGetLastError());
    return TURDB_OK;
  }

  // candidate ids per statement: [nq][stride]
  uint32_t stride;
  if (use_index) {
    if (ef < K) ef = K;
    if (ef > 2048) return fail(TURDB_ERR_UNSUPPORTED, "index-backed scan needs ef >= limit + offset = %u > 2048", K);
    stride = ef;  // TopKExec over every row the index's beam returns
  } else {
    // first slice (>= 2K rows, all archived) + the arrivals of ~log3(n) later slices (~1.1 K each without slack)
    uint32_t want = std::max(256u, 2 * K) + 12 * K + 512;
    stride = 1024;
    while (stride < want) stride <<= 1;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_ids = take((size_t)nq * stride * 4), o_cnt = take((size_t)nq * 4), o_flags = take((size_t)nq * 4),
               o_rows = take(use_index ? (size_t)nq * stride * 8 : 0), o_dist = take(use_index ? (size_t)nq * stride * 4 : 0);
  uint8_t* scr = nullptr;
  CUDA_TRY(cudaMallocFromPoolAsync(&scr, off, idx->pool, stream));
  uint32_t* c_ids = (uint32_t*)(scr + o_ids);
  uint32_t* c_cnt = (uint32_t*)(scr + o_cnt);
  uint32_t* c_flags = (uint32_t*)(scr + o_flags);
  int32_t rc;
  if (use_index) {
    cudaMemsetAsync(c_flags, 0, (size_t)nq * 4, stream);
    rc = turdb_cuda_search_batch_device(idx, d_queries, query_dim, nq, stride, ef, op, nullptr, (uint64_t*)(scr + o_rows), c_ids,
                                        (float*)(scr + o_dist), c_cnt, nullptr, stream);
  } else {
    rc = exact_filter_archive(idx, d_queries, nq, K, op, stride, c_cnt, c_ids, c_flags, stream);
  }
  if (rc != TURDB_OK) {
    cudaFreeAsync(scr, stream);
    return rc;
  }
  const size_t heap_bytes = (size_t)2 * K * 12 + 32;
  const size_t smem_replay = (size_t)stride * 12 + heap_bytes, smem_scan = (size_t)256 * 12 + heap_bytes;
  cudaError_t e = cudaSuccess;
  auto raise = [&](auto kern, size_t smem) {
    if (e == cudaSuccess && smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  };
  const uint32_t scan_grid = (uint32_t)std::min<uint64_t>(nq, 2ull * idx->num_sms);
  if (op == 0) {
    raise(sql_replay_kernel<0>, smem_replay);
    raise(sql_stream_scan_kernel<0>, smem_scan);
    if (e == cudaSuccess) {
      sql_replay_kernel<0><<<nq, 32, smem_replay, stream>>>(idx->ix, d_queries, nq, limit, offset, c_ids, c_cnt, stride, stride, c_flags,
                                                           proj_op, d_out_row_ids, d_out_keys, d_out_proj, d_out_counts);
      if (!use_index)
        sql_stream_scan_kernel<0><<<scan_grid, 256, smem_scan, stream>>>(idx->ix, d_queries, nq, limit, offset, c_flags, proj_op,
                                                                        d_out_row_ids, d_out_keys, d_out_proj, d_out_counts);
    }
  } else {
    raise(sql_replay_kernel<1>, smem_replay);
    raise(sql_stream_scan_kernel<1>, smem_scan);
    if (e == cudaSuccess) {
      sql_replay_kernel<1><<<nq, 32, smem_replay, stream>>>(idx->ix, d_queries, nq, limit, offset, c_ids, c_cnt, stride, stride, c_flags,
                                                           proj_op, d_out_row_ids, d_out_keys, d_out_proj, d_out_counts);
      if (!use_index)
        sql_stream_scan_kernel<1><<<scan_grid, 256, smem_scan, stream>>>(idx->ix, d_queries, nq, limit, offset, c_flags, proj_op,
                                                                        d_out_row_ids, d_out_keys, d_out_proj, d_out_counts);
    }
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFreeAsync(scr, stream);
  if (e != cudaSuccess) return fail(TURDB_ERR_CUDA, "sql_topk launch failed: %s", cudaGetErrorString(e));
  return TURDB_OK;
}

extern "C" int32_t turdb_cuda_sql_topk_batch(turdb_cuda_index* idx, const float* queries, uint32_t query_dim, uint32_t nq,
                                             uint32_t limit, uint32_t offset, uint8_t op, uint8_t proj_op, int32_t use_index,
                                             uint32_t ef, uint64_t* out_row_ids, double* out_keys, double* out_proj,
                                             uint32_t* out_counts) {
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
  const size_t off_rows = (qbytes + 255) & ~(size_t)255, off_keys = off_rows + nq * ll * 8, off_proj = off_keys + nq * ll * 8,
               off_counts = off_proj + nq * ll * 8, total = off_counts + (size_t)nq * 4;
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
  int32_t rc = turdb_cuda_sql_topk_batch_device(idx, (const float*)slab, query_dim, nq, limit, offset, op, proj_op, use_index, ef,
                                                (uint64_t*)(slab + off_rows), (double*)(slab + off_keys),
                                                out_proj ? (double*)(slab + off_proj) : nullptr, (uint32_t*)(slab + off_counts),
                                                stream);
  if (rc != TURDB_OK) {
    cleanup();
    return rc;
  }
  e = cudaMemcpyAsync(out_counts, slab + off_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess && limit) {
    e = cudaMemcpyAsync(out_row_ids, slab + off_rows, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_keys, slab + off_keys, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && out_proj) e = cudaMemcpyAsync(out_proj, slab + off_proj, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost, stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cleanup();
  if (e != cudaSuccess) return fail(TURDB_ERR_CUDA, "sql_topk failed: %s", cudaGetErrorString(e));
  return TURDB_OK;
}
