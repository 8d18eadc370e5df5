// exact_abi.inl — C-ABI entry points of the exact path and the shard merge (included by turdb_cuda.cu)

extern "C" int32_t turdb_cuda_merge_topk_device(int32_t device, const uint64_t* d_gathered_row_ids,
                                                const float* d_gathered_dist, const uint32_t* d_gathered_counts,
                                                uint32_t n_shards, uint32_t nq, uint32_t k, uint64_t* d_out_row_ids,
                                                float* d_out_dist, uint32_t* d_out_counts, void* stream_) {
  if (n_shards == 0 || n_shards > 32) return fail(TURDB_ERR_INVALID_ARGUMENT, "n_shards must be 1..32");
  if (nq == 0) return TURDB_OK;
  if (!d_gathered_row_ids || !d_gathered_dist || !d_gathered_counts || !d_out_counts || (k && (!d_out_row_ids || !d_out_dist)))
    return fail(TURDB_ERR_INVALID_ARGUMENT, "null pointer");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const uint32_t threads = 128;
  const uint32_t blocks = (uint32_t)(((uint64_t)nq * 32 + threads - 1) / threads);
  merge_topk_kernel<<<blocks, threads, 0, stream>>>((const uint8_t*)d_gathered_row_ids, (const uint8_t*)d_gathered_dist,
                                                    (const uint8_t*)d_gathered_counts, (size_t)nq * k * 8, (size_t)nq * k * 4,
                                                    (size_t)nq * 4, n_shards, nq, k, d_out_row_ids, d_out_dist, d_out_counts);
  CUDA_TRY(cudaGetLastError());
  return TURDB_OK;
}

// The same merge over ONE packed block per shard — row ids [nq][k] u64 | distances [nq][k] f32 | counts [nq] u32, blocks
// shard_stride_bytes apart (a multiple of 8) — i.e. over the output of a single all-gather of each rank's block.
extern "C" int32_t turdb_cuda_merge_topk_packed_device(int32_t device, const void* d_gathered, uint64_t shard_stride_bytes,
                                                       uint32_t n_shards, uint32_t nq, uint32_t k, uint64_t* d_out_row_ids,
                                                       float* d_out_dist, uint32_t* d_out_counts, void* stream_) {
  if (n_shards == 0 || n_shards > 32) return fail(TURDB_ERR_INVALID_ARGUMENT, "n_shards must be 1..32");
  if (nq == 0) return TURDB_OK;
  const uint64_t need = (uint64_t)nq * k * 12 + (uint64_t)nq * 4;
  if (!d_gathered || !d_out_counts || (k && (!d_out_row_ids || !d_out_dist))) return fail(TURDB_ERR_INVALID_ARGUMENT, "null pointer");
  if (shard_stride_bytes < need || (shard_stride_bytes & 7)) return fail(TURDB_ERR_INVALID_ARGUMENT, "shard stride %llu: need a multiple of 8, >= %llu", (unsigned long long)shard_stride_bytes, (unsigned long long)need);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const uint32_t threads = 128;
  const uint32_t blocks = (uint32_t)(((uint64_t)nq * 32 + threads - 1) / threads);
  const uint8_t* b = (const uint8_t*)d_gathered;
  merge_topk_kernel<<<blocks, threads, 0, stream>>>(b, b + (size_t)nq * k * 8, b + (size_t)nq * k * 12, shard_stride_bytes,
                                                    shard_stride_bytes, shard_stride_bytes, n_shards, nq, k, d_out_row_ids,
                                                    d_out_dist, d_out_counts);
  CUDA_TRY(cudaGetLastError());
  return TURDB_OK;
}

// One host thread, one sub-index per GPU (turdb_cuda_shards_search_batch).  Every shard searches
// the same replicated query batch on its own device and stream; the per-shard top-k lists are copied device to
// device into shard 0's gather buffer ([n_shards][nq][k], the layout the NCCL all-gather of the multi-process
// path produces) and merged there by the same merge_topk_kernel.  The two paths return identical results; this one
// is for a single-process host (the Rust/C++ caller), the torch.distributed one for one-process-per-GPU serving.
extern "C" int32_t turdb_cuda_shards_search_batch(turdb_cuda_index* const* shards, uint32_t n_shards, const float* queries,
                                                  uint32_t query_dim, uint32_t nq, uint32_t k, uint32_t ef, uint8_t metric,
                                                  uint64_t* out_row_ids, float* out_dist, uint32_t* out_counts) {
  if (!shards || n_shards == 0 || n_shards > 32) return fail(TURDB_ERR_INVALID_ARGUMENT, "n_shards must be 1..32");
  for (uint32_t s = 0; s < n_shards; ++s) {
    if (!shards[s]) return fail(TURDB_ERR_INVALID_ARGUMENT, "shard %u is null", s);
    if (query_dim != shards[s]->ix.dim)
      return fail(TURDB_ERR_DIMENSION_MISMATCH, "query dimension %u does not match index dimension %u", query_dim, shards[s]->ix.dim);
  }
  if (nq == 0) return TURDB_OK;
  if (k == 0 || !queries || !out_row_ids || !out_dist || !out_counts) return fail(TURDB_ERR_INVALID_ARGUMENT, "null pointer or k == 0");
  struct Lane {
    cudaStream_t stream = nullptr;
    uint8_t* slab = nullptr;
    cudaEvent_t done = nullptr;
  };
  std::vector<Lane> lanes(n_shards);
  const size_t qbytes = (size_t)nq * query_dim * 4;
  const size_t off_rows = (qbytes + 255) & ~(size_t)255, off_dist = off_rows + (size_t)nq * k * 8,
               off_counts = off_dist + (size_t)nq * k * 4, local_total = off_counts + (size_t)nq * 4;
  // shard 0 additionally holds the gathered lists and the merged result
  const size_t g_rows = (local_total + 255) & ~(size_t)255, g_dist = g_rows + (size_t)n_shards * nq * k * 8,
               g_cnt = g_dist + (size_t)n_shards * nq * k * 4, m_rows = (g_cnt + (size_t)n_shards * nq * 4 + 255) & ~(size_t)255,
               m_dist = m_rows + (size_t)nq * k * 8, m_cnt = m_dist + (size_t)nq * k * 4, total0 = m_cnt + (size_t)nq * 4;
  int32_t rc = TURDB_OK;
  std::string err;
  auto cleanup = [&]() {
    for (uint32_t s = 0; s < n_shards; ++s) {
      DeviceGuard g(shards[s]->device);
      if (lanes[s].slab) cudaFreeAsync(lanes[s].slab, lanes[s].stream);
      if (lanes[s].stream) {
        cudaStreamSynchronize(lanes[s].stream);
        cudaStreamDestroy(lanes[s].stream);
      }
      if (lanes[s].done) cudaEventDestroy(lanes[s].done);
    }
  };
  for (uint32_t s = 0; s < n_shards && rc == TURDB_OK; ++s) {
    turdb_cuda_index* idx = shards[s];
    DeviceGuard g(idx->device);
    cudaError_t e = cudaStreamCreateWithFlags(&lanes[s].stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&lanes[s].done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&lanes[s].slab, s == 0 ? total0 : local_total, idx->pool, lanes[s].stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lanes[s].slab, queries, qbytes, cudaMemcpyHostToDevice, lanes[s].stream);
    if (e != cudaSuccess) {
      rc = fail(TURDB_ERR_CUDA, "shard %u setup failed: %s", s, cudaGetErrorString(e));
      break;
    }
    rc = turdb_cuda_search_batch_device(idx, (const float*)lanes[s].slab, query_dim, nq, k, ef, metric, nullptr,
                                        (uint64_t*)(lanes[s].slab + off_rows), nullptr, (float*)(lanes[s].slab + off_dist),
                                        (uint32_t*)(lanes[s].slab + off_counts), nullptr, lanes[s].stream);
  }
  if (rc != TURDB_OK) {
    err = g_last_error;
    cleanup();
    g_last_error = err;
    return rc;
  }
  // gather into shard 0 (peer copies ordered after each shard's search AND after shard 0's slab exists: it is a
  // stream-ordered allocation on lanes[0].stream), then merge on shard 0's stream
  cudaError_t e = cudaSuccess;
  cudaEvent_t slab0_ready = nullptr;
  {
    DeviceGuard g(shards[0]->device);
    e = cudaEventCreateWithFlags(&slab0_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(slab0_ready, lanes[0].stream);
  }
  for (uint32_t s = 0; s < n_shards && e == cudaSuccess; ++s) {
    DeviceGuard g(shards[s]->device);
    uint8_t* dst = lanes[0].slab;
    if (s != 0) e = cudaStreamWaitEvent(lanes[s].stream, slab0_ready, 0);
    if (e != cudaSuccess) break;
    e = cudaMemcpyPeerAsync(dst + g_rows + (size_t)s * nq * k * 8, shards[0]->device, lanes[s].slab + off_rows, shards[s]->device,
                            (size_t)nq * k * 8, lanes[s].stream);
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(dst + g_dist + (size_t)s * nq * k * 4, shards[0]->device, lanes[s].slab + off_dist, shards[s]->device,
                              (size_t)nq * k * 4, lanes[s].stream);
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(dst + g_cnt + (size_t)s * nq * 4, shards[0]->device, lanes[s].slab + off_counts, shards[s]->device,
                              (size_t)nq * 4, lanes[s].stream);
    if (e == cudaSuccess) e = cudaEventRecord(lanes[s].done, lanes[s].stream);
  }
  if (e == cudaSuccess) {
    DeviceGuard g(shards[0]->device);
    for (uint32_t s = 0; s < n_shards && e == cudaSuccess; ++s) e = cudaStreamWaitEvent(lanes[0].stream, lanes[s].done, 0);
    if (e == cudaSuccess) {
      uint8_t* b = lanes[0].slab;
      rc = turdb_cuda_merge_topk_device(shards[0]->device, (const uint64_t*)(b + g_rows), (const float*)(b + g_dist),
                                        (const uint32_t*)(b + g_cnt), n_shards, nq, k, (uint64_t*)(b + m_rows), (float*)(b + m_dist),
                                        (uint32_t*)(b + m_cnt), lanes[0].stream);
      if (rc == TURDB_OK) {
        e = cudaMemcpyAsync(out_row_ids, b + m_rows, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, lanes[0].stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, b + m_dist, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, lanes[0].stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_counts, b + m_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, lanes[0].stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(lanes[0].stream);
      }
    }
  }
  if (rc == TURDB_OK && e != cudaSuccess) rc = fail(TURDB_ERR_CUDA, "sharded search failed: %s", cudaGetErrorString(e));
  err = g_last_error;
  if (slab0_ready) {
    DeviceGuard g(shards[0]->device);
    cudaStreamSynchronize(lanes[0].stream);
    cudaEventDestroy(slab0_ready);
  }
  cleanup();
  g_last_error = err;
  return rc;
}

// 1: problems of up to three K-chunks (dim <= 192; L2: dim <= 189) run 128-vector tiles over four accumulators
#ifndef TURDB_EXACT_SHORT_K_TILE128
#define TURDB_EXACT_SHORT_K_TILE128 0
#endif

// ---- TMA descriptors: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// rows x kp BF16, row-major; box = 64 (K) x box_rows (128 queries / kTileN vectors), 128-byte swizzle, out-of-range rows
// read as zero
static bool make_bf16_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t kp, uint32_t box_rows, int fp16) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {kp, rows};
  cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
  cuuint32_t box[2] = {kChunkK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void exact_init_kernel(float* thresh, uint32_t* cand_cnt, uint32_t* kept, uint32_t* qflags, uint32_t* arch_cnt,
                                  uint32_t nq, uint32_t first_cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) {
    thresh[i] = -INFINITY;
    cand_cnt[i] = first_cnt;  // the first slice is written densely (ExactArgs::dense): its column count
    kept[i] = 0;
    qflags[i] = 0;
    if (arch_cnt) arch_cnt[i] = 0;
  }
}

// What the filter leaves behind for its caller (device pointers into one scratch allocation, freed by the caller).
struct ExactFilterResult {
  uint8_t* scr = nullptr;
  uint32_t* cand_cnt = nullptr;  // [nq]       entries kept per query (sorted best first)
  uint32_t* cand_id = nullptr;   // [nq][cap]
  uint32_t cap = 0;
};

// The certified tensor-core filter (exact_search.cuh): BF16 scores over geometrically growing slices of the corpus, per
// query threshold = kprime-th best filter key so far minus twice the score-error bound e(q) (query_slack_kernel) — a
// rejected row's exact key lies below the exact keys of kprime admitted rows.  With an archive every arrival's id is
// also appended to [nq][arch_cap] (the SQL operator replays them).  qflags[q] != 0: a buffer of query q overflowed.
static int32_t exact_filter_run(turdb_cuda_index* idx, const float* d_queries, uint32_t nq, uint32_t kprime, uint8_t metric,
                                uint32_t cap, uint32_t first_rows, uint32_t arch_cap, uint32_t* d_arch_cnt,
                                uint32_t* d_arch_id, uint32_t* d_qflags, ExactFilterResult* out, cudaStream_t stream) {
  const uint64_t n = idx->ix.n;
  const uint32_t dim = idx->ix.dim, ds = idx->ix.ds;
  // 16-bit copies of the arena, built once per index: 0 raw rows (IP), 1 rows scaled by 1/|x| (cosine), 2 raw rows plus
  // three columns carrying -|x|^2/2 (L2: the contraction itself yields the ranking key; an earlier version added the bias
  // in the epilogue — 16 broadcast LDS.128 + 64 FADD per warp and tile, 1.5x slower than IP at 128-d).
  const int copy = metric == kCosine ? 1 : (metric == kL2 ? 2 : 0);
  const uint32_t kp = (dim + (copy == 2 ? 3 : 0) + kChunkK - 1) / kChunkK * kChunkK;
  const uint32_t k_chunks = kp / kChunkK;
  if (k_chunks > 32)
    return fail(TURDB_ERR_UNSUPPORTED, "the exact path supports dim <= %u for this metric (dim %u)", copy == 2 ? 2045u : 2048u, dim);
  // up to 512 dims the query block (128 x K BF16) stays resident in shared memory; above, its K chunks are streamed
  // with the vector tile's (twice the TMA traffic per tile, but any K fits)
  const uint32_t stream_a = k_chunks > 8 ? 1u : 0u;
  // vectors per MMA tile: 256 (two accumulators) — or, at short K, 128 with four accumulators and two alternating sets of
  // epilogue warps (exact_search.cuh); TURDB_EXACT_TILE_N=128/256 overrides the choice (measurement)
  uint32_t tile_n = TURDB_EXACT_SHORT_K_TILE128 && k_chunks <= 3 ? 128u : 256u;
  if (const char* ev = getenv("TURDB_EXACT_TILE_N")) tile_n = (atoi(ev) == 128 && !stream_a) ? 128u : 256u;

  // each copy comes with the maxima of its rounding-error and row norms (the inputs of the filter's error bound)
  {
    std::lock_guard<std::mutex> lk(idx->mu);
    uint16_t*& dst = copy == 1 ? idx->d_arena_bf16n : (copy == 2 ? idx->d_arena_bf16l2 : idx->d_arena_bf16);
    if (!dst) {
      if (!idx->d_bf16_max2) {
        CUDA_TRY(cudaMalloc(&idx->d_bf16_max2, 8 * 4));  // [3 copies][2 maxima] + max |value| + max |x|^2 of the arena
        CUDA_TRY(cudaMemsetAsync(idx->d_bf16_max2, 0, 8 * 4, stream));
      }
      const bool force_bf16 = getenv("TURDB_EXACT_FORCE_BF16") != nullptr;
      if (copy != 1) {  // raw rows: FP16 only if every value fits comfortably (the cosine copy is unit-length rows: always FP16)
        max_abs_kernel<<<(unsigned)((n * dim + 255) / 256), 256, 0, stream>>>(idx->d_arena, dim, ds, n, idx->d_bf16_max2 + 6);
        max_nonneg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(idx->d_norm2, n, idx->d_bf16_max2 + 7);
        uint32_t mbits[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(mbits, idx->d_bf16_max2 + 6, 8, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        float mabs, n2max;
        memcpy(&mabs, &mbits[0], 4);
        memcpy(&n2max, &mbits[1], 4);
        int f16 = (mabs <= 16384.0f && !force_bf16) ? 1 : 0;
        if (copy == 2) {
          // FP16 bias columns hold t = -|x|^2 / (2 S) with |t| <= 16384 and the query side carries S <= 32768 (both exact
          // powers of two); beyond that range the copy is BF16 (FP32 range, S = 1)
          float S = 1.f;
          while (f16 && 0.5f * n2max / S > 16384.0f && S < 65536.f) S *= 2.f;
          if (S > 32768.f) {
            f16 = 0;
            S = 1.f;
          }
          idx->l2_scale = f16 ? S : 1.f;
        }
        idx->half_fp16[copy] = f16;
      } else if (force_bf16) {
        idx->half_fp16[1] = 0;
      }
      CUDA_TRY(cudaMalloc(&dst, (size_t)n * kp * 2));
      const uint64_t total = n * kp;
      to_half_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(idx->d_arena, dim, ds, kp, n,
                                                                            copy != 0 ? idx->d_norm2 : nullptr, idx->half_fp16[copy], dst,
                                                                            copy == 2 ? 1 : 0, idx->l2_scale);
      bf16_rowerr_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(idx->d_arena, dim, ds, kp, n,
                                                                                 copy == 1 ? idx->d_norm2 : nullptr, dst,
                                                                                 idx->half_fp16[copy], idx->d_bf16_max2 + 2 * copy);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(stream));
      idx->device_bytes += (size_t)n * kp * 2;
    }
  }
  const uint16_t* d_xb = copy == 1 ? idx->d_arena_bf16n : (copy == 2 ? idx->d_arena_bf16l2 : idx->d_arena_bf16);
  const int fp16 = idx->half_fp16[copy];

  // scratch: Qb | thresh | cand_cnt | kept | slack | cand_id | cand_key
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_qb = take((size_t)nq * kp * 2), o_th = take((size_t)nq * 4),
               o_cnt = take((size_t)nq * 4), o_kept = take((size_t)nq * 4), o_slack = take((size_t)nq * 4),
               o_id = take((size_t)nq * cap * 4), o_key = take((size_t)nq * cap * 4);
  uint8_t* scr = nullptr;
  CUDA_TRY(cudaMallocFromPoolAsync(&scr, off, idx->pool, stream));
  uint16_t* d_qb = (uint16_t*)(scr + o_qb);
  float* d_th = (float*)(scr + o_th);
  uint32_t* d_cnt = (uint32_t*)(scr + o_cnt);
  uint32_t* d_kept = (uint32_t*)(scr + o_kept);
  float* d_slack = (float*)(scr + o_slack);
  uint32_t* d_cid = (uint32_t*)(scr + o_id);
  float* d_ckey = (float*)(scr + o_key);
  auto bail = [&](int32_t rc) {
    cudaFreeAsync(scr, stream);
    return rc;
  };

  {
    const uint64_t total = (uint64_t)nq * kp;
    to_half_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_queries, dim, dim, kp, nq, nullptr, fp16, d_qb,
                                                                          copy == 2 ? 2 : 0, idx->l2_scale);
    const uint32_t first_tiles = std::max(1u, std::min(first_rows, cap / 2) / tile_n);
    const uint32_t first_cnt = (uint32_t)std::min<uint64_t>(n, (uint64_t)first_tiles * tile_n);
    exact_init_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(d_th, d_cnt, d_kept, d_qflags, d_arch_cnt, nq, first_cnt);
    // diagnostic only (measuring what the certificate costs): TURDB_EXACT_SLACK_SCALE=0 turns the slack band off, which
    // makes the filter the uncertified heuristic of round 1
    float slack_scale = 1.0f;
    if (const char* ev = getenv("TURDB_EXACT_SLACK_SCALE")) slack_scale = (float)atof(ev);
    query_slack_kernel<<<(unsigned)(((uint64_t)nq * 32 + 255) / 256), 256, 0, stream>>>(d_queries, dim, kp, nq, d_qb, fp16,
                                                                                         idx->d_bf16_max2 + 2 * copy, metric,
                                                                                         slack_scale, d_slack,
                                                                                         copy == 2 ? idx->l2_scale : 0.f);
  }
  // Two-CTA form (tcgen05 cta_group::2, clusters of 2): each CTA of a pair stages half of every vector tile, so the L2 -> SM
  // traffic and the shared-memory fill per flop halve (exact_search.cuh).  TURDB_EXACT_PAIR=0/1 overrides the default.
  // Measured, 10k queries x 1M vectors (profiles/r02_exact_probe_final.json): with the CTA-scope remote arrive the pair form
  // wins at every K that was tried — 128-d 2.96 against 3.05 ms, 256-d 4.06 against 4.43, 384-d 5.47 against 6.02, 512-d
  // 7.1 against 10.2 (a resident 128 KB query block leaves the one-CTA form two stages), 768-d 10.2 against 11.1.
  // One-chunk problems (dim <= 64) keep the one-CTA form (not measured).
  int pair = k_chunks >= 2 ? 1 : 0;
  if (const char* ev = getenv("TURDB_EXACT_PAIR")) pair = atoi(ev) != 0;
  if (idx->num_sms < 2) pair = 0;
  cudaError_t e = cudaSuccess;
  {
    static std::mutex attr_mu;  // cudaFuncSetAttribute is process-wide state
    std::lock_guard<std::mutex> lk(attr_mu);
    auto set_smem = [&](auto kern, size_t bytes) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    };
    set_smem(exact_gemm_filter_kernel<false, 256>, (size_t)idx->max_smem_optin);
    set_smem(exact_gemm_filter_kernel<true, 256>, (size_t)idx->max_smem_optin);
    set_smem(exact_gemm_filter_kernel<false, 128>, (size_t)idx->max_smem_optin);
    set_smem(exact_gemm_filter_pair_kernel<false, 256>, (size_t)idx->max_smem_optin);
    set_smem(exact_gemm_filter_pair_kernel<true, 256>, (size_t)idx->max_smem_optin);
    set_smem(exact_gemm_filter_pair_kernel<false, 128>, (size_t)idx->max_smem_optin);
  }
  if (e != cudaSuccess) return bail(fail(TURDB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)));

  // shared-memory plan of one form: stage size, pipeline depth, total dynamic shared memory (0 stages = does not fit)
  const size_t fixed_smem = exact_fixed_smem(k_chunks, stream_a != 0);
  size_t stage_bytes = 0, gemm_smem = 0;
  uint32_t n_stages = 0;
  auto plan = [&](int as_pair) {
    stage_bytes = exact_stage_bytes(stream_a != 0, as_pair != 0, tile_n);
    n_stages = (size_t)idx->max_smem_optin < fixed_smem + 2 * stage_bytes
                   ? 0u
                   : (uint32_t)std::min<size_t>(kMaxStages, ((size_t)idx->max_smem_optin - fixed_smem) / stage_bytes);
    gemm_smem = fixed_smem + (size_t)n_stages * stage_bytes;
  };
  // workers: CTAs, or CTA pairs (as many clusters of 2 as the device keeps resident at this shared-memory size)
  uint32_t n_workers = (uint32_t)idx->num_sms;
  if (pair) {
    plan(1);
    int max_clusters = 0;
    if (n_stages) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(idx->num_sms & ~1), 1, 1);
      cfg.blockDim = dim3(kExactThreads, 1, 1);
      cfg.dynamicSmemBytes = gemm_smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaError_t oe = stream_a       ? cudaOccupancyMaxActiveClusters(&max_clusters, exact_gemm_filter_pair_kernel<true, 256>, &cfg)
                       : tile_n == 128 ? cudaOccupancyMaxActiveClusters(&max_clusters, exact_gemm_filter_pair_kernel<false, 128>, &cfg)
                                       : cudaOccupancyMaxActiveClusters(&max_clusters, exact_gemm_filter_pair_kernel<false, 256>, &cfg);
      if (oe != cudaSuccess) {
        cudaGetLastError();
        max_clusters = 0;
      }
    }
    if (max_clusters > 0) n_workers = (uint32_t)std::min<int>(max_clusters, idx->num_sms / 2);
    else pair = 0;  // no cluster of 2 fits here: one CTA per tile
  }
  if (!pair) plan(0);
  if (n_stages < 2) return bail(fail(TURDB_ERR_UNSUPPORTED, "not enough shared memory for the exact path at dim %u", dim));
  if (getenv("TURDB_EXACT_VERBOSE"))
    fprintf(stderr, "[turdb exact] form=%s tile_n=%u workers=%u stages=%u smem=%zu k_chunks=%u stream_a=%u fp16=%d\n", pair ? "two-CTA" : "one-CTA",
            tile_n, n_workers, n_stages, gemm_smem, k_chunks, stream_a, fp16);
  CUtensorMap map_q, map_x;
  if (!make_bf16_map(&map_q, d_qb, nq, kp, kTileM, fp16) || !make_bf16_map(&map_x, d_xb, n, kp, pair ? tile_n / 2 : tile_n, fp16))
    return bail(fail(TURDB_ERR_CUDA, "cuTensorMapEncodeTiled failed"));

  // Slice growth g: with the threshold frozen at the kprime-th best of the m rows seen so far, the next (g - 1) m rows bring
  // about (g - 1) kprime arrivals per query.  Measured at 1M x 384, kprime 40 (profiles/r02_exact_growth.json): g = 4
  // 8.4 ms, 8: 9.0, 13: 9.6, 16: 10.2 — arrivals cost more than passes.  TURDB_EXACT_GROWTH overrides it (measurement).
  // With the radix-select threshold kernel a pass is cheap: g = 2: 6.60 ms, g = 3: 6.52 ms (kprime = k = 10).
  uint32_t growth = 3, diag = 0;
  if (const char* ev = getenv("TURDB_EXACT_GROWTH")) growth = (uint32_t)std::max(2, atoi(ev));
  if (const char* ev = getenv("TURDB_EXACT_DIAG")) diag = (uint32_t)atoi(ev);
  const uint32_t n_tiles = (uint32_t)((n + tile_n - 1) / tile_n);
  const uint32_t q_rows = pair ? 2 * kTileM : kTileM;  // queries per work item
  const uint32_t n_qblocks = (nq + q_rows - 1) / q_rows;
  // first slice: every column becomes a candidate (threshold -inf), so it must fit the buffer
  uint32_t lo = 0, span = std::max(1u, std::min(first_rows, cap / 2) / tile_n);
  while (lo < n_tiles) {
    const uint32_t hi = std::min(n_tiles, lo + span);
    ExactArgs a{};
    a.n_vec = (uint32_t)n;
    a.nq = nq;
    a.k_chunks = k_chunks;
    a.n_stages = n_stages;
    a.stream_a = stream_a;
    a.fp16 = (uint32_t)fp16;
    a.tile_lo = lo;
    a.tile_hi = hi;
    const uint32_t tiles = hi - lo;
    // tiles per work item: the resident query block is reloaded once per item (a ~1.5 us bubble), so short-K tiles get
    // longer items; never fewer than ~4 items per worker and pass
    const uint32_t tpi_cap = 16u * std::max(1u, 6u / k_chunks) * (256u / tile_n);
    uint32_t tpi = (uint32_t)std::min<uint64_t>(tpi_cap, std::max<uint64_t>(1, ((uint64_t)tiles * n_qblocks) / (4ull * n_workers)));
    a.tiles_per_item = tpi;
    a.n_qblocks = n_qblocks;
    a.n_items = n_qblocks * ((tiles + tpi - 1) / tpi);
    a.thresh = d_th;
    a.cand_cnt = d_cnt;
    a.cand_id = d_cid;
    a.cand_key = d_ckey;
    a.cap = cap;
    a.qflags = d_qflags;
    a.dbg = idx->d_dbg;
    a.diag = diag;
    a.dense = lo == 0 ? 1u : 0u;
    const uint32_t grid = std::min<uint32_t>(a.n_items, n_workers) * (pair ? 2u : 1u);
    if (pair) {  // __cluster_dims__(2, 1, 1) on the kernel: the grid is a multiple of 2
      if (stream_a) exact_gemm_filter_pair_kernel<true, 256><<<grid, kExactThreads, gemm_smem, stream>>>(map_q, map_x, a);
      else if (tile_n == 128) exact_gemm_filter_pair_kernel<false, 128><<<grid, kExactThreads, gemm_smem, stream>>>(map_q, map_x, a);
      else exact_gemm_filter_pair_kernel<false, 256><<<grid, kExactThreads, gemm_smem, stream>>>(map_q, map_x, a);
    } else {
      if (stream_a) exact_gemm_filter_kernel<true, 256><<<grid, kExactThreads, gemm_smem, stream>>>(map_q, map_x, a);
      else if (tile_n == 128) exact_gemm_filter_kernel<false, 128><<<grid, kExactThreads, gemm_smem, stream>>>(map_q, map_x, a);
      else exact_gemm_filter_kernel<false, 256><<<grid, kExactThreads, gemm_smem, stream>>>(map_q, map_x, a);
    }
    exact_threshold_kernel<<<(nq + kThreshWarps - 1) / kThreshWarps, 32 * kThreshWarps, 0, stream>>>(nq, kprime, cap, d_cnt, d_cid, d_ckey, d_th, d_slack, d_kept, d_qflags,
                                                         d_arch_cnt, d_arch_id, arch_cap);
    e = cudaGetLastError();
    if (e != cudaSuccess) return bail(fail(TURDB_ERR_CUDA, "exact pass launch failed: %s", cudaGetErrorString(e)));
    lo = hi;
    span = hi * (growth - 1);  // the next slice is (growth - 1) x everything seen so far: ~ln(growth) * kprime arrivals per
                               // query plus the slack band
  }
  out->scr = scr;
  out->cand_cnt = d_cnt;
  out->cand_id = d_cid;
  out->cap = cap;
  return TURDB_OK;
}

static int32_t exact_filter_archive(turdb_cuda_index* idx, const float* d_queries, uint32_t nq, uint32_t K, uint8_t metric,
                                    uint32_t arch_cap, uint32_t* d_arch_cnt, uint32_t* d_arch_id, uint32_t* d_qflags,
                                    cudaStream_t stream) {
  const uint32_t kprime = (uint32_t)std::min<uint64_t>(std::max(K, 1u), std::max<uint64_t>(idx->ix.n, 1));
  uint32_t cap = 2048;
  while (cap < 8 * kprime) cap <<= 1;
  ExactFilterResult r;
  int32_t rc = exact_filter_run(idx, d_queries, nq, kprime, metric, cap, std::max(256u, 2 * K), arch_cap, d_arch_cnt, d_arch_id,
                                d_qflags, &r, stream);
  if (rc != TURDB_OK) return rc;
  CUDA_TRY(cudaFreeAsync(r.scr, stream));
  return TURDB_OK;
}

extern "C" int32_t turdb_cuda_bruteforce_topk_device(turdb_cuda_index* idx, const float* d_queries, uint32_t query_dim,
                                                     uint32_t nq, uint32_t k, uint8_t metric, uint32_t rerank_factor,
                                                     uint64_t* d_out_row_ids, uint32_t* d_out_node_ids,
                                                     float* d_out_dist, uint32_t* d_out_counts, void* stream_) {
  if (!idx) return fail(TURDB_ERR_INVALID_ARGUMENT, "idx is null");
  if (query_dim != idx->ix.dim)
    return fail(TURDB_ERR_DIMENSION_MISMATCH, "query dimension %u does not match index dimension %u", query_dim, idx->ix.dim);
  if (metric > 2) return fail(TURDB_ERR_INVALID_ARGUMENT, "metric %u unknown", metric);
  if (nq == 0) return TURDB_OK;
  if (!d_queries || !d_out_counts || (k && (!d_out_row_ids || !d_out_dist)))
    return fail(TURDB_ERR_INVALID_ARGUMENT, "null query/output pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", idx->device);
  const uint64_t n = idx->ix.n;
  if (n == 0 || k == 0) {
    uint64_t total = std::max<uint64_t>((uint64_t)nq * std::max(k, 1u), nq);
    fill_empty_results_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_out_row_ids, d_out_node_ids, d_out_dist,
                                                                                   d_out_counts, nullptr, nq, k);
    CUDA_TRY(cudaGetLastError());
    return TURDB_OK;
  }
  if (k > TURDB_EXACT_MAX_K) return fail(TURDB_ERR_UNSUPPORTED, "bruteforce_topk: k = %u > %u", k, TURDB_EXACT_MAX_K);
  if (!rerank_factor) rerank_factor = 1;  // the certified band makes k' = k sufficient; more only widens the working set
  // kprime >= k rows stay in the working set: the slack band around the kprime-th key is what makes the filter exact;
  // a larger kprime only makes the band's lower edge less sensitive to outliers
  const uint32_t kprime = (uint32_t)std::min<uint64_t>(std::min<uint64_t>((uint64_t)k * rerank_factor, 2048), n);
  uint32_t cap = 2048;
  while (cap < 8 * std::max(kprime, k)) cap <<= 1;
  uint32_t* d_qflags = nullptr;
  CUDA_TRY(cudaMallocFromPoolAsync(&d_qflags, (size_t)nq * 4, idx->pool, stream));
  ExactFilterResult r;
  int32_t rc = exact_filter_run(idx, d_queries, nq, std::max(kprime, std::min<uint32_t>(k, (uint32_t)n)), metric, cap, cap / 2, 0,
                                nullptr, nullptr, d_qflags, &r, stream);
  if (rc != TURDB_OK) {
    cudaFreeAsync(d_qflags, stream);
    return rc;
  }
  const uint32_t ds = idx->ix.ds;
  const size_t rr_smem = (size_t)ds * 4 + (size_t)(cap / 2) * 8, st_smem = (size_t)ds * 4 + (size_t)k * 8 + 64 * 4;
  cudaError_t e = cudaSuccess;
  auto launch = [&](auto rerank, auto scan) {
    if (rr_smem > 48 * 1024) e = cudaFuncSetAttribute(rerank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem);
    if (e == cudaSuccess && st_smem > 48 * 1024) e = cudaFuncSetAttribute(scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem);
    if (e != cudaSuccess) return;
    rerank<<<nq, 128, rr_smem, stream>>>(idx->ix, d_queries, nq, k, cap, r.cand_cnt, r.cand_id, d_qflags, d_out_row_ids,
                                         d_out_node_ids, d_out_dist, d_out_counts);
    // queries whose buffers overflowed: the scan itself (exits at once when none is flagged)
    scan<<<(unsigned)std::min<uint64_t>(nq, 2ull * idx->num_sms), 256, st_smem, stream>>>(idx->ix, d_queries, nq, k, d_qflags,
                                                                                          d_out_row_ids, d_out_node_ids,
                                                                                          d_out_dist, d_out_counts);
  };
  switch (metric) {
    case kCosine: launch(exact_rerank_kernel<kCosine>, exact_stream_topk_kernel<kCosine>); break;
    case kIP: launch(exact_rerank_kernel<kIP>, exact_stream_topk_kernel<kIP>); break;
    default: launch(exact_rerank_kernel<kL2>, exact_stream_topk_kernel<kL2>);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFreeAsync(r.scr, stream);
  cudaFreeAsync(d_qflags, stream);
  if (e != cudaSuccess) return fail(TURDB_ERR_CUDA, "rerank launch failed: %s", cudaGetErrorString(e));
  return TURDB_OK;
}

extern "C" int32_t turdb_cuda_bruteforce_topk(turdb_cuda_index* idx, const float* queries, uint32_t query_dim,
                                              uint32_t nq, uint32_t k, uint8_t metric, uint32_t rerank_factor,
                                              uint64_t* out_row_ids, uint32_t* out_node_ids, float* out_dist,
                                              uint32_t* out_counts) {
  if (!idx) return fail(TURDB_ERR_INVALID_ARGUMENT, "idx is null");
  if (query_dim != idx->ix.dim)
    return fail(TURDB_ERR_DIMENSION_MISMATCH, "query dimension %u does not match index dimension %u", query_dim, idx->ix.dim);
  if (nq == 0) return TURDB_OK;
  if (!queries || !out_counts || (k && (!out_row_ids || !out_dist)))
    return fail(TURDB_ERR_INVALID_ARGUMENT, "null query/output pointer");
  DeviceGuard guard(idx->device);
  if (!guard.ok) return fail(TURDB_ERR_CUDA, "cudaSetDevice(%d) failed", idx->device);
  cudaStream_t stream;
  CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const size_t kk = std::max(k, 1u);
  const size_t qbytes = (size_t)nq * query_dim * 4;
  size_t off_rows = (qbytes + 255) & ~(size_t)255, off_dist = off_rows + nq * kk * 8, off_nodes = off_dist + nq * kk * 4,
         off_counts = off_nodes + nq * kk * 4, total = off_counts + (size_t)nq * 4;
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
  int32_t rc = turdb_cuda_bruteforce_topk_device(idx, (const float*)slab, query_dim, nq, k, metric, rerank_factor,
                                                 (uint64_t*)(slab + off_rows), (uint32_t*)(slab + off_nodes),
                                                 (float*)(slab + off_dist), (uint32_t*)(slab + off_counts), stream);
  if (rc != TURDB_OK) {
    cleanup();
    return rc;
  }
  if (k) {
    e = cudaMemcpyAsync(out_row_ids, slab + off_rows, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, slab + off_dist, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && out_node_ids)
      e = cudaMemcpyAsync(out_node_ids, slab + off_nodes, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_counts, slab + off_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cleanup();
  if (e != cudaSuccess) return fail(TURDB_ERR_CUDA, "bruteforce_topk failed: %s", cudaGetErrorString(e));
  return TURDB_OK;
}
