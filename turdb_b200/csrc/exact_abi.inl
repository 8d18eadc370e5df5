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
  merge_topk_kernel<<<blocks, threads, 0, stream>>>(d_gathered_row_ids, d_gathered_dist, d_gathered_counts, n_shards,
                                                    nq, k, d_out_row_ids, d_out_dist, d_out_counts);
  CUDA_TRY(cudaGetLastError());
  return TURDB_OK;
}

extern "C" int32_t turdb_cuda_bruteforce_topk_device(turdb_cuda_index* idx, const float* d_queries, uint32_t query_dim,
                                                     uint32_t nq, uint32_t k, uint8_t metric, uint32_t rerank_factor,
                                                     uint64_t* d_out_row_ids, uint32_t* d_out_node_ids,
                                                     float* d_out_dist, uint32_t* d_out_counts, void* stream_) {
  (void)idx; (void)d_queries; (void)query_dim; (void)nq; (void)k; (void)metric; (void)rerank_factor;
  (void)d_out_row_ids; (void)d_out_node_ids; (void)d_out_dist; (void)d_out_counts; (void)stream_;
  return fail(TURDB_ERR_UNSUPPORTED, "bruteforce_topk: not built yet");
}

extern "C" int32_t turdb_cuda_bruteforce_topk(turdb_cuda_index* idx, const float* queries, uint32_t query_dim,
                                              uint32_t nq, uint32_t k, uint8_t metric, uint32_t rerank_factor,
                                              uint64_t* out_row_ids, uint32_t* out_node_ids, float* out_dist,
                                              uint32_t* out_counts) {
  (void)idx; (void)queries; (void)query_dim; (void)nq; (void)k; (void)metric; (void)rerank_factor;
  (void)out_row_ids; (void)out_node_ids; (void)out_dist; (void)out_counts;
  return fail(TURDB_ERR_UNSUPPORTED, "bruteforce_topk: not built yet");
}
