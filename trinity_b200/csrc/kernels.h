// Launch wrappers implemented in kernels.cu
#pragma once
#include "device_types.h"
#include <cuda_runtime.h>

namespace trn {
uint32_t    exec_stage_bytes(int codec);
size_t      exec_smem_bytes(uint32_t tile_shift, uint32_t nslots, int mode, int codec);
int         exec_max_ctas_per_sm(uint32_t tile_shift, uint32_t nslots, int mode, int codec);
cudaError_t launch_exec_tiles(const ExecParams &P, int grid, cudaStream_t stream);
uint32_t    exec_docs_stage_bytes();
uint32_t    exec_docs_cand_smem_bytes(bool with_membership); // per-warp shared memory of the candidate-driven path (membership bytes: trees with terms that are not necessary)
size_t      exec_docs_smem_bytes(uint32_t exec_shift, uint32_t nslots, uint32_t stageBytes);
int         exec_docs_max_ctas_per_sm(uint32_t exec_shift, uint32_t nslots, uint32_t stageBytes, bool tree = false, bool lucene = false);
cudaError_t launch_exec_docs(const ExecParams &P, int grid, cudaStream_t stream);
cudaError_t launch_query_scan(const unsigned long long *match_counts, uint32_t nq, uint64_t *q_offsets, cudaStream_t stream);
cudaError_t launch_item_scan(const DevQuery *queries, uint32_t nq, const uint32_t *item_cnt, const uint64_t *q_offsets, uint64_t *item_dst, cudaStream_t stream);
cudaError_t launch_gather(uint32_t total_items, const uint64_t *item_off, const uint32_t *item_cnt, const uint64_t *item_dst, const uint32_t *seg_docids,
                          const float *seg_scores, uint32_t *out_docids, float *out_scores, cudaStream_t stream);
cudaError_t launch_topk_select(const DevQuery *queries, uint32_t nq, const uint2 *cand, const uint32_t *cand_cursor, uint32_t k, uint32_t *out_docids,
                               float *out_scores, uint32_t *out_counts, cudaStream_t stream);
cudaError_t launch_topk_merge(const uint32_t *docids, const float *scores, uint32_t nshards, uint32_t nq, uint32_t k, uint32_t *out_docids, float *out_scores,
                              cudaStream_t stream);
cudaError_t launch_decode_terms(const DevIndex &ix, const uint32_t *term_ids, const uint32_t *unit_base, const uint64_t *out_base, uint32_t nterms,
                                uint32_t total_units, uint32_t *docids, uint32_t *freqs, unsigned long long *sums, int grid, cudaStream_t stream);
cudaError_t launch_decode_google(const DevIndex &ix, const uint32_t *term_ids, const uint32_t *unit_base, const uint64_t *out_base, uint32_t nterms,
                                 uint32_t total_units, uint32_t *docids, uint32_t *freqs, unsigned long long *sums, int grid, cudaStream_t stream);
size_t      score_flat_smem_bytes(uint32_t tile_shift, int threads);
uint32_t    score_flat_max_leaves();
cudaError_t launch_build_luts(const FlatLeaf *leaves, uint32_t nleaves, float *luts, cudaStream_t stream);
cudaError_t launch_score_flat(const ScoreParams &S, int threads, int num_sms, cudaStream_t stream);
cudaError_t launch_decode_stream(const DevIndex &ix, const DecUnit *units, const uint64_t *out_base, uint32_t total_units, uint32_t *docids, uint32_t *freqs,
                                 unsigned long long *sums, int num_sms, cudaStream_t stream);
cudaError_t launch_enc_scan(const uint32_t *in, uint64_t n, unsigned long long *partials, unsigned long long *out, cudaStream_t stream);
cudaError_t launch_enc_google_sizes(const EncParams &E, cudaStream_t stream);
cudaError_t launch_enc_term_sizes(const EncParams &E, unsigned long long *chunk_bytes, cudaStream_t stream);
cudaError_t launch_enc_google_write(const EncParams &E, cudaStream_t stream);
uint32_t    kernel_max_k();
} // namespace trn
