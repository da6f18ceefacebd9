// Host-callable launchers of the device kernels (one translation unit each).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bb_common.cuh"

struct CUtensorMap_st;

namespace bb {

// ---- K1x exact fp32 scan (bb_scan_fp32.cu)
constexpr int kFp32QRows = 128;   // query rows per unit
constexpr int kFp32Tile = 128;    // catalog rows per tile
cudaError_t launch_scan_fp32(const ScanParams& P, const float* Q, const float* E, cudaStream_t stream);

// ---- K1 tcgen05 bf16 scan (bb_scan_tc.cu)
constexpr int kTcTile = 256;      // catalog rows per MMA tile (UMMA N); 512 / kTcTile TMEM accumulator stages (N=128 is smem-bandwidth bound: measured 995 vs 1524 TFLOP/s MMA-only)
constexpr int kTcRowsPerCta = 128;
constexpr int kTcMaxKBlocks = 7;  // resident query operand: 7 x 64 bf16 columns
constexpr int kTcSubStreams = 1;  // candidate streams per (split, row): the four column-slice warps of a row group share one
// cta_group: 1 or 2.  tmap_q / tmap_e are CUtensorMap objects (host memory, passed by value to the kernel).
// tmap_ql / tmap_el: tensor maps of the second bf16 terms (query / catalog) -> the exact-fp32 split scan; nullptr = bf16 scan
cudaError_t launch_scan_tc(const ScanParams& P, const void* tmap_q, const void* tmap_e, int cta_group, int n_groups,
                           cudaStream_t stream, const void* tmap_ql = nullptr, const void* tmap_el = nullptr);

// ---- K2 operand preparation (bb_aux.cu)
cudaError_t launch_prep_catalog(const float* emb, const float* cf, int64_t n_rows, int dim, int cf_dim, int k_pad,
                                int metric, float* out_f32, void* out_bf16, void* out_lo, unsigned* norm_max, cudaStream_t stream);
cudaError_t launch_prep_queries(const float* q_emb, const int32_t* q_item, const float* q_cf, const float* w_c,
                                const float* w_f, float w_c_all, float w_f_all, int n_queries, int b_pad, int dim, int cf_dim,
                                int k_pad, int metric, const float* cat_f32, const void* cat_bf16, int64_t id_offset,
                                int64_t n_items, float* out_f32, void* out_bf16, void* out_lo, int32_t* err_flag,
                                cudaStream_t stream);
cudaError_t launch_pack_meta(const int32_t* num_parts, const int16_t* year, const int16_t* theme, int64_t n_rows,
                             int2* out, cudaStream_t stream);
// exact fp32 path for B <= kSmallBatch queries (bb_scan_small.cu): dense score strip [B_template x n_items] + per-split
// selection into the candidate slots (slot = split * b_pad + q, P.cap keys each); merge_emit finishes
constexpr int kSmallBatch = 16;
int small_scan_splits(int64_t n_items);
cudaError_t launch_scan_small(const ScanParams& P, const float* q_f32, const float* cat_f32, float* scores, int sm_count,
                              cudaStream_t stream);
// Y[n_rows x L] = CSR(n_rows x *) * X[* x L], fp32 values, fp64 X / Y (bb_spmm.cu); L <= 128
cudaError_t launch_csr_spmm_f64(int64_t n_rows, int64_t nnz, const int64_t* indptr, const int32_t* indices, const float* values,
                                const double* X, int L, double* Y, int sm_count, cudaStream_t stream);
cudaError_t launch_gather_rows(const float* cat_f32, const void* cat_bf16, int k_pad, int dim, int64_t id_offset,
                               int64_t n_items, const int32_t* ids, int n, float* out, int32_t* err_flag,
                               cudaStream_t stream);

// ---- K3 constraint mask (bb_aux.cu)
cudaError_t launch_mask(const ScanParams& P, int n_constraints, uint32_t* out_bitmaps, unsigned long long* out_counts,
                        cudaStream_t stream);

// ---- K4 + K6 merge / re-score / emit (bb_aux.cu)
// comm != nullptr: collective search of a row-sharded catalog -- the k final keys of every query are stored into the
// owner rank's inbox (peer memory) instead of out_scores / out_ids
// rescore: 0 = rank by the scan's scores, n > 0 = re-score the best n candidates in exact fp32
cudaError_t launch_merge_emit(const ScanParams& P, int k, int rescore, const float* q_f32, const float* cat_f32,
                              bool as_l2, float* out_scores, int32_t* out_ids, const CommParams* comm, cudaStream_t stream,
                              int32_t* row_flags = nullptr, int32_t* n_flagged = nullptr, const unsigned* cat_norm_max = nullptr);
cudaError_t launch_merge_inbox(const CommParams& comm, int n_mine, int k, unsigned long long expected, bool as_l2,
                               float* out_scores, int32_t* out_ids, int32_t* err_flag, cudaStream_t stream);
// threshold (and level) seeds from the pilot pass's group maxima (P.pilot_max / P.pilot_groups)
cudaError_t launch_seed_thresholds(const ScanParams& P, int64_t pilot_items, cudaStream_t stream);
cudaError_t launch_merge_lists(int n_queries, int n_lists, int len, const float* scores, const int32_t* ids, int k,
                               float* out_scores, int32_t* out_ids, cudaStream_t stream);

// ---- K5 reference-semantics hybrid combine (bb_aux.cu)
cudaError_t launch_hybrid_combine(int n_rows, int len_c, const int32_t* c_ids, const float* c_scores, int len_f,
                                  const int32_t* f_ids, const float* f_scores, const double* weights,
                                  int weights_per_row, int k, double* out_scores, int32_t* out_ids,
                                  cudaStream_t stream);

// ---- item-item CF (bb_itemitem.cu)
cudaError_t launch_item_item(int64_t n_users, int n_items, const int64_t* indptr, const int32_t* indices,
                             const float* values, int k, float* out_scores, int32_t* out_ids, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream, int* n_launches);
size_t item_item_workspace_bytes(int64_t n_users, int n_items, int64_t nnz);

}  // namespace bb
