/*
 * brickscore.h -- C ABI of libbrickscore.so, the B200 (sm_100a) candidate-scoring engine.
 *
 * This is the drop-in boundary for the Brickbrain-Rec-Engine scoring hot path.
 * The reference (pure Python) has no FFI layer; each entry point below replaces
 * the reference code named in its comment (paths relative to the reference repo
 * root, src/scripts/...).  The Python adapters in brickbrain_rec_engine_b200/
 * bind these with ctypes and keep the reference's method signatures
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; caller-allocated outputs; opaque handle.
 *   - every pointer argument may be HOST or DEVICE memory (detected with
 *     cudaPointerGetAttributes).  With device pointers the call only enqueues
 *     work on `stream`; with host pointers it copies and synchronizes `stream`
 *     before returning.
 *   - return 0 on success, negative errno-style code on failure; the message is
 *     available from bb_last_error().  Nothing throws across the ABI.
 *   - ids are int32 GLOBAL catalog row indices (shard row + id_offset);
 *     missing results are id -1 / score -inf.
 *   - ordering everywhere: descending score, ties -> lower id first.
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with BB_ERR_CUDA.
 */
#ifndef BRICKSCORE_H
#define BRICKSCORE_H

#include <stdint.h>

#if defined(__GNUC__)
#define BB_API __attribute__((visibility("default")))
#else
#define BB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BB_OK 0
#define BB_ERR_INVALID (-22)     /* EINVAL: bad argument / shape */
#define BB_ERR_NOMEM (-12)       /* ENOMEM: device or host allocation failed */
#define BB_ERR_CUDA (-5)         /* EIO: CUDA runtime / driver error, or no device */
#define BB_ERR_STATE (-1)        /* EPERM: call order (e.g. search before load) */
#define BB_ERR_UNSUPPORTED (-95) /* EOPNOTSUPP */

/* bb_config.metric */
#define BB_METRIC_COSINE 0 /* rows and queries L2-normalized (zero rows stay zero): sklearn cosine_similarity */
#define BB_METRIC_IP 1     /* raw inner product */

/* bb_config.flags */
#define BB_CFG_KEEP_FP32 1u  /* keep the fp32 operand copy (exact path, re-scoring) */
#define BB_CFG_BUILD_BF16 2u /* build the bf16 operand copy (tcgen05 path) */
#define BB_CFG_BUILD_SPLIT 4u /* build the second bf16 term of the operands (+2 B per element): BB_PREC_FP32 searches then run on
                                 the tensor pipe (split-bf16 scan, exact re-score, proof of completeness, FFMA fallback per row).
                                 Default (flags = 0): built when the plane is <= 4 GiB */

/* bb_query_batch.precision */
#define BB_PREC_FP32 0         /* exact fp32 search, the parity path: ids and scores of an fp32 evaluation.  With the split plane:
                                  tensor-pipe scan + exact re-score + per-row proof (flagged rows are re-scanned with FFMAs);
                                  without it, or for collective searches: FFMA scan; <= 16 queries: one HBM pass */
#define BB_PREC_BF16 1         /* tcgen05 bf16 scan, fp32 accumulate, ranks by bf16 scores */
#define BB_PREC_BF16_RESCORE 2 /* tcgen05 bf16 scan over-fetches, survivors re-scored in exact fp32 */

/* bb_query_batch.flags */
#define BB_Q_SCORES_AS_L2 1u /* report 2 - 2*score (FAISS IndexFlatL2 distance on unit vectors) */
#define BB_Q_SHARE_THRESHOLDS 2u /* collective search of a row-sharded catalog (union thresholds over peer memory), see bb_comm_import */
#define BB_Q_EMB_PRENORMALIZED 4u /* q_emb rows are stored catalog rows (already unit norm): use them as they are */

/* bb_constraint.flags */
#define BB_C_ALWAYS_FALSE 1u /* THEMES_REQUIRED matched no theme: "1=0" (hard_constraint_filter.py:408-409) */

typedef struct bb_handle bb_handle;

typedef struct bb_config {
  uint32_t struct_size; /* sizeof(bb_config) */
  int32_t device;       /* CUDA ordinal */
  int64_t n_items;      /* catalog rows held by THIS handle (one shard) */
  int64_t id_offset;    /* global id of local row 0 (row-sharded catalogs) */
  int32_t dim;          /* D, content embedding width (0 = none) */
  int32_t cf_dim;       /* F, CF factor width (0 = none) */
  int32_t k_max;        /* largest k that will be requested (<= 128) */
  int32_t max_batch;    /* largest n_queries per bb_search call */
  int32_t theme_bits;   /* width of a theme bitmap, multiple of 32 (0 -> 1024) */
  int32_t metric;       /* BB_METRIC_* */
  uint32_t flags;       /* BB_CFG_* (0 -> KEEP_FP32 | BUILD_BF16) */
} bb_config;

/*
 * One AND-of-predicates hard constraint, compiled on the host from the
 * reference's HardConstraint list (hard_constraint_filter.py:366-480): every
 * numeric predicate there is an integer compare on sets.num_parts / sets.year,
 * so an AND of them is a pair of inclusive ranges.  SQL NULL theme_id is -1 and
 * fails whenever theme_set >= 0 (three-valued logic of `= ANY` / `NOT (= ANY)`).
 */
typedef struct bb_constraint {
  int32_t parts_min, parts_max; /* inclusive; "num_parts > 0" (:343) is parts_min >= 1 */
  int16_t year_min, year_max;   /* inclusive */
  int32_t theme_set;            /* index into bb_set_theme_sets bitmaps (allowed ids), -1 = no theme predicate */
  int32_t allow_bitmap;         /* index into bb_set_allow_bitmaps (explicit valid_set_filter), -1 = none */
  int32_t excl_list[3];         /* indices into bb_set_exclusion_lists (rated / owned / wishlisted), -1 = none */
  uint32_t flags;               /* BB_C_* */
} bb_constraint;

typedef struct bb_query_batch {
  uint32_t struct_size;             /* sizeof(bb_query_batch) */
  int32_t n_queries;                /* B */
  const float* q_emb;               /* [B x dim] raw query embeddings, or NULL */
  const int32_t* q_item;            /* [B] global ids whose stored embedding is the query (liked set), or NULL */
  const float* q_cf;                /* [B x cf_dim] user factors, or NULL */
  const float* w_content;           /* [B] per-query content weight, or NULL -> w_content_all */
  const float* w_cf;                /* [B] per-query CF weight, or NULL -> w_cf_all */
  float w_content_all, w_cf_all;    /* reference default 0.4 / 0.6 (recommendation_system.py:609-610) */
  const bb_constraint* constraints; /* n_constraints entries, or NULL */
  int32_t n_constraints;            /* 0, 1 (shared by all queries), B (one per query), or C distinct sets when
                                       constraint_index is given */
  const int32_t* constraint_index;  /* [B] index of each query's entry in constraints[], or NULL.  Batches whose
                                       queries share a few distinct constraint sets (<= 64) are masked through
                                       per-set validity bitmaps built once per call by the mask kernel */
  const int32_t* self_id;           /* [B] global id excluded per query (-1 none), or NULL */
  int32_t precision;                /* BB_PREC_* */
  uint32_t flags;                   /* BB_Q_* */
  /* Per-CALL validity bitmaps over LOCAL rows, [n_allow x ceil(n_items/32)] words, host or device memory: the
   * valid_set_filter of ONE request (recommendation_system.py:229,454), e.g. the device bitmap written by
   * bb_apply_constraints.  When given, bb_constraint.allow_bitmap indexes THIS array instead of the handle-resident
   * one (bb_set_allow_bitmaps), so concurrent requests never share filter state.  NULL = handle-resident bitmaps. */
  const uint32_t* allow_bitmaps;
  int32_t n_allow;
  /* Per-QUERY exclusion lists: [B x n_excl_per_query] indices into bb_set_exclusion_lists (-1 = none): the user's
   * already-rated items (recommendation_system.py:441-451).  Kept out of bb_constraint so that queries that share a
   * constraint set still share one descriptor (and one validity bitmap) while every user carries an own list. */
  int32_t n_excl_per_query;         /* 0..3 */
  const int32_t* excl_lists;
} bb_query_batch;

typedef struct bb_info {
  uint32_t struct_size;
  int32_t sm_count;
  int32_t k_pad;            /* operand width D+F padded to a multiple of 64 */
  int32_t cta_group;        /* 2 = tcgen05 cta_group::2 CTA pairs, 1 = single CTA */
  int64_t launches;         /* kernels launched by this handle so far */
  int64_t last_units;       /* (query block, catalog split) work units of the last search */
  int64_t last_candidates;  /* candidates left in the buffers at the end of the last scan (debug) */
  int64_t last_appends;     /* candidates appended by the last scan's epilogue */
  int64_t last_slow_chunks; /* 32-column chunks that entered the epilogue slow path */
  int64_t last_compactions; /* candidate-buffer compactions (warp sorts) */
  int64_t mma_wait_epilogue_cycles; /* sum over MMA issuers: cycles waiting for a free TMEM accumulator */
  int64_t mma_wait_tma_cycles;      /* sum over MMA issuers: cycles waiting for operand stages */
  int64_t mma_total_cycles;         /* sum over MMA issuers: cycles in the unit loop */
  double bytes_fp32, bytes_bf16; /* resident operand bytes */
  double scan_ms;           /* device time of the scan kernels since bb_set_profiling(h, 1) (CUDA events) */
  int64_t scan_launches;    /* scan kernel launches covered by scan_ms */
  /* sum over epilogue warps of the last scan, in cycles: [0] waiting for a full accumulator, [1] slow path
   * (threshold passed: mask + append), [2] in-scan compactions, [3] unit-end compaction + refine, [4] total,
   * [5] accumulator hold (full -> released), [6] TMEM loads (issue -> data in registers) */
  int64_t epilogue_cycles[7];
  /* (profiling build) per scheduling round of the last scan, summed over CTA pairs: unit duration and the part of
   * it spent waiting for the query operand (includes the wave barrier); rounds >= 15 share the last bucket */
  int64_t round_cycles[16];
  int64_t round_wait_cycles[16];
  int64_t last_wasted_chunks; /* slow-path chunks whose passing scores were all masked out (no append) */
  int64_t last_flagged_rows;  /* exact-fp32 search on the tensor pipe: rows of the last search that fell back to the FFMA scan */
} bb_info;

BB_API const char* bb_version(void);
BB_API const char* bb_last_error(const bb_handle* h); /* h may be NULL: last error of bb_create on this thread */

BB_API int bb_create(bb_handle** out, const bb_config* cfg);
BB_API void bb_destroy(bb_handle* h);
BB_API int bb_get_info(const bb_handle* h, bb_info* info);
/* enable != 0: bracket every scan-kernel launch with CUDA events on the launching stream and accumulate
 * their device time into bb_info.scan_ms / scan_launches (read back, synchronizing, by bb_get_info);
 * enabling resets the counters.  Used by bench.py for the roofline figure. */
BB_API int bb_set_profiling(bb_handle* h, int enable);

/*
 * Loads / replaces catalog rows [row_begin, row_begin + n_rows): embeddings
 * (normalized once here -- replaces the per-query sklearn normalize of the whole
 * catalog inside cosine_similarity, recommendation_system.py:213-214), CF item
 * factors (svd components_.T, recommendation_system.py:402-407) and the three
 * integer metadata columns the constraint SQL reads (src/db/rebrickable_schema.sql:38-46).
 * emb / item_factors may be NULL when dim / cf_dim is 0; metadata NULL -> num_parts 1, year 0, theme NULL.
 */
BB_API int bb_load_catalog(bb_handle* h, int64_t row_begin, int64_t n_rows, const float* emb, const float* item_factors,
                    const int32_t* num_parts, const int16_t* year, const int16_t* theme_id, void* stream);

/* Copies stored (normalized) embedding rows out: [n x dim].  ids are global and must live in this shard. */
BB_API int bb_fetch_embeddings(bb_handle* h, int32_t n, const int32_t* ids, float* out, void* stream);

/*
 * Per-user exclusion lists in CSR form (sorted ascending GLOBAL ids per list):
 * already-rated items (recommendation_system.py:441-451) and the owned /
 * wishlisted anti-joins (hard_constraint_filter.py:441-451).
 */
BB_API int bb_set_exclusion_lists(bb_handle* h, int64_t n_lists, const int64_t* offsets, const int32_t* item_ids);
/* Allowed-theme bitmaps, n_sets x (theme_bits/32) words; bit t set = theme id t allowed. */
BB_API int bb_set_theme_sets(bb_handle* h, int32_t n_sets, const uint32_t* bitmaps);
/* Explicit validity bitmaps over LOCAL rows, n x ceil(n_items/32) words (valid_set_filter lists,
 * recommendation_system.py:229,454).  bitmaps may be a device pointer produced by bb_apply_constraints. */
BB_API int bb_set_allow_bitmaps(bb_handle* h, int32_t n_bitmaps, const uint32_t* bitmaps);

/*
 * The fused hot path: score = w_c * <q_hat, e_hat> + w_f * <u, v> over the whole
 * shard, hard-constraint mask, self / exclusion-list removal, top-k.
 * Replaces recommendation_system.py:213-247 (content), :432-461 (CF), the
 * full-catalog form of :789-843 (blend) and the filter walk over
 * hard_constraint_filter.py:263-316 results; also the vector-store search behind
 * lego_nlp_recommeder.py:296-305,1394 and hf_nlp_recommender.py:1310-1349.
 * out_scores [B x k] float, out_ids [B x k] int32.
 */
BB_API int bb_search(bb_handle* h, const bb_query_batch* q, int32_t k, float* out_scores, int32_t* out_ids, void* stream);

/*
 * HardConstraintFilter.apply_constraints as a mask kernel (hard_constraint_filter.py:263-316,
 * SQL semantics of :318-480): n constraints -> n bitmaps over LOCAL rows
 * (ceil(n_items/32) words each) and n valid counts.  Exclusion lists are applied;
 * allow_bitmap is ANDed if set.
 */
BB_API int bb_apply_constraints(bb_handle* h, const bb_constraint* constraints, int32_t n, uint32_t* out_bitmaps,
                         int64_t* out_counts, void* stream);

/*
 * HybridRecommender._combine_recommendations (recommendation_system.py:789-843),
 * reference semantics: per row the union of a content list and a CF list
 * (id -1 = empty slot), a missing side contributes 0, fp64 blend, descending,
 * first k.  weights: [2] shared or [n_rows x 2].
 */
BB_API int bb_hybrid_combine(bb_handle* h, int32_t n_rows, int32_t len_c, const int32_t* c_ids, const float* c_scores,
                      int32_t len_f, const int32_t* f_ids, const float* f_scores, const double* weights,
                      int32_t weights_per_row, int32_t k, double* out_scores, int32_t* out_ids, void* stream);

/*
 * Cross-shard merge after the all-gather: n_lists candidate lists of
 * length len per query, laid out [n_lists x n_queries x len] -> top-k.
 */
BB_API int bb_merge_topk(bb_handle* h, int32_t n_queries, int32_t n_lists, int32_t len, const float* scores,
                  const int32_t* ids, int32_t k, float* out_scores, int32_t* out_ids, void* stream);

/*
 * Item-item collaborative filtering (BASELINE config 3; absent in the reference,
 * schema stub at src/db/user_interaction_schema.sql:136-145): cosine between item
 * columns of the user x item CSR rating matrix, top-k neighbours per item
 * (self excluded).  out_* are [n_items x k].
 */
BB_API int bb_item_item_topk(bb_handle* h, int64_t n_users, int32_t n_items, const int64_t* indptr, const int32_t* indices,
                      const float* values, int32_t k, float* out_scores, int32_t* out_ids, void* stream);

/*
 * Row-sharded catalogs: every rank (one process per GPU, or several handles in one process) holds a
 * contiguous range of rows; queries are replicated.  Each handle created with max_batch > 0 owns a COMMUNICATION
 * BLOCK of device memory that the other ranks map (cudaIpc across processes, peer access inside one):
 *   - threshold + level words: while scanning, every rank publishes per query "I hold >= k'/8, k'/4, k'/2, k'
 *     candidates with score >= T"; if `need` ranks each hold m candidates >= T and need * m >= k', T is a lower bound
 *     of the GLOBAL k'-th best.  A poller warp of the scan kernel reads the peers' words over NVLink (plain loads) and
 *     raises the local threshold, so every shard prunes against the global running top-k' instead of its own;
 *   - an inbox + arrival counters: query q is owned by rank q / ceil(B / world); the merge kernel of every rank
 *     stores its k final (score, id) keys of q straight into the owner's inbox (peer stores fused into the kernel)
 *     and the owner merges the `world` lists of its slice.
 * Setup is collective: every rank exports its handle, the handles are exchanged (e.g. all-gather) and imported, in
 * rank order, before any rank searches.  bb_search_sharded is collective too: all ranks issue the same sequence of
 * sharded searches with the same replicated batch.  Its outputs are the results of THIS rank's slice of the queries,
 * rows [rank * slice, min(B, (rank + 1) * slice)), slice = ceil(B / world): out_* are [n_mine x k].
 * bb_search with BB_Q_SHARE_THRESHOLDS only shares the thresholds and leaves the shard-local candidates in out_*
 * (fewer than k entries, -1 padded, where the global bound pruned the rest) for an external exchange (NCCL all-gather
 * + bb_merge_topk).
 */
#define BB_IPC_HANDLE_BYTES 64
BB_API int bb_comm_export(bb_handle* h, void* out_handle);
BB_API int bb_comm_import(bb_handle* h, int32_t rank, int32_t world, const void* handles /* world x 64 B, own entry ignored */);
/* same process (several GPUs, or several shards on one GPU in tests): handles[world], own entry ignored */
BB_API int bb_comm_link_local(bb_handle* h, int32_t rank, int32_t world, bb_handle* const* handles);
BB_API int bb_search_sharded(bb_handle* h, const bb_query_batch* q, int32_t k, float* out_scores, int32_t* out_ids, void* stream);
/* The two halves of bb_search_sharded, for callers that drive several shards from ONE thread (a process with several
 * GPUs, or several shards on one GPU): scatter on every shard first, then collect on every shard.  scatter only
 * enqueues work; collect waits on the device until all ranks' keys of this rank's slice have arrived. */
BB_API int bb_sharded_scatter(bb_handle* h, const bb_query_batch* q, int32_t k, void* stream);
BB_API int bb_sharded_collect(bb_handle* h, float* out_scores, int32_t* out_ids, void* stream);
/* Pipelined form for back-to-back collective searches: scatters the NEXT search and launches the collect of the previous
 * (not yet collected) scatter into prev_out_* (DEVICE memory, [n_mine x k] of that search) in between -- after the next
 * search's operand preparation, masks and threshold seeds, before its scan.  The owner-side merge waits for the slowest
 * peer's keys; here that wait overlaps work that does not depend on the peers.  prev_out_* are complete once the work
 * enqueued by this call has run; the last search of a sequence is finished with bb_sharded_collect. */
BB_API int bb_sharded_scatter_after(bb_handle* h, const bb_query_batch* q, int32_t k, float* prev_out_scores,
                                    int32_t* prev_out_ids, void* stream);

/*
 * Sparse (CSR ratings, fp32 values) x dense (fp64, row-major) product, Y[n_rows x L] = A[n_rows x n_cols] * X[n_cols x L]:
 * the A @ Q, A.T @ Q (pass the CSR of A.T) and A @ V products of the randomized truncated SVD that trains the
 * collaborative-filtering factors (reference: recommendation_system.py:368-409, sklearn TruncatedSVD).  L <= 128.
 * All pointers are DEVICE pointers (the training loop keeps its matrices resident).  Rows with more than 2048
 * entries are summed with fp64 atomics (last-bit run-to-run differences); shorter rows are deterministic.
 */
BB_API int bb_csr_spmm_f64(bb_handle* h, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                           const float* values, const double* x, int32_t l, double* y, void* cuda_stream);

/* Test / diagnosis aid: dense score matrix [B x n_items] from the selected scan pipeline. */
BB_API int bb_debug_scores(bb_handle* h, const bb_query_batch* q, float* out_scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BRICKSCORE_H */
