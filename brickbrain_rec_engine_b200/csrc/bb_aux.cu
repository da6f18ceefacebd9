// Bandwidth kernels around the scan: operand preparation (K2), constraint mask
// (K3), candidate merge / exact re-score / emit (K4 + K6), cross-shard list merge
// and the reference-semantics hybrid combine (K5).  All are HBM- or latency-bound
// integer / fp32 work: coalesced 128-bit accesses, one warp per row or per query.
#include <cuda_bf16.h>

#include "bb_common.cuh"
#include "bb_kernels.h"

namespace bb {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ------------------------------------------------------------------ K2: catalog rows
// One warp per row: L2 norm of the embedding part (sklearn normalize semantics:
// zero rows are divided by 1, recommendation_system.py:214), write
// [e_hat | v | 0-pad] as fp32 and bf16.
__global__ void __launch_bounds__(256)
prep_catalog_kernel(const float* __restrict__ emb, const float* __restrict__ cf, int64_t n_rows, int dim, int cf_dim,
                    int k_pad, int metric, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                    __nv_bfloat16* __restrict__ out_lo, unsigned* __restrict__ norm_max) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  float inv_norm_den = 1.f;
  if (dim > 0 && emb != nullptr && metric == BB_METRIC_COSINE) {
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float x = __ldg(emb + row * dim + c);
      ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    inv_norm_den = (nrm == 0.f) ? 1.f : nrm;
  }
  // emb == nullptr with dim > 0: factor-only reload (re-trained CF factors) -- the embedding columns are kept
  float ss_op = 0.f;
  for (int c = (emb == nullptr ? dim : 0) + lane; c < k_pad; c += 32) {
    float x = 0.f;
    if (c < dim) x = __fdiv_rn(__ldg(emb + row * dim + c), inv_norm_den);
    else if (c < dim + cf_dim && cf != nullptr) x = __ldg(cf + row * cf_dim + (c - dim));
    if (out_f32 != nullptr) out_f32[row * k_pad + c] = x;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    if (out_bf16 != nullptr) out_bf16[row * k_pad + c] = hi;
    // second bf16 term of the operand split x = hi + lo + r, |r| <= 2^-18 |x| (exact-fp32 search on the tensor pipe)
    if (out_lo != nullptr) out_lo[row * k_pad + c] = __float2bfloat16_rn(x - __bfloat162float(hi));
    ss_op = fmaf(x, x, ss_op);
  }
  if (norm_max != nullptr) {
    // largest operand-row norm of the catalog (positive floats order like their bit patterns); with a factor-only reload
    // the kept embedding part (unit norm or zero) is bounded by 1
    ss_op = warp_sum(ss_op) + (emb == nullptr && dim > 0 ? 1.f : 0.f);
    if (lane == 0) atomicMax(norm_max, __float_as_uint(sqrtf(ss_op) * 1.0001f));
  }
}

// ------------------------------------------------------------------ K2: query rows
// One warp per query: [w_c * q_hat | w_f * u | 0-pad]; q_hat is either the
// normalized raw embedding or a stored catalog row (liked set, global id).
__global__ void __launch_bounds__(256)
prep_queries_kernel(const float* __restrict__ q_emb, const int32_t* __restrict__ q_item, const float* __restrict__ q_cf,
                    const float* __restrict__ w_c, const float* __restrict__ w_f, float w_c_all, float w_f_all,
                    int n_queries, int b_pad, int dim, int cf_dim, int k_pad, int metric, const float* __restrict__ cat_f32,
                    const __nv_bfloat16* __restrict__ cat_bf16, int64_t id_offset, int64_t n_items,
                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, __nv_bfloat16* __restrict__ out_lo,
                    int32_t* err_flag) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= b_pad) return;
  if (row >= n_queries) {                     // padding rows of the last query block: zero operands
    for (int c = lane; c < k_pad; c += 32) {
      out_f32[static_cast<int64_t>(row) * k_pad + c] = 0.f;
      out_bf16[static_cast<int64_t>(row) * k_pad + c] = __float2bfloat16_rn(0.f);
      if (out_lo != nullptr) out_lo[static_cast<int64_t>(row) * k_pad + c] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  const float wc = w_c != nullptr ? __ldg(w_c + row) : w_c_all;
  const float wf = w_f != nullptr ? __ldg(w_f + row) : w_f_all;
  int64_t src_item = -1;
  if (q_item != nullptr) {
    src_item = static_cast<int64_t>(__ldg(q_item + row)) - id_offset;
    if (src_item < 0 || src_item >= n_items) {
      if (lane == 0) atomicExch(err_flag, 1);
      src_item = -2;
    }
  }
  float den = 1.f;
  if (q_emb != nullptr && src_item == -1 && metric == BB_METRIC_COSINE) {
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float x = __ldg(q_emb + static_cast<int64_t>(row) * dim + c);
      ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    den = (nrm == 0.f) ? 1.f : nrm;
  }
  for (int c = lane; c < k_pad; c += 32) {
    float x = 0.f;
    if (c < dim) {
      if (src_item >= 0) {
        x = cat_f32 != nullptr ? __ldg(cat_f32 + src_item * k_pad + c) : __bfloat162float(cat_bf16[src_item * k_pad + c]);
        x = __fmul_rn(wc, x);
      } else if (q_emb != nullptr && src_item == -1) {
        x = __fmul_rn(wc, __fdiv_rn(__ldg(q_emb + static_cast<int64_t>(row) * dim + c), den));
      }
    } else if (c < dim + cf_dim && q_cf != nullptr) {
      x = __fmul_rn(wf, __ldg(q_cf + static_cast<int64_t>(row) * cf_dim + (c - dim)));
    }
    out_f32[static_cast<int64_t>(row) * k_pad + c] = x;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    out_bf16[static_cast<int64_t>(row) * k_pad + c] = hi;
    if (out_lo != nullptr) out_lo[static_cast<int64_t>(row) * k_pad + c] = __float2bfloat16_rn(x - __bfloat162float(hi));
  }
}

__global__ void pack_meta_kernel(const int32_t* __restrict__ num_parts, const int16_t* __restrict__ year,
                                 const int16_t* __restrict__ theme, int64_t n, int2* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t p = num_parts != nullptr ? num_parts[i] : 1;
  const int32_t y = year != nullptr ? year[i] : 0;
  const int32_t t = theme != nullptr ? theme[i] : -1;
  out[i] = make_int2(p, (y & 0xffff) | (t << 16));
}

__global__ void gather_rows_kernel(const float* __restrict__ cat_f32, const __nv_bfloat16* __restrict__ cat_bf16,
                                   int k_pad, int dim, int64_t id_offset, int64_t n_items,
                                   const int32_t* __restrict__ ids, int n, float* __restrict__ out, int32_t* err_flag) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const int64_t item = static_cast<int64_t>(ids[r]) - id_offset;
  if (item < 0 || item >= n_items) {
    if (lane == 0) atomicExch(err_flag, 1);
    return;
  }
  for (int c = lane; c < dim; c += 32)
    out[static_cast<int64_t>(r) * dim + c] =
        cat_f32 != nullptr ? cat_f32[item * k_pad + c] : __bfloat162float(cat_bf16[item * k_pad + c]);
}

// ------------------------------------------------------------------ K3: constraint mask
// One pass over the metadata for ALL constraints of the launch (<= kMaskCons, descriptors in shared memory): a warp
// handles 32 items per step (lane = item, one 8-byte metadata read), evaluates every descriptor on the registers and
// lane c keeps the ballot word of constraint c, so a step ends with one store per constraint.
// Algorithmic bytes: 8 B metadata read per item + 1/8 B written per item per constraint.
constexpr int kMaskCons = 32;
__global__ void __launch_bounds__(256)
mask_kernel(const ScanParams P, int c_begin, int n_cons, uint32_t* __restrict__ out_bitmaps,
            unsigned long long* __restrict__ out_counts) {
  __shared__ bb_constraint s_cons[kMaskCons];
  for (int i = threadIdx.x; i < n_cons; i += blockDim.x) s_cons[i] = P.cons[c_begin + i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t n_words = (P.n_items + 31) >> 5;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  unsigned long long count = 0;                  // lane c: valid items of constraint c seen by this warp
  constexpr int U = 4;                           // bitmap words per warp step: four independent 256-byte metadata reads in flight
  for (int64_t w0 = warp0 * U; w0 < n_words; w0 += n_warps * U) {
    int2 m[U];
    bool in[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t item = ((w0 + u) << 5) + lane;
      in[u] = item < P.n_items;
      m[u] = in[u] ? __ldg(P.meta + item) : make_int2(0, 0);
    }
    unsigned mine[U] = {0u, 0u, 0u, 0u};
    for (int c = 0; c < n_cons; ++c) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned bits = __ballot_sync(kFull, in[u] && constraint_ok_meta(P, s_cons[c], ((w0 + u) << 5) + lane, m[u]));
        if (lane == c) mine[u] = bits;
      }
    }
    if (lane < n_cons) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (w0 + u < n_words) {
          out_bitmaps[static_cast<int64_t>(c_begin + lane) * n_words + w0 + u] = mine[u];
          count += __popc(mine[u]);
        }
      }
    }
  }
  if (lane < n_cons && count) atomicAdd(out_counts + c_begin + lane, count);
}

// ------------------------------------------------------------------ running top-k' merger (one warp)
// stage[0..127] = best so far (sorted descending after a flush), stage[128..255] = new keys.
struct WarpMerger {
  uint64_t* stage;
  int n_new;
  int lane;
  uint64_t floor_key;      // k'-th best so far: offers at or below it are dropped before they reach the stage
  __device__ void init(uint64_t* s, int l) {
    stage = s; n_new = 0; lane = l; floor_key = 0ull;
    for (int i = lane; i < kCap; i += 32) stage[i] = 0ull;
    __syncwarp();
  }
  // stage[0..127] is sorted (or empty) and stage[128..] holds <= 128 new keys: sort only the new keys and merge
  // (one compare against the reversed tail + one bitonic merge phase), as in the scan's compaction.
  // The tail's sort network follows the number of new keys (a query of a sharded search gathers a few dozen).
  template <int TE>
  __device__ void sort_tail(uint64_t (&t)[kEpl / 2]) {
    uint64_t tt[TE];
#pragma unroll
    for (int i = 0; i < TE; ++i) tt[i] = t[i];
    warp_sort_desc<TE>(tt, lane);
#pragma unroll
    for (int i = 0; i < kEpl / 2; ++i) t[i] = i < TE ? tt[i < TE ? i : 0] : 0ull;
  }
  __device__ void flush() {
    constexpr int PE = kEpl / 2;
    if (n_new == 0) return;                          // (the prefix is sorted already)
    uint64_t p[PE], t[PE];
#pragma unroll
    for (int i = 0; i < PE; ++i) {
      p[i] = stage[(i << 5) | lane];
      t[i] = stage[kCap / 2 + ((i << 5) | lane)];
    }
    __syncwarp();
    if (n_new <= 32) sort_tail<1>(t);
    else if (n_new <= 64) sort_tail<2>(t);
    else sort_tail<PE>(t);
#pragma unroll
    for (int i = 0; i < PE; ++i) {
      const uint64_t r = __shfl_sync(kFull, t[PE - 1 - i], 31 - lane);
      p[i] = p[i] > r ? p[i] : r;
    }
    warp_bitonic_merge_desc<PE>(p, lane);
#pragma unroll
    for (int i = 0; i < PE; ++i) {
      stage[(i << 5) | lane] = p[i];
      stage[kCap / 2 + ((i << 5) | lane)] = 0ull;
    }
    floor_key = __shfl_sync(kFull, p[PE - 1], 31);
    n_new = 0;
    __syncwarp();
  }
  // top-128 of this merger's sorted prefix and another sorted 128-key list (pairwise tree merge)
  __device__ void absorb_sorted(const uint64_t* other) {
    constexpr int PE = kEpl / 2;
    uint64_t p[PE];
#pragma unroll
    for (int i = 0; i < PE; ++i) {
      const int e = (i << 5) | lane;
      const uint64_t r = other[kCap / 2 - 1 - e];
      const uint64_t a = stage[e];
      p[i] = a > r ? a : r;
    }
    __syncwarp();
    warp_bitonic_merge_desc<PE>(p, lane);
#pragma unroll
    for (int i = 0; i < PE; ++i) stage[(i << 5) | lane] = p[i];
    __syncwarp();
  }
  // every lane offers one key (0 = nothing)
  __device__ void offer(uint64_t key) {
    if (key <= floor_key) key = 0ull;
    const unsigned m = __ballot_sync(kFull, key != 0ull);
    if (m == 0) return;
    if (key != 0ull) stage[kCap / 2 + n_new + __popc(m & ((1u << lane) - 1u))] = key;
    n_new += __popc(m);
    __syncwarp();
    if (n_new + 32 > kCap / 2) flush();
  }
};

// ------------------------------------------------------------------ K4 + K6: merge, re-score, emit
// One CTA (4 warps) per query.  Phase 1 (all warps): gather every split's surviving
// candidates that are still >= the final shared threshold and keep the best k_sel.
// Phase 2 (all warps, optional): exact fp32 re-score of the survivors against the
// fp32 operand copy (gather of k_sel rows of k_pad floats).  Phase 3: final sort, emit.
// `verify` (exact-fp32 search on the tensor pipe): the scan ranked by split-bf16 scores (error <= delta, see below) and the
// best `rescore` candidates were re-scored in exact fp32.  Everything that was NOT re-scored has an approximate score of
// at most that of the last re-scored candidate, a_bound; if the k-th exact score exceeds a_bound + delta, no outsider can
// reach the top k and the row is proven exact.  Otherwise the row is flagged for the FFMA scan (row_flags / n_flagged).
struct VerifyParams {
  int32_t* row_flags;            // [n_queries] out: 1 = re-scan this row exactly
  int32_t* n_flagged;
  const unsigned* cat_norm_max;  // largest operand-row norm of the catalog (float bits)
  float rel;                     // delta = rel * |query operand| * cat_norm_max
};

template <bool kScatter>
__global__ void __launch_bounds__(128)
merge_emit_kernel(const ScanParams P, int k, int rescore, const float* __restrict__ q_f32,
                  const float* __restrict__ cat_f32, int as_l2, float* __restrict__ out_scores,
                  int32_t* __restrict__ out_ids, const __grid_constant__ CommParams C, const VerifyParams V) {
  __shared__ uint64_t stage_all[4][kCap];
  uint64_t* stage = stage_all[0];
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (P.row_active != nullptr && __ldg(P.row_active + q) == 0) return;      // fallback pass: flagged rows only

  // Phase 1: stream s of the row belongs to warp s % 4.  A warp takes its streams four at a time: the counts were fetched
  // lane-parallel, the first 32 keys of the four streams are loaded together (the walk is otherwise a chain of dependent
  // L2 round trips), and -- a unit leaves its buffer as a SORTED, published prefix -- a stream whose first group holds
  // nothing at or above the final threshold is finished (with ~60 streams per row a couple of keys per stream survive).
  {
    WarpMerger mg;
    mg.init(stage_all[warp], lane);
    const uint32_t thr = thr_load(P, q);
    const int n_slots = P.n_slot_splits * P.n_sub;
    for (int s0 = warp; s0 < n_slots; s0 += 128) {           // lane l of this pass: stream s0 + 4 * l
      const int my_s = s0 + 4 * lane;
      const int64_t my_slot = static_cast<int64_t>(my_s) * P.b_pad + q;
      const int my_cnt = my_s < n_slots ? __ldcg(P.cand_cnt + my_slot) : 0;
      const int my_sorted = (my_s < n_slots && P.pub_cnt != nullptr) ? __ldcg(P.pub_cnt + my_slot) : 0;
      const int n_here = min(32, (n_slots - s0 + 3) / 4);
      for (int j0 = 0; j0 < n_here; j0 += 4) {
        uint64_t first[4];
        int cnt[4], sorted[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          cnt[u] = __shfl_sync(kFull, my_cnt, (j0 + u) & 31);
          sorted[u] = __shfl_sync(kFull, my_sorted, (j0 + u) & 31);
          if (j0 + u >= n_here) cnt[u] = 0;
          const uint64_t* buf = P.cand + (static_cast<int64_t>(s0 + 4 * (j0 + u)) * P.b_pad + q) * P.cap;
          first[u] = lane < cnt[u] ? __ldcg(buf + lane) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (cnt[u] == 0) continue;
          const uint64_t* buf = P.cand + (static_cast<int64_t>(s0 + 4 * (j0 + u)) * P.b_pad + q) * P.cap;
          uint64_t key = first[u];
          for (int base = 0; base < cnt[u]; base += 32) {
            if (base > 0) key = (base + lane < cnt[u]) ? __ldcg(buf + base + lane) : 0ull;
            if (static_cast<uint32_t>(key >> 32) < thr) key = 0ull;
            const bool any = __any_sync(kFull, key != 0ull);
            mg.offer(key);
            if (!any && base + 32 <= sorted[u]) base = (sorted[u] > base + 32 ? sorted[u] : base + 32) - 32;   // rest of the sorted prefix: below the threshold
          }
        }
      }
    }
    mg.flush();
  }
  __syncthreads();
  // pairwise tree merge of the four sorted partial tops: (0,1) and (2,3) in parallel, then (0,2)
  if ((warp == 0 || warp == 2) && stage_all[warp + 1][0] != 0ull) {
    WarpMerger m2;
    m2.stage = stage_all[warp];
    m2.lane = lane;
    m2.absorb_sorted(stage_all[warp + 1]);
  }
  __syncthreads();
  if (warp == 0 && stage_all[2][0] != 0ull) {
    WarpMerger m2;
    m2.stage = stage_all[0];
    m2.lane = lane;
    m2.absorb_sorted(stage_all[2]);
  }
  __syncthreads();

  if (rescore > 0) {
    // `rescore` = number of best candidates (by bf16 score) that are re-scored; the rest of the over-fetch only existed
    // because buffers are kept in multiples of 32 and is dropped here.
    const uint64_t bound_key = stage[rescore - 1];     // (read by every thread before anything is overwritten)
    __syncthreads();
    for (int c = rescore + threadIdx.x; c < kMaxSel; c += blockDim.x) stage[c] = 0ull;
    __syncthreads();
    // survivors (sorted, zeros at the end): a shard of a collective search keeps only what beats the GLOBAL bound
    int n_cand = 0;
    for (int c = 0; c < rescore; c += 32) n_cand += __popc(__ballot_sync(kFull, c + lane < rescore && stage[c + lane] != 0ull));
    // exact fp32 dot of the query operand with each surviving catalog row.  Eight lanes per candidate, four
    // candidates per warp at a time: every lane keeps up to 14 independent 16-byte gathers in flight, which is what
    // hides the latency of these scattered row reads.
    const float4* qv = reinterpret_cast<const float4*>(q_f32 + static_cast<int64_t>(q) * P.k_pad);
    const int n4 = P.k_pad >> 2;
    const int sub = lane >> 3, l8 = lane & 7;
    for (int c0 = warp * 4; c0 < n_cand; c0 += 16) {
      const int c = c0 + sub;
      const uint64_t key = c < n_cand ? stage[c] : 0ull;
      float acc = 0.f;
      int32_t gid = -1;
      if (key != 0ull) {
        gid = key_id(key);
        const float4* ev = reinterpret_cast<const float4*>(cat_f32 + (static_cast<int64_t>(gid) - P.id_offset) * P.k_pad);
#pragma unroll 4
        for (int i = l8; i < n4; i += 8) {
          const float4 a = __ldg(qv + i), b = __ldg(ev + i);
          acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
        }
      }
      acc += __shfl_xor_sync(kFull, acc, 4);
      acc += __shfl_xor_sync(kFull, acc, 2);
      acc += __shfl_xor_sync(kFull, acc, 1);
      __syncwarp();
      if (l8 == 0 && key != 0ull) stage[c] = make_key(acc, static_cast<uint32_t>(gid));
    }
    __syncthreads();
    if (warp == 0) {
      if (n_cand <= 32) {
        uint64_t v[1] = {stage[lane]};
        __syncwarp();
        warp_sort_desc<1>(v, lane);
        stage[lane] = v[0];
      } else if (n_cand <= 64) {
        uint64_t v[2] = {stage[lane], stage[32 + lane]};
        __syncwarp();
        warp_sort_desc<2>(v, lane);
        stage[lane] = v[0];
        stage[32 + lane] = v[1];
      } else {
        uint64_t v[kEpl / 2];
#pragma unroll
        for (int i = 0; i < kEpl / 2; ++i) v[i] = stage[(i << 5) | lane];
        __syncwarp();
        warp_sort_desc<kEpl / 2>(v, lane);
#pragma unroll
        for (int i = 0; i < kEpl / 2; ++i) stage[(i << 5) | lane] = v[i];
      }
      if (V.row_flags != nullptr) {
        // proof of completeness (see VerifyParams)
        float ss = 0.f;
        for (int i = lane; i < P.k_pad; i += 32) {
          const float x = __ldg(q_f32 + static_cast<int64_t>(q) * P.k_pad + i);
          ss = fmaf(x, x, ss);
        }
        ss = warp_sum(ss);
        __syncwarp();
        if (lane == 0) {
          bool flagged = false;
          if (bound_key != 0ull) {
            const uint64_t kth = stage[k - 1];
            const float delta = V.rel * sqrtf(ss) * __uint_as_float(__ldg(V.cat_norm_max)) + 1e-30f;
            flagged = kth == 0ull || !(key_score(kth) > key_score(bound_key) + delta);
          }
          V.row_flags[q] = flagged ? 1 : 0;
          if (flagged) atomicAdd(V.n_flagged, 1);
        }
      }
    }
    __syncthreads();
  }

  if constexpr (kScatter) {
    // row-sharded catalog: this rank's k best keys of query q go straight into the OWNER's inbox (peer stores over
    // NVLink when the owner is another GPU), then one arrival is counted there
    const int owner = q / C.slice;
    unsigned long long* dst = C.inbox[owner] + inbox_offset(C, C.parity, C.rank, q - owner * C.slice);
    for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = stage[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd_system(C.arrivals[owner] + C.parity, 1ull);
    return;
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = stage[i];
    float s = -INFINITY;
    int32_t id = -1;
    if (key != 0ull) {
      s = key_score(key);
      id = key_id(key);
      if (as_l2) s = 2.f - 2.f * s;
    } else if (as_l2) {
      s = INFINITY;
    }
    out_scores[static_cast<int64_t>(q) * k + i] = s;
    out_ids[static_cast<int64_t>(q) * k + i] = id;
  }
}

// ------------------------------------------------------------------ owner-side merge of a collective search
// One warp per query of this rank's slice: wait until every rank's keys of the slice have arrived (cumulative
// counter, bounded wait), merge the `world` lists, emit.
__global__ void __launch_bounds__(128)
merge_inbox_kernel(const __grid_constant__ CommParams C, int n_mine, int k, unsigned long long expected, int as_l2,
                   float* __restrict__ out_scores, int32_t* __restrict__ out_ids, int32_t* err_flag) {
  __shared__ uint64_t stage_all[4][kCap];
  __shared__ int s_ok;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const unsigned long long* cnt = C.arrivals[C.rank] + C.parity;
    int ok = 1;
    const long long t0 = clock64();
    while (true) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(cnt) : "memory");
      if (v >= expected) break;
      __nanosleep(500);
      if (clock64() - t0 > 6000000000ll) {        // ~3 s: a peer died; fail the call instead of hanging the GPU
        atomicExch(err_flag, 108);
        ok = 0;
        break;
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  if (!s_ok) return;
  const int q = blockIdx.x * 4 + warp;
  if (q >= n_mine) return;
  uint64_t* stage = stage_all[warp];
  WarpMerger mg;
  mg.init(stage, lane);
  for (int r = 0; r < C.world; ++r) {
    const unsigned long long* src = C.inbox[C.rank] + inbox_offset(C, C.parity, r, q);
    for (int i = 0; i < k; i += 32) mg.offer(i + lane < k ? __ldcg(src + i + lane) : 0ull);
  }
  mg.flush();
  for (int i = lane; i < k; i += 32) {
    const uint64_t key = stage[i];
    float s = key ? key_score(key) : -INFINITY;
    if (as_l2) s = key ? 2.f - 2.f * s : INFINITY;
    out_scores[static_cast<int64_t>(q) * k + i] = s;
    out_ids[static_cast<int64_t>(q) * k + i] = key ? key_id(key) : -1;
  }
}

// ------------------------------------------------------------------ threshold seeding from the pilot pass
// One warp per query row: the pilot pass left the row's best valid score of every 8-column group of the leading
// pilot_items rows.  Group maxima are scores of DISTINCT valid items, except that the query's own exclusion lists were
// not applied: with E listed ids inside the pilot range, the (m + E)-th largest maximum is a lower bound of the row's
// m-th best valid score.  m = k' seeds the threshold word, m = k'/8, k'/4, k'/2 this rank's level claims.
struct WarpMerger32 {
  uint32_t* stage;     // [256]: [0,128) sorted best so far, [128,256) new
  int n_new, lane;
  uint32_t floor_key;
  __device__ void init(uint32_t* s, int l) {
    stage = s; n_new = 0; lane = l; floor_key = 0u;
    for (int i = lane; i < 256; i += 32) stage[i] = 0u;
    __syncwarp();
  }
  // stage[0..127] sorted (or empty), stage[128..] <= 128 new keys: sort the new keys, keep the best 128 of the union
  // (compare against the reversed tail + one bitonic merge phase)
  __device__ void flush() {
    if (n_new == 0) return;
    uint32_t p[4], t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      p[i] = stage[(i << 5) | lane];
      t[i] = stage[128 + ((i << 5) | lane)];
    }
    __syncwarp();
    warp_sort_desc<4, uint32_t>(t, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t r = __shfl_sync(kFull, t[3 - i], 31 - lane);
      p[i] = p[i] > r ? p[i] : r;
    }
    warp_bitonic_merge_desc<4, uint32_t>(p, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      stage[(i << 5) | lane] = p[i];
      stage[128 + ((i << 5) | lane)] = 0u;
    }
    floor_key = __shfl_sync(kFull, p[3], 31);
    n_new = 0;
    __syncwarp();
  }
  __device__ void offer(uint32_t key) {
    if (key <= floor_key) key = 0u;
    const unsigned m = __ballot_sync(kFull, key != 0u);
    if (m == 0) return;
    if (key != 0u) stage[128 + n_new + __popc(m & ((1u << lane) - 1u))] = key;
    n_new += __popc(m);
    __syncwarp();
    if (n_new + 32 > 128) flush();
  }
};

__device__ __forceinline__ int64_t lower_bound_ids(const int32_t* ids, int64_t lo, int64_t hi, int64_t gid) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(__ldg(ids + mid)) < gid) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128)
seed_thresholds_kernel(const ScanParams P, int64_t pilot_items) {
  __shared__ uint32_t stage_all[4][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 4 + warp;
  if (row >= P.n_queries) return;
  WarpMerger32 mg;
  mg.init(stage_all[warp], lane);
  const float* src = P.pilot_max + static_cast<int64_t>(row) * P.pilot_groups;
  for (int g = 0; g < P.pilot_groups; g += 32) {
    uint32_t key = 0u;
    if (g + lane < P.pilot_groups) {
      key = f2ord(__ldcg(src + g + lane));
      if (key <= kOrdNegInf) key = 0u;          // (no valid item in the group)
    }
    mg.offer(key);
  }
  mg.flush();
  // listed ids inside the pilot range (they may sit behind up to E of the maxima)
  // (a single list per query is walked exactly by the pilot's epilogue -- the pilot only runs on the variant that walks
  // lists: nothing to allow for)
  int64_t excl = 0;
  if (lane == 0 && P.q_excl_n > 1) {
    for (int e = 0; e < P.q_excl_n; ++e) {
      const int32_t l = __ldg(P.q_excl + static_cast<int64_t>(row) * P.q_excl_n + e);
      if (l < 0) continue;
      const int64_t lo = __ldg(P.excl_off + l), hi = __ldg(P.excl_off + l + 1);
      excl += lower_bound_ids(P.excl_ids, lo, hi, P.id_offset + pilot_items) - lower_bound_ids(P.excl_ids, lo, hi, P.id_offset);
    }
  }
  excl = __shfl_sync(kFull, excl, 0);
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (l < 3 && P.lvl_stride == 0) continue;
      const int64_t e = (P.k_sel >> (3 - l)) + excl - 1;
      if (e >= 128) continue;
      const uint32_t v = mg.stage[e];
      if (v <= kOrdNegInf) continue;
      if (l == 3) thr_raise(P, row, v); else lvl_raise(P, row, l, v);
    }
  }
}

// ------------------------------------------------------------------ cross-shard merge of (score, id) lists
__global__ void __launch_bounds__(128)
merge_lists_kernel(int n_queries, int n_lists, int len, const float* __restrict__ scores,
                   const int32_t* __restrict__ ids, int k, float* __restrict__ out_scores,
                   int32_t* __restrict__ out_ids) {
  __shared__ uint64_t stage_all[4][kCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 4 + warp;
  if (q >= n_queries) return;
  uint64_t* stage = stage_all[warp];
  WarpMerger mg;
  mg.init(stage, lane);
  for (int l = 0; l < n_lists; ++l) {
    const int64_t base = (static_cast<int64_t>(l) * n_queries + q) * len;
    for (int i = 0; i < len; i += 32) {
      uint64_t key = 0ull;
      if (i + lane < len) {
        const int32_t id = __ldg(ids + base + i + lane);
        if (id >= 0) key = make_key(__ldg(scores + base + i + lane), static_cast<uint32_t>(id));
      }
      mg.offer(key);
    }
  }
  mg.flush();
  for (int i = lane; i < k; i += 32) {
    const uint64_t key = stage[i];
    out_scores[static_cast<int64_t>(q) * k + i] = key ? key_score(key) : -INFINITY;
    out_ids[static_cast<int64_t>(q) * k + i] = key ? key_id(key) : -1;
  }
}

// ------------------------------------------------------------------ K5: reference-semantics hybrid combine
// recommendation_system.py:789-843 -- union of the two lists, missing side = 0,
// python-float (fp64) blend without FMA contraction, rank by (-score, +id).
constexpr int kHybMax = 512;
__global__ void __launch_bounds__(128)
hybrid_combine_kernel(int len_c, const int32_t* __restrict__ c_ids, const float* __restrict__ c_scores, int len_f,
                      const int32_t* __restrict__ f_ids, const float* __restrict__ f_scores,
                      const double* __restrict__ weights, int weights_per_row, int k, double* __restrict__ out_scores,
                      int32_t* __restrict__ out_ids) {
  __shared__ int32_t u_id[kHybMax];
  __shared__ double u_sc[kHybMax];
  __shared__ int n_union;
  const int row = blockIdx.x, tid = threadIdx.x;
  const double wc = weights[weights_per_row ? 2 * row : 0], wf = weights[weights_per_row ? 2 * row + 1 : 1];
  const int32_t* ci = c_ids + static_cast<int64_t>(row) * len_c;
  const float* cs = c_scores + static_cast<int64_t>(row) * len_c;
  const int32_t* fi = f_ids + static_cast<int64_t>(row) * len_f;
  const float* fs = f_scores + static_cast<int64_t>(row) * len_f;
  if (tid == 0) n_union = 0;
  __syncthreads();
  for (int i = tid; i < len_c; i += blockDim.x) {
    const int32_t id = ci[i];
    if (id < 0) continue;
    double f = 0.0;
    for (int j = 0; j < len_f; ++j)
      if (fi[j] == id) f = static_cast<double>(fs[j]);
    const int slot = atomicAdd(&n_union, 1);
    u_id[slot] = id;
    u_sc[slot] = __dadd_rn(__dmul_rn(wc, static_cast<double>(cs[i])), __dmul_rn(wf, f));
  }
  for (int j = tid; j < len_f; j += blockDim.x) {
    const int32_t id = fi[j];
    if (id < 0) continue;
    bool in_c = false;
    for (int i = 0; i < len_c; ++i) in_c |= (ci[i] == id);
    if (in_c) continue;
    const int slot = atomicAdd(&n_union, 1);
    u_id[slot] = id;
    u_sc[slot] = __dadd_rn(__dmul_rn(wc, 0.0), __dmul_rn(wf, static_cast<double>(fs[j])));
  }
  __syncthreads();
  const int n = n_union;
  for (int i = tid; i < k; i += blockDim.x) {
    out_scores[static_cast<int64_t>(row) * k + i] = -INFINITY;
    out_ids[static_cast<int64_t>(row) * k + i] = -1;
  }
  __syncthreads();
  for (int e = tid; e < n; e += blockDim.x) {
    const double s = u_sc[e];
    const int32_t id = u_id[e];
    int rank = 0;
    for (int o = 0; o < n; ++o) {
      const double so = u_sc[o];
      rank += (so > s) || (so == s && u_id[o] < id);
    }
    if (rank < k) {
      out_scores[static_cast<int64_t>(row) * k + rank] = s;
      out_ids[static_cast<int64_t>(row) * k + rank] = id;
    }
  }
}

inline int blocks_for(int64_t n, int per_block) { return static_cast<int>((n + per_block - 1) / per_block); }

}  // namespace

cudaError_t launch_prep_catalog(const float* emb, const float* cf, int64_t n_rows, int dim, int cf_dim, int k_pad,
                                int metric, float* out_f32, void* out_bf16, void* out_lo, unsigned* norm_max, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  prep_catalog_kernel<<<blocks_for(n_rows, 8), 256, 0, stream>>>(emb, cf, n_rows, dim, cf_dim, k_pad, metric, out_f32,
                                                                static_cast<__nv_bfloat16*>(out_bf16),
                                                                static_cast<__nv_bfloat16*>(out_lo), norm_max);
  return cudaGetLastError();
}

cudaError_t launch_prep_queries(const float* q_emb, const int32_t* q_item, const float* q_cf, const float* w_c,
                                const float* w_f, float w_c_all, float w_f_all, int n_queries, int b_pad, int dim, int cf_dim,
                                int k_pad, int metric, const float* cat_f32, const void* cat_bf16, int64_t id_offset,
                                int64_t n_items, float* out_f32, void* out_bf16, void* out_lo, int32_t* err_flag,
                                cudaStream_t stream) {
  prep_queries_kernel<<<blocks_for(b_pad, 8), 256, 0, stream>>>(
      q_emb, q_item, q_cf, w_c, w_f, w_c_all, w_f_all, n_queries, b_pad, dim, cf_dim, k_pad, metric, cat_f32,
      static_cast<const __nv_bfloat16*>(cat_bf16), id_offset, n_items, out_f32, static_cast<__nv_bfloat16*>(out_bf16),
      static_cast<__nv_bfloat16*>(out_lo), err_flag);
  return cudaGetLastError();
}

cudaError_t launch_pack_meta(const int32_t* num_parts, const int16_t* year, const int16_t* theme, int64_t n_rows,
                             int2* out, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  pack_meta_kernel<<<blocks_for(n_rows, 256), 256, 0, stream>>>(num_parts, year, theme, n_rows, out);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* cat_f32, const void* cat_bf16, int k_pad, int dim, int64_t id_offset,
                               int64_t n_items, const int32_t* ids, int n, float* out, int32_t* err_flag,
                               cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  gather_rows_kernel<<<blocks_for(n, 8), 256, 0, stream>>>(cat_f32, static_cast<const __nv_bfloat16*>(cat_bf16), k_pad,
                                                          dim, id_offset, n_items, ids, n, out, err_flag);
  return cudaGetLastError();
}

cudaError_t launch_mask(const ScanParams& P, int n_constraints, uint32_t* out_bitmaps, unsigned long long* out_counts,
                        cudaStream_t stream) {
  const int64_t n_words = (P.n_items + 31) >> 5;
  int bx = static_cast<int>((n_words + 7) / 8);
  if (bx > 148 * 8) bx = 148 * 8;       // grid-stride: 8 resident CTAs per SM
  if (bx < 1) bx = 1;
  for (int c0 = 0; c0 < n_constraints; c0 += kMaskCons) {
    const int n = n_constraints - c0 < kMaskCons ? n_constraints - c0 : kMaskCons;
    mask_kernel<<<bx, 256, 0, stream>>>(P, c0, n, out_bitmaps, out_counts);
  }
  return cudaGetLastError();
}

cudaError_t launch_merge_emit(const ScanParams& P, int k, int rescore, const float* q_f32, const float* cat_f32,
                              bool as_l2, float* out_scores, int32_t* out_ids, const CommParams* comm, cudaStream_t stream,
                              int32_t* row_flags, int32_t* n_flagged, const unsigned* cat_norm_max) {
  VerifyParams V{};
  V.row_flags = row_flags;
  V.n_flagged = n_flagged;
  V.cat_norm_max = cat_norm_max;
  V.rel = 1.0f / 4096.0f;          // 2^-12: operand split (<= 1.2e-5), tensor-pipe and re-score accumulation (K <= 1344 terms)
  if (comm != nullptr) {
    merge_emit_kernel<true><<<P.n_queries, 128, 0, stream>>>(P, k, rescore, q_f32, cat_f32, 0, nullptr, nullptr, *comm, V);
  } else {
    CommParams none{};
    merge_emit_kernel<false><<<P.n_queries, 128, 0, stream>>>(P, k, rescore, q_f32, cat_f32, as_l2 ? 1 : 0, out_scores,
                                                              out_ids, none, V);
  }
  return cudaGetLastError();
}

cudaError_t launch_merge_inbox(const CommParams& comm, int n_mine, int k, unsigned long long expected, bool as_l2,
                               float* out_scores, int32_t* out_ids, int32_t* err_flag, cudaStream_t stream) {
  if (n_mine <= 0) return cudaSuccess;
  merge_inbox_kernel<<<blocks_for(n_mine, 4), 128, 0, stream>>>(comm, n_mine, k, expected, as_l2 ? 1 : 0, out_scores, out_ids,
                                                                err_flag);
  return cudaGetLastError();
}

cudaError_t launch_seed_thresholds(const ScanParams& P, int64_t pilot_items, cudaStream_t stream) {
  seed_thresholds_kernel<<<blocks_for(P.n_queries, 4), 128, 0, stream>>>(P, pilot_items);
  return cudaGetLastError();
}

cudaError_t launch_merge_lists(int n_queries, int n_lists, int len, const float* scores, const int32_t* ids, int k,
                               float* out_scores, int32_t* out_ids, cudaStream_t stream) {
  merge_lists_kernel<<<blocks_for(n_queries, 4), 128, 0, stream>>>(n_queries, n_lists, len, scores, ids, k, out_scores,
                                                                  out_ids);
  return cudaGetLastError();
}

cudaError_t launch_hybrid_combine(int n_rows, int len_c, const int32_t* c_ids, const float* c_scores, int len_f,
                                  const int32_t* f_ids, const float* f_scores, const double* weights,
                                  int weights_per_row, int k, double* out_scores, int32_t* out_ids,
                                  cudaStream_t stream) {
  hybrid_combine_kernel<<<n_rows, 128, 0, stream>>>(len_c, c_ids, c_scores, len_f, f_ids, f_scores, weights,
                                                    weights_per_row, k, out_scores, out_ids);
  return cudaGetLastError();
}

}  // namespace bb
