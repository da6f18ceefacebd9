// Exact fp32 scoring for SMALL batches (B <= kSmallBatch): the reference's own call pattern is one query per call
// (get_similar_sets / get_recommendations, recommendation_system.py:194-247, 411-481), where the FFMA tile scan of
// bb_scan_fp32.cu spends its time on 127 padded query rows.  Here the catalog is streamed ONCE at HBM speed:
//
//   K_a score_small_kernel:  eight lanes per catalog row (float4 loads, 128 B contiguous per row and step), the B
//        query operands in shared memory; scores[q][item] = <query operand q, catalog operand item> (fp32 FMA,
//        three shuffles per (row, query)).  Bound: HBM, 4*k_pad B per item (1792 B at 384+64); the dense score
//        strip written back is B*4 B per item.
//   K_b select_small_kernel: one CTA per (catalog split, query): scores above the running k'-th best are staged and
//        validated in batches at the merger's flush (bounds, self, hard constraints, allow / exclusion lists through
//        item_ok); the split's sorted top-k' goes to the query's candidate slot, and the regular merge_emit kernel finishes.
#include "bb_common.cuh"
#include "bb_kernels.h"

namespace bb {
namespace {

constexpr int kRowsPerGroup = 2;                       // catalog rows an 8-lane group scores per query-operand read
constexpr int kItemsPerWarp = 4 * kRowsPerGroup;       // 8 lanes per catalog row, 4 groups per warp

template <int B>
__global__ void __launch_bounds__(256) score_small_kernel(const float* __restrict__ q_f32, const float* __restrict__ cat_f32,
                                                          int64_t n_items, int k_pad, float* __restrict__ scores) {
  extern __shared__ float4 s_q[];   // [B][k_pad / 4]
  const int n4 = k_pad >> 2;
  for (int i = threadIdx.x; i < B * n4; i += blockDim.x) s_q[i] = reinterpret_cast<const float4*>(q_f32)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t n_groups = (n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  for (int64_t g = warp0; g < n_groups; g += n_warps) {
    // the group's rows: item0 + r * 4 (so that the four groups of a warp touch 4 consecutive rows per step)
    const int64_t item0 = g * kItemsPerWarp + sub;
    float acc[kRowsPerGroup][B];
#pragma unroll
    for (int r = 0; r < kRowsPerGroup; ++r)
#pragma unroll
      for (int q = 0; q < B; ++q) acc[r][q] = 0.f;
#pragma unroll 2
    for (int i = l8; i < n4; i += 8) {
      float4 e[kRowsPerGroup];
#pragma unroll
      for (int r = 0; r < kRowsPerGroup; ++r) {
        const int64_t item = item0 + r * 4;
        // streamed once: do not keep it in L2 ahead of the score strip
        e[r] = item < n_items ? __ldcs(reinterpret_cast<const float4*>(cat_f32 + item * k_pad) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < B; ++q) {
        const float4 a = s_q[q * n4 + i];                  // one shared-memory read serves kRowsPerGroup rows
#pragma unroll
        for (int r = 0; r < kRowsPerGroup; ++r) {
          acc[r][q] = fmaf(a.x, e[r].x, acc[r][q]);
          acc[r][q] = fmaf(a.y, e[r].y, acc[r][q]);
          acc[r][q] = fmaf(a.z, e[r].z, acc[r][q]);
          acc[r][q] = fmaf(a.w, e[r].w, acc[r][q]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerGroup; ++r) {
      const int64_t item = item0 + r * 4;
#pragma unroll
      for (int q = 0; q < B; ++q) {
        float v = acc[r][q];
        v += __shfl_xor_sync(kFull, v, 4);
        v += __shfl_xor_sync(kFull, v, 2);
        v += __shfl_xor_sync(kFull, v, 1);
        if (l8 == 0 && item < n_items) scores[static_cast<int64_t>(q) * n_items + item] = v;
      }
    }
  }
}

// one warp's running top-kMaxSel (same staging scheme as merge_emit's WarpMerger, kept local to this file)
struct SmallMerger {
  uint64_t* stage;   // [kCap]: [0, kCap/2) sorted best so far, [kCap/2, kCap) new keys
  int n_new, lane;
  uint64_t floor_key;
  __device__ void init(uint64_t* s, int l) {
    stage = s; n_new = 0; lane = l; floor_key = 0ull;
    for (int i = lane; i < kCap; i += 32) stage[i] = 0ull;
    __syncwarp();
  }
  // `validate`: the new keys were staged unchecked; every lane checks its own (bounds, self, hard constraints, allow /
  // exclusion lists) NOW, four at a time and all lanes busy -- checking at offer time made every loop iteration with a
  // passing score wait for a chain of dependent global loads (self id, descriptor, metadata, list search)
  __device__ void flush(const ScanParams& P, int q, bool validate) {
    uint64_t v[kEpl];
#pragma unroll
    for (int i = 0; i < kEpl; ++i) v[i] = stage[(i << 5) | lane];
    if (validate) {
#pragma unroll
      for (int i = kEpl / 2; i < kEpl; ++i)
        if (v[i] != 0ull && !item_ok(P, q, static_cast<int64_t>(key_id(v[i])) - P.id_offset)) v[i] = 0ull;
    }
    __syncwarp();
    warp_sort_desc<kEpl>(v, lane);
#pragma unroll
    for (int i = 0; i < kEpl; ++i) stage[(i << 5) | lane] = (i < kEpl / 2) ? v[i] : 0ull;
    floor_key = __shfl_sync(kFull, v[kEpl / 2 - 1], 31);
    n_new = 0;
    __syncwarp();
  }
  __device__ void offer(uint64_t key, const ScanParams& P, int q, bool validate) {
    const unsigned m = __ballot_sync(kFull, key != 0ull);
    if (m == 0) return;
    if (key != 0ull) stage[kCap / 2 + n_new + __popc(m & ((1u << lane) - 1u))] = key;
    n_new += __popc(m);
    __syncwarp();
    if (n_new + 32 > kCap / 2) flush(P, q, validate);
  }
};

__global__ void __launch_bounds__(128) select_small_kernel(const ScanParams P, const float* __restrict__ scores) {
  __shared__ uint64_t stages[4][kCap];
  const int split = blockIdx.x, q = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t per = (P.n_items + P.n_slot_splits - 1) / P.n_slot_splits;
  const int64_t lo = static_cast<int64_t>(split) * per, hi = min(P.n_items, lo + per);
  const float* row = scores + static_cast<int64_t>(q) * P.n_items;
  SmallMerger mg;
  mg.init(stages[warp], lane);
  // (eight independent loads in flight per lane: the loop is otherwise a chain of L2 round trips)
  constexpr int kUnroll = 8;
  for (int64_t base = lo + static_cast<int64_t>(warp) * 32 * kUnroll; base < hi; base += 128 * kUnroll) {
    float v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t item = base + j * 32 + lane;
      v[j] = item < hi ? __ldcs(row + item) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t item = base + j * 32 + lane;
      uint64_t key = 0ull;
      if (item < hi) {
        key = make_key(v[j], static_cast<uint32_t>(item + P.id_offset));
        if (key <= mg.floor_key) key = 0ull;       // cannot enter the top-k' any more; survivors are validated at the flush
      }
      mg.offer(key, P, q, true);
    }
  }
  mg.flush(P, q, true);
  __syncthreads();
  if (warp == 0) {
    // fold the other warps' sorted halves in, then publish the split's top-k' for merge_emit
    for (int w2 = 1; w2 < 4; ++w2) {
      for (int x = 0; x < kCap / 2; x += 32) mg.offer(stages[w2][x + lane], P, q, false);   // (already validated)
    }
    mg.flush(P, q, false);
    const int64_t slot = static_cast<int64_t>(split) * P.b_pad + q;
    uint64_t* out = P.cand + slot * P.cap;
    int cnt = 0;
    for (int x = 0; x < P.k_sel; x += 32) {
      const uint64_t key = stages[0][x + lane];
      out[x + lane] = key;
      cnt += __popc(__ballot_sync(kFull, key != 0ull));
    }
    if (lane == 0) P.cand_cnt[slot] = cnt;
  }
}

template <int B>
cudaError_t launch_score_b(const float* q_f32, const float* cat_f32, int64_t n_items, int k_pad, float* scores, int sm_count,
                           cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(B) * k_pad * sizeof(float);
  const int64_t groups = (n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  const int grid = static_cast<int>(max(static_cast<int64_t>(1), min((groups + 7) / 8, static_cast<int64_t>(sm_count) * 8)));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(score_small_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  score_small_kernel<B><<<grid, 256, smem, stream>>>(q_f32, cat_f32, n_items, k_pad, scores);
  return cudaGetLastError();
}

}  // namespace

int small_scan_splits(int64_t n_items) {
  int64_t s = (n_items + 1023) / 1024;          // the selection is a latency-bound single-CTA loop: keep it short (one
  if (s < 1) s = 1;                             // pass of 8 loads per thread at catalog size)
  if (s > 128) s = 128;
  return static_cast<int>(s);
}

cudaError_t launch_scan_small(const ScanParams& P, const float* q_f32, const float* cat_f32, float* scores, int sm_count,
                              cudaStream_t stream) {
  cudaError_t e;
  const int B = P.n_queries;
  if (B <= 1) e = launch_score_b<1>(q_f32, cat_f32, P.n_items, P.k_pad, scores, sm_count, stream);
  else if (B <= 2) e = launch_score_b<2>(q_f32, cat_f32, P.n_items, P.k_pad, scores, sm_count, stream);
  else if (B <= 4) e = launch_score_b<4>(q_f32, cat_f32, P.n_items, P.k_pad, scores, sm_count, stream);
  else if (B <= 8) e = launch_score_b<8>(q_f32, cat_f32, P.n_items, P.k_pad, scores, sm_count, stream);
  else e = launch_score_b<kSmallBatch>(q_f32, cat_f32, P.n_items, P.k_pad, scores, sm_count, stream);
  if (e != cudaSuccess) return e;
  select_small_kernel<<<dim3(P.n_slot_splits, B), 128, 0, stream>>>(P, scores);
  return cudaGetLastError();
}

}  // namespace bb
