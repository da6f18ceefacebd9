// K1x -- exact fp32 scan: FFMA tile GEMM with the fused mask + running-threshold
// top-k epilogue.  This is the parity path (BB_PREC_FP32): scores are plain fp32
// FMAs over the fp32 operand copy, bit-comparable with the reference's
// cosine_similarity / np.dot to summation-order noise (recommendation_system.py:213-214, :438).
//
// One CTA per (query block of 128 rows, catalog split); tiles of 128 items.
// Roofline: fp32 FFMA-bound (2*128*128*K flop per tile vs 128*K*4 B of catalog).
#include "bb_common.cuh"
#include "bb_kernels.h"

namespace bb {

namespace {
constexpr int XM = 128;      // query rows per CTA
constexpr int XN = 128;      // catalog rows per tile
constexpr int XK = 16;       // k-slab
constexpr int XLD = XM + 4;  // smem leading dim of the transposed slabs
constexpr int SLD = XN + 4;  // smem leading dim of the score tile

struct Smem {
  float a[2][XK][XLD];
  float b[2][XK][XLD];
  float s[XM][SLD];
};

__device__ __forceinline__ void load_slab(const float* __restrict__ base, int64_t row, int64_t n_rows, int k_pad,
                                          int k0, float4& r0, float4& r1) {
  if (row < n_rows) {
    const float4* p = reinterpret_cast<const float4*>(base + row * k_pad + k0);
    r0 = __ldg(p);
    r1 = __ldg(p + 1);
  } else {
    r0 = make_float4(0.f, 0.f, 0.f, 0.f);
    r1 = r0;
  }
}

__device__ __forceinline__ void store_slab(float (*dst)[XLD], int row, int kq, const float4& r0, const float4& r1) {
  dst[kq + 0][row] = r0.x; dst[kq + 1][row] = r0.y; dst[kq + 2][row] = r0.z; dst[kq + 3][row] = r0.w;
  dst[kq + 4][row] = r1.x; dst[kq + 5][row] = r1.y; dst[kq + 6][row] = r1.z; dst[kq + 7][row] = r1.w;
}
}  // namespace

__global__ void __launch_bounds__(256, 2)
scan_fp32_kernel(const ScanParams P, const float* __restrict__ Q, const float* __restrict__ E) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int unit = blockIdx.x;
  const int qb = unit % P.n_qblocks, split = unit / P.n_qblocks;
  int t0, t1;
  split_tile_range(P, split, t0, t1);

  const int ld_row = tid >> 1, ld_kq = (tid & 1) * 8;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t q_row0 = static_cast<int64_t>(qb) * XM;

  // row-owner state (warps 0..3 own rows warp*32 + lane)
  const int own_row = static_cast<int>(q_row0) + warp * 32 + lane;
  // fallback pass of the tensor-pipe exact search: only flagged rows are scanned; a block without one leaves at once
  bool row_on = warp < 4 && own_row < P.n_queries;
  if (P.row_active != nullptr) {
    row_on = row_on && __ldg(P.row_active + own_row) != 0;
    if (!__syncthreads_or(row_on ? 1 : 0)) return;
  }
  RowState st;
  st.thr = row_on ? -INFINITY : INFINITY;
  st.cnt = 0;
  st.sorted = 0;
  st.slot = static_cast<int64_t>(split) * P.n_sub * P.b_pad + (warp < 4 ? own_row : 0);
  st.buf = P.cand + st.slot * P.cap;

  const int n_slabs = P.k_pad / XK;
  for (int t = t0; t < t1; ++t) {
    const int64_t item0 = static_cast<int64_t>(t) * XN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra0, ra1, rb0, rb1;
    load_slab(Q, q_row0 + ld_row, P.b_pad, P.k_pad, ld_kq, ra0, ra1);
    load_slab(E, item0 + ld_row, P.n_items, P.k_pad, ld_kq, rb0, rb1);
    store_slab(sm.a[0], ld_row, ld_kq, ra0, ra1);
    store_slab(sm.b[0], ld_row, ld_kq, rb0, rb1);
    __syncthreads();

    for (int s = 0; s < n_slabs; ++s) {
      const int cur = s & 1;
      if (s + 1 < n_slabs) {
        load_slab(Q, q_row0 + ld_row, P.b_pad, P.k_pad, (s + 1) * XK + ld_kq, ra0, ra1);
        load_slab(E, item0 + ld_row, P.n_items, P.k_pad, (s + 1) * XK + ld_kq, rb0, rb1);
      }
#pragma unroll
      for (int k = 0; k < XK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[cur][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[cur][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[cur][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[cur][k][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (s + 1 < n_slabs) {
        store_slab(sm.a[cur ^ 1], ld_row, ld_kq, ra0, ra1);
        store_slab(sm.b[cur ^ 1], ld_row, ld_kq, rb0, rb1);
      }
      __syncthreads();
    }

    // score tile -> smem [row][col]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
      *reinterpret_cast<float4*>(&sm.s[r][tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(&sm.s[r][64 + tx * 4]) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    __syncthreads();

    if (warp < 4) {
      const int r = warp * 32 + lane;
      if (P.dump != nullptr) {
        if (own_row < P.n_queries)
          for (int c = 0; c < XN; ++c)
            if (item0 + c < P.n_items) P.dump[static_cast<int64_t>(own_row) * P.n_items + item0 + c] = sm.s[r][c];
      } else {
        if (row_on) st.thr = fmaxf(st.thr, ord2f(thr_load(P, own_row)));
        for (int g = 0; g < XN / 32; ++g) {
          float4 v[8];
          float m = -INFINITY;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            v[c] = *reinterpret_cast<const float4*>(&sm.s[r][g * 32 + c * 4]);
            m = fmaxf(m, fmaxf(fmaxf(v[c].x, v[c].y), fmaxf(v[c].z, v[c].w)));
          }
          if (__any_sync(kFull, m >= st.thr)) {
            if (m >= st.thr) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float e[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const int64_t item = item0 + g * 32 + c * 4 + x;
                  if (e[x] >= st.thr && item_ok(P, own_row, item)) row_append(P, st, e[x], item);
                }
              }
            }
            warp_compact_if_needed(P, st, own_row, lane);
          }
        }
      }
    }
    __syncthreads();
  }

  if (warp < 4 && P.dump == nullptr) {
    const unsigned fresh = __ballot_sync(kFull, row_on && st.cnt > 0);
    warp_compact_if_needed(P, st, own_row, lane, fresh);
    P.cand_cnt[st.slot] = st.cnt;
    if (P.stat_candidates != nullptr && st.cnt > 0) atomicAdd(P.stat_candidates, static_cast<unsigned long long>(st.cnt));
  }
}

cudaError_t launch_scan_fp32(const ScanParams& P, const float* Q, const float* E, cudaStream_t stream) {
  const size_t smem = sizeof(Smem);
  cudaError_t e = cudaFuncSetAttribute(scan_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  const int units = P.n_qblocks * P.n_splits;
  scan_fp32_kernel<<<units, 256, smem, stream>>>(P, Q, E);
  return cudaGetLastError();
}

}  // namespace bb
