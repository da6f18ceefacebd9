// CSR (ratings) x dense fp64 product: the two sparse products of the randomized truncated SVD that trains the
// collaborative-filtering factors (reference: recommendation_system.py:368-409 -> sklearn TruncatedSVD: A @ Q and
// A.T @ Q inside the power iterations, Q.T @ A for the projection, A @ V for the user factors).
//
//   Y[r, :] = sum_{e in row r} values[e] * X[indices[e], :]          (fp32 ratings, fp64 X / accumulate / Y)
//
// A.T @ Q is the same kernel on the CSR of A.T (built once by the caller).
//
// Bound: HBM / L2 gather.  Algorithmic bytes per launch = nnz * (8 + L * 8) + n_rows * (8 + L * 8): each stored
// rating reads its (index, value) pair and gathers one L-wide row of X; rows of X are contiguous (L * 8 B, 480 B at
// L = 60), so a warp's gather is one or two fully used 128 B lines per 16 columns.
//
// One warp per CSR row: 32 (index, value) pairs are loaded coalesced, broadcast by shuffle, and each lane
// accumulates columns lane, lane+32, ...  Rows longer than kLongRow entries (popular items of A.T) are left to a
// second kernel that spreads each of them over many warps (see csr_spmm_long_rows_kernel).
#include "bb_common.cuh"
#include "bb_kernels.h"

namespace bb {
namespace {

constexpr int kLongRow = 2048;
constexpr int kMaxLpt = 4;     // L <= 128 columns

template <int LPT>
__device__ __forceinline__ void spmm_accumulate(const int32_t* __restrict__ indices, const float* __restrict__ values,
                                                const double* __restrict__ X, int L, int64_t e0, int64_t e1, int lane,
                                                double (&acc)[LPT]) {
  for (int64_t base = e0; base < e1; base += 32) {
    const int64_t e = base + lane;
    int32_t c = 0;
    float v = 0.f;
    if (e < e1) {
      c = __ldg(indices + e);
      v = __ldg(values + e);
    }
    const int cnt = static_cast<int>(min(static_cast<int64_t>(32), e1 - base));
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const int32_t cj = __shfl_sync(kFull, c, j);
      const double vj = static_cast<double>(__shfl_sync(kFull, v, j));
      const double* xr = X + static_cast<int64_t>(cj) * L;
#pragma unroll
      for (int t = 0; t < LPT; ++t) {
        const int col = lane + 32 * t;
        if (col < L) acc[t] = fma(vj, __ldg(xr + col), acc[t]);
      }
    }
  }
}

template <int LPT>
__global__ void __launch_bounds__(256) csr_spmm_rows_kernel(int64_t n_rows, const int64_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices, const float* __restrict__ values,
                                                            const double* __restrict__ X, int L, double* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const int64_t e0 = __ldg(indptr + r), e1 = __ldg(indptr + r + 1);
    double acc[LPT];
#pragma unroll
    for (int t = 0; t < LPT; ++t) acc[t] = 0.0;
    if (e1 - e0 <= kLongRow) spmm_accumulate<LPT>(indices, values, X, L, e0, e1, lane, acc);   // (long rows: zeroed here, summed by the chunk kernel)
#pragma unroll
    for (int t = 0; t < LPT; ++t) {
      const int col = lane + 32 * t;
      if (col < L) Y[r * L + col] = acc[t];
    }
  }
}

// Long rows (popular items in the CSR of A.T: one row can hold a few percent of all ratings): the stored entries are
// cut into fixed chunks of kChunk consecutive entries, one warp per chunk; the warp finds the rows its chunk overlaps
// (binary search in indptr), accumulates the part of every LONG row inside the chunk and adds it to Y with fp64
// atomics (the row kernel zeroed those rows).  Short rows are skipped here.  The summation order of a long row's
// chunks is not fixed, so its sums can differ in the last bits from run to run.
constexpr int kChunk = 1024;
template <int LPT>
__global__ void __launch_bounds__(256) csr_spmm_long_rows_kernel(int64_t n_rows, int64_t nnz, const int64_t* __restrict__ indptr,
                                                                 const int32_t* __restrict__ indices, const float* __restrict__ values,
                                                                 const double* __restrict__ X, int L, double* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t n_chunks = (nnz + kChunk - 1) / kChunk;
  for (int64_t c = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; c < n_chunks; c += warps) {
    const int64_t c0 = c * kChunk, c1 = min(nnz, c0 + kChunk);
    // last row r with indptr[r] <= c0
    int64_t lo = 0, hi = n_rows;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(indptr + mid) <= c0) lo = mid; else hi = mid;
    }
    for (int64_t r = lo; r < n_rows; ++r) {
      const int64_t e0 = __ldg(indptr + r), e1 = __ldg(indptr + r + 1);
      if (e0 >= c1) break;
      if (e1 - e0 <= kLongRow) continue;
      double acc[LPT];
#pragma unroll
      for (int t = 0; t < LPT; ++t) acc[t] = 0.0;
      spmm_accumulate<LPT>(indices, values, X, L, max(e0, c0), min(e1, c1), lane, acc);
#pragma unroll
      for (int t = 0; t < LPT; ++t) {
        const int col = lane + 32 * t;
        if (col < L) atomicAdd(Y + r * L + col, acc[t]);
      }
    }
  }
}

template <int LPT>
cudaError_t launch_spmm_lpt(int64_t n_rows, int64_t nnz, const int64_t* indptr, const int32_t* indices, const float* values, const double* X,
                            int L, double* Y, int sm_count, cudaStream_t stream) {
  const int64_t want = (n_rows * 32 + 255) / 256;
  const int grid = static_cast<int>(max(static_cast<int64_t>(1), min(want, static_cast<int64_t>(sm_count) * 32)));
  csr_spmm_rows_kernel<LPT><<<grid, 256, 0, stream>>>(n_rows, indptr, indices, values, X, L, Y);
  const int64_t n_chunks = (nnz + kChunk - 1) / kChunk;
  const int grid2 = static_cast<int>(max(static_cast<int64_t>(1), min((n_chunks * 32 + 255) / 256, static_cast<int64_t>(sm_count) * 32)));
  csr_spmm_long_rows_kernel<LPT><<<grid2, 256, 0, stream>>>(n_rows, nnz, indptr, indices, values, X, L, Y);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_csr_spmm_f64(int64_t n_rows, int64_t nnz, const int64_t* indptr, const int32_t* indices, const float* values,
                                const double* X, int L, double* Y, int sm_count, cudaStream_t stream) {
  if (L <= 32) return launch_spmm_lpt<1>(n_rows, nnz, indptr, indices, values, X, L, Y, sm_count, stream);
  if (L <= 64) return launch_spmm_lpt<2>(n_rows, nnz, indptr, indices, values, X, L, Y, sm_count, stream);
  if (L <= 96) return launch_spmm_lpt<3>(n_rows, nnz, indptr, indices, values, X, L, Y, sm_count, stream);
  return launch_spmm_lpt<kMaxLpt>(n_rows, nnz, indptr, indices, values, X, L, Y, sm_count, stream);
}

}  // namespace bb
