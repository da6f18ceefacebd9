// Item-item collaborative filtering (BASELINE config 3; no reference counterpart,
// the schema stub is src/db/user_interaction_schema.sql:136-145): cosine between
// item columns of the sparse user x item rating matrix, top-k neighbours per item.
//
// Two paths.  Catalog-sized item sets (N*N*4 B <= 6 GiB, e.g. the reference's 25,216 sets = 2.5 GB) take the
// dense-Gram path below (global fire-and-forget reductions + one streaming select pass).  Larger item sets take
// the accumulator path, where the N x N Gram matrix is never materialized:
//   1. CSR -> CSC on device (histogram, scan, scatter) + column square norms,
//   2. items ordered by descending degree bucket (hot columns first: LPT balance),
//   3. one CTA per item i: a shared-memory accumulator over all items j receives
//      r_ui * r_uj for every user u that rated i (shared-memory atomics; integer
//      ratings make the sums exact and order independent), then the same CTA
//      normalizes and selects the top-k without the row ever reaching HBM.
// Algorithmic work (sparse): sum_u nnz_u^2 multiply-adds; bytes >= nnz * 8.
#include "bb_common.cuh"
#include "bb_kernels.h"

namespace bb {
namespace {

constexpr int kIIThreads = 256;
constexpr int kIIWarps = kIIThreads / 32;

__global__ void ii_hist_kernel(const int32_t* __restrict__ indices, const float* __restrict__ values, int64_t nnz,
                               unsigned long long* __restrict__ col_cnt, float* __restrict__ sq) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int32_t c = indices[i];
  atomicAdd(col_cnt + c, 1ull);
  const float v = values[i];
  atomicAdd(sq + c, v * v);
}

// single-block exclusive scan (n_items is catalog sized; this is microseconds)
__global__ void ii_scan_kernel(const unsigned long long* __restrict__ cnt, int n, long long* __restrict__ ptr,
                               unsigned long long* __restrict__ cursor) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int lo = t * per, hi = min(n, lo + per);
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += static_cast<long long>(cnt[i]);
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int i = 0; i < 1024; ++i) { const long long v = part[i]; part[i] = run; run += v; }
    ptr[n] = run;
  }
  __syncthreads();
  long long run = part[t];
  for (int i = lo; i < hi; ++i) {
    ptr[i] = run;
    cursor[i] = static_cast<unsigned long long>(run);
    run += static_cast<long long>(cnt[i]);
  }
}

__global__ void ii_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  const float* __restrict__ values, int64_t n_users,
                                  unsigned long long* __restrict__ cursor, int32_t* __restrict__ csc_user,
                                  float* __restrict__ csc_val) {
  const int lane = threadIdx.x & 31;
  const int64_t u = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_users) return;
  for (int64_t e = indptr[u] + lane; e < indptr[u + 1]; e += 32) {
    const unsigned long long pos = atomicAdd(cursor + indices[e], 1ull);
    csc_user[pos] = static_cast<int32_t>(u);
    csc_val[pos] = values[e];
  }
}

// order items by descending floor(log2(degree+1)) bucket
__global__ void ii_bucket_count_kernel(const unsigned long long* __restrict__ cnt, int n, int* __restrict__ bucket_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = 63 - __clzll(static_cast<long long>(cnt[i]) + 1);
  atomicAdd(bucket_cnt + (63 - b), 1);
}
__global__ void ii_bucket_scan_kernel(int* bucket_cnt, int* bucket_pos) {
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < 64; ++b) { bucket_pos[b] = run; run += bucket_cnt[b]; }
  }
}
__global__ void ii_bucket_scatter_kernel(const unsigned long long* __restrict__ cnt, int n, int* __restrict__ bucket_pos,
                                         int32_t* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = 63 - __clzll(static_cast<long long>(cnt[i]) + 1);
  order[atomicAdd(bucket_pos + (63 - b), 1)] = i;
}

struct IIMerger {          // per-warp running top-128 (same scheme as the scan merge)
  uint64_t* stage;
  int n_new;
  int lane;
  uint64_t floor_key;      // keys below the current 128th best are dropped early
  __device__ void init(uint64_t* s, int l) {
    stage = s; n_new = 0; lane = l; floor_key = 0;
    for (int i = lane; i < kCap; i += 32) stage[i] = 0ull;
    __syncwarp();
  }
  __device__ void flush() {
    uint64_t v[kEpl];
#pragma unroll
    for (int i = 0; i < kEpl; ++i) v[i] = stage[(i << 5) | lane];
    __syncwarp();
    warp_sort_desc<kEpl>(v, lane);
#pragma unroll
    for (int i = 0; i < kEpl; ++i) stage[(i << 5) | lane] = (i < kEpl / 2) ? v[i] : 0ull;
    floor_key = __shfl_sync(kFull, v[kEpl / 2 - 1], 31);
    n_new = 0;
    __syncwarp();
  }
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

__global__ void __launch_bounds__(kIIThreads)
ii_main_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
               const long long* __restrict__ col_ptr, const int32_t* __restrict__ csc_user,
               const float* __restrict__ csc_val, const float* __restrict__ sq, const int32_t* __restrict__ order,
               int n_items, int range, int k, int* __restrict__ work_counter, float* __restrict__ out_scores,
               int32_t* __restrict__ out_ids) {
  extern __shared__ __align__(16) unsigned char ii_smem[];
  float* acc = reinterpret_cast<float*>(ii_smem);                                   // [range]
  uint64_t* stages = reinterpret_cast<uint64_t*>(ii_smem + static_cast<size_t>(range) * 4);   // [kIIWarps][kCap]
  __shared__ int s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  while (true) {
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int slot = s_item;
    __syncthreads();
    if (slot >= n_items) break;
    const int i = order[slot];
    const float sq_i = sq[i];
    IIMerger mg;
    mg.init(stages + warp * kCap, lane);

    for (int j0 = 0; j0 < n_items; j0 += range) {
      const int j1 = min(n_items, j0 + range);
      for (int j = tid; j < j1 - j0; j += kIIThreads) acc[j] = 0.f;
      __syncthreads();
      // accumulate r_ui * r_uj for every user of column i (one warp per user)
      for (long long e = col_ptr[i] + warp; e < col_ptr[i + 1]; e += kIIWarps) {
        const int32_t u = csc_user[e];
        const float r_ui = csc_val[e];
        for (int64_t p = indptr[u] + lane; p < indptr[u + 1]; p += 32) {
          const int32_t j = indices[p];
          if (j >= j0 && j < j1) atomicAdd(acc + (j - j0), r_ui * values[p]);
        }
      }
      __syncthreads();
      // normalize + select: each warp scans a strided slice of the accumulator
      for (int base = warp * 32; base < j1 - j0; base += kIIThreads) {
        const int jj = base + lane;
        uint64_t key = 0ull;
        if (jj < j1 - j0) {
          const float g = acc[jj];
          const int j = j0 + jj;
          if (g > 0.f && j != i) key = make_key(g / (sqrtf(sq_i) * sqrtf(sq[j])), static_cast<uint32_t>(j));
        }
        mg.offer(key);
      }
      __syncthreads();
    }
    mg.flush();
    __syncthreads();
    // warp 0 merges the 8 per-warp tops
    if (warp == 0) {
      IIMerger fin;
      uint64_t* s0 = stages;                       // reuse warp 0's stage as the final one
      // copy other warps' results through registers, then offer
      uint64_t mine[kEpl / 2];
#pragma unroll
      for (int x = 0; x < kEpl / 2; ++x) mine[x] = s0[(x << 5) | lane];
      __syncwarp();
      fin.init(s0, lane);
#pragma unroll
      for (int x = 0; x < kEpl / 2; ++x) fin.offer(mine[x]);
      for (int w = 1; w < kIIWarps; ++w)
        for (int x = 0; x < kCap / 2; x += 32) fin.offer(stages[w * kCap + x + lane]);
      fin.flush();
      // emit: positives first, then zero-similarity items in ascending id order
      int n_pos = 0;
      for (int x = 0; x < k; x += 32) {
        const uint64_t key = (x + lane < k) ? s0[x + lane] : 0ull;
        n_pos += __popc(__ballot_sync(kFull, key != 0ull));
        if (x + lane < k && key != 0ull) {
          out_scores[static_cast<int64_t>(i) * k + x + lane] = key_score(key);
          out_ids[static_cast<int64_t>(i) * k + x + lane] = key_id(key);
        }
      }
      if (n_pos < k && lane == 0) {
        int filled = n_pos;
        for (int j = 0; j < n_items && filled < k; ++j) {
          if (j == i) continue;
          bool used = false;
          for (int x = 0; x < n_pos; ++x) used |= (key_id(s0[x]) == j);
          if (used) continue;
          out_scores[static_cast<int64_t>(i) * k + filled] = 0.f;
          out_ids[static_cast<int64_t>(i) * k + filled] = j;
          ++filled;
        }
        for (; filled < k; ++filled) {
          out_scores[static_cast<int64_t>(i) * k + filled] = -INFINITY;
          out_ids[static_cast<int64_t>(i) * k + filled] = -1;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- dense-Gram path (catalog-sized item sets)
// For N items with N*N*4 B within budget the Gram matrix G = X^T X is accumulated in HBM with fire-and-forget
// global reductions (RED.ADD.F32: ~1 per lane per cycle per SM, no serialization behind hot items), one warp per
// user: every ordered pair (a, b) of the user's rated items adds r_a * r_b to G[a][b].  A second kernel streams
// each row once (coalesced), normalizes by the diagonal and selects the top-k.  Integer ratings keep every partial
// sum exact (< 2^24), so the result does not depend on the order of the atomics.
__global__ void __launch_bounds__(256)
ii_gram_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
               int64_t n_users, int n_items, float* __restrict__ gram) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t u = warp0; u < n_users; u += n_warps) {
    const int64_t lo = indptr[u], hi = indptr[u + 1];
    const int d = static_cast<int>(hi - lo);
    // pairs (a, b): lanes sweep the flattened d*d index space
    for (int p = lane; p < d * d; p += 32) {
      const int a = p / d, b = p - a * d;
      const int32_t ia = __ldg(indices + lo + a), ib = __ldg(indices + lo + b);
      atomicAdd(gram + static_cast<int64_t>(ia) * n_items + ib, __ldg(values + lo + a) * __ldg(values + lo + b));
    }
  }
}

__global__ void __launch_bounds__(kIIThreads)
ii_select_kernel(const float* __restrict__ gram, int n_items, int k, float* __restrict__ out_scores,
                 int32_t* __restrict__ out_ids) {
  __shared__ uint64_t stages[kIIWarps * kCap];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
    const float* row = gram + static_cast<int64_t>(i) * n_items;
    const float sq_i = row[i];
    IIMerger mg;
    mg.init(stages + warp * kCap, lane);
    for (int base = warp * 32; base < n_items; base += kIIThreads) {
      const int j = base + lane;
      uint64_t key = 0ull;
      if (j < n_items) {
        const float g = __ldg(row + j);
        if (g > 0.f && j != i) key = make_key(g / (sqrtf(sq_i) * sqrtf(__ldg(gram + static_cast<int64_t>(j) * n_items + j))), static_cast<uint32_t>(j));
      }
      mg.offer(key);
    }
    mg.flush();
    __syncthreads();
    if (warp == 0) {
      uint64_t* s0 = stages;
      uint64_t mine[kEpl / 2];
#pragma unroll
      for (int x = 0; x < kEpl / 2; ++x) mine[x] = s0[(x << 5) | lane];
      __syncwarp();
      IIMerger fin;
      fin.init(s0, lane);
#pragma unroll
      for (int x = 0; x < kEpl / 2; ++x) fin.offer(mine[x]);
      for (int w = 1; w < kIIWarps; ++w)
        for (int x = 0; x < kCap / 2; x += 32) fin.offer(stages[w * kCap + x + lane]);
      fin.flush();
      int n_pos = 0;
      for (int x = 0; x < k; x += 32) {
        const uint64_t key = (x + lane < k) ? s0[x + lane] : 0ull;
        n_pos += __popc(__ballot_sync(kFull, key != 0ull));
        if (x + lane < k && key != 0ull) {
          out_scores[static_cast<int64_t>(i) * k + x + lane] = key_score(key);
          out_ids[static_cast<int64_t>(i) * k + x + lane] = key_id(key);
        }
      }
      if (n_pos < k && lane == 0) {          // fewer than k co-rated neighbours: zero-similarity items by ascending id
        int filled = n_pos;
        for (int j = 0; j < n_items && filled < k; ++j) {
          if (j == i) continue;
          bool used = false;
          for (int x = 0; x < n_pos; ++x) used |= (key_id(s0[x]) == j);
          if (used) continue;
          out_scores[static_cast<int64_t>(i) * k + filled] = 0.f;
          out_ids[static_cast<int64_t>(i) * k + filled] = j;
          ++filled;
        }
        for (; filled < k; ++filled) {
          out_scores[static_cast<int64_t>(i) * k + filled] = -INFINITY;
          out_ids[static_cast<int64_t>(i) * k + filled] = -1;
        }
      }
    }
    __syncthreads();
  }
}

constexpr size_t kGramBudgetBytes = static_cast<size_t>(6) << 30;    // dense path when N*N*4 B <= 6 GiB (N <= ~40k)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

size_t item_item_workspace_bytes(int64_t n_users, int n_items, int64_t nnz) {
  (void)n_users;
  const size_t gram = static_cast<size_t>(n_items) * n_items * sizeof(float);
  if (gram <= kGramBudgetBytes) return align_up(gram, 256) + 256;
  size_t b = 0;
  b += align_up(static_cast<size_t>(n_items + 1) * 8, 256);   // col_cnt
  b += align_up(static_cast<size_t>(n_items + 1) * 8, 256);   // col_ptr
  b += align_up(static_cast<size_t>(n_items + 1) * 8, 256);   // cursor
  b += align_up(static_cast<size_t>(n_items) * 4, 256);       // sq
  b += align_up(static_cast<size_t>(n_items) * 4, 256);       // order
  b += align_up(static_cast<size_t>(nnz) * 4, 256);           // csc_user
  b += align_up(static_cast<size_t>(nnz) * 4, 256);           // csc_val
  b += 1024;                                                  // bucket counters + work counter
  return b;
}

cudaError_t launch_item_item(int64_t n_users, int n_items, const int64_t* indptr, const int32_t* indices,
                             const float* values, int k, float* out_scores, int32_t* out_ids, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream, int* n_launches) {
  // nnz lives in indptr[n_users] on the device; the caller passes it through workspace_bytes sizing,
  // so read it back here (one 8-byte copy).
  int64_t nnz = 0;
  cudaError_t e = cudaMemcpyAsync(&nnz, indptr + n_users, sizeof(int64_t), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return e;
  if (item_item_workspace_bytes(n_users, n_items, nnz) > workspace_bytes) return cudaErrorInvalidValue;
  int dev0 = 0, sms0 = 148;
  cudaGetDevice(&dev0);
  cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
  const size_t gram_bytes = static_cast<size_t>(n_items) * n_items * sizeof(float);
  if (gram_bytes <= kGramBudgetBytes) {
    float* gram = static_cast<float*>(workspace);
    e = cudaMemsetAsync(gram, 0, gram_bytes, stream);
    if (e != cudaSuccess) return e;
    ii_gram_kernel<<<sms0 * 8, 256, 0, stream>>>(indptr, indices, values, n_users, n_items, gram);
    ii_select_kernel<<<sms0 * 4, kIIThreads, 0, stream>>>(gram, n_items, k, out_scores, out_ids);
    if (n_launches != nullptr) *n_launches = 2;
    return cudaGetLastError();
  }

  unsigned char* w = static_cast<unsigned char*>(workspace);
  auto take = [&](size_t bytes) { unsigned char* p = w; w += align_up(bytes, 256); return p; };
  unsigned long long* col_cnt = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(n_items + 1) * 8));
  long long* col_ptr = reinterpret_cast<long long*>(take(static_cast<size_t>(n_items + 1) * 8));
  unsigned long long* cursor = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(n_items + 1) * 8));
  float* sq = reinterpret_cast<float*>(take(static_cast<size_t>(n_items) * 4));
  int32_t* order = reinterpret_cast<int32_t*>(take(static_cast<size_t>(n_items) * 4));
  int32_t* csc_user = reinterpret_cast<int32_t*>(take(static_cast<size_t>(nnz) * 4));
  float* csc_val = reinterpret_cast<float*>(take(static_cast<size_t>(nnz) * 4));
  int* counters = reinterpret_cast<int*>(take(1024));   // [0..63] bucket_cnt, [64..127] bucket_pos, [128] work counter

  cudaMemsetAsync(col_cnt, 0, static_cast<size_t>(n_items + 1) * 8, stream);
  cudaMemsetAsync(sq, 0, static_cast<size_t>(n_items) * 4, stream);
  cudaMemsetAsync(counters, 0, 1024, stream);
  int launches = 0;
  if (nnz > 0) {
    ii_hist_kernel<<<static_cast<int>((nnz + 255) / 256), 256, 0, stream>>>(indices, values, nnz, col_cnt, sq);
    ++launches;
  }
  ii_scan_kernel<<<1, 1024, 0, stream>>>(col_cnt, n_items, col_ptr, cursor);
  ++launches;
  if (n_users > 0) {
    ii_scatter_kernel<<<static_cast<int>((n_users * 32 + 255) / 256), 256, 0, stream>>>(indptr, indices, values, n_users,
                                                                                      cursor, csc_user, csc_val);
    ++launches;
  }
  ii_bucket_count_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(col_cnt, n_items, counters);
  ii_bucket_scan_kernel<<<1, 32, 0, stream>>>(counters, counters + 64);
  ii_bucket_scatter_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(col_cnt, n_items, counters + 64, order);
  launches += 3;

  // accumulator range: as many items as fit beside the per-warp merge stages
  const int max_smem = 227 * 1024;
  const int stage_bytes = kIIWarps * kCap * 8;
  int range = (max_smem - stage_bytes - 1024) / 4;
  if (range > n_items) range = n_items;
  range = (range + 31) / 32 * 32;
  const size_t smem = static_cast<size_t>(range) * 4 + stage_bytes;
  e = cudaFuncSetAttribute(ii_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
  int grid = sms * ctas_per_sm;
  if (grid > n_items) grid = n_items;
  ii_main_kernel<<<grid, kIIThreads, smem, stream>>>(indptr, indices, values, col_ptr, csc_user, csc_val, sq, order,
                                                     n_items, range, k, counters + 128, out_scores, out_ids);
  ++launches;
  if (n_launches != nullptr) *n_launches = launches;
  return cudaGetLastError();
}

}  // namespace bb
