// Item-item collaborative filtering (BASELINE config 3; no reference counterpart,
// the schema stub is src/db/user_interaction_schema.sql:136-145): cosine between
// item columns of the sparse user x item rating matrix, top-k neighbours per item.
//
// The N x N Gram matrix is never materialized.  Pipeline (all on the device):
//   1. CSR -> CSC with a multi-block counting sort: every CTA owns a contiguous user range, histograms its ratings per
//      item in SHARED memory, a column scan turns the per-CTA histograms into per-CTA write cursors, and the CTAs
//      scatter (user, rating) with shared-memory cursors.  No global atomic ever lands on a hot item's counter (a
//      Zipf catalog sends ~1M ratings to its most popular item; the first version serialized on that address).
//      Catalogs whose histogram does not fit in shared memory (> ~56k items) use global-atomic kernels instead.
//   2. work units = (item i, chunk of <= C of i's raters), hot items first (LPT).  One CTA per unit: a dense
//      shared-memory accumulator over all items j receives r_ui * r_uj for every rater u of the chunk (1024 threads:
//      a warp fetches 32 raters' row bounds at once, then walks the rows with coalesced reads; shared-memory float
//      atomics; integer ratings keep the sums exact and order independent).
//   3. an item with ONE unit is normalized and selected straight from shared memory; a hot item's units add their
//      partial rows to a global row buffer (coalesced fire-and-forget reductions) and the unit that finishes last
//      selects from it.
// Algorithmic work (sparse): sum_u nnz_u^2 multiply-adds; bytes >= nnz * 8.  The first version (global reductions into
// a dense 2.5 GB Gram + a streaming select) took 27.5 ms at 1M users x 25,216 items and was bound by same-address
// atomics; this one is bound by shared-memory atomic throughput.
#include <cstdlib>
#include "bb_common.cuh"
#include "bb_kernels.h"

namespace bb {
namespace {

constexpr int kAccThreads = 1024;                 // threads of the counting-sort CTAs (latency-bound gathers: more warps in flight)
constexpr int kSortCtas = 148;                    // CTAs of the counting sort (one per SM)
constexpr int kSmemHistItems = 56 * 1024;         // largest item count whose uint32 histogram fits in shared memory

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ 1. CSR -> CSC (counting sort)
__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// users [u0, u1) of sort CTA g: ranges balanced by rating count
__device__ __forceinline__ void sort_user_range(const int64_t* indptr, int64_t n_users, int64_t nnz, int g, int64_t& u0, int64_t& u1) {
  const int64_t t0 = nnz * g / kSortCtas, t1 = nnz * (g + 1) / kSortCtas;
  u0 = g == 0 ? 0 : lower_bound_i64(indptr, n_users + 1, t0);
  u1 = g == kSortCtas - 1 ? n_users : lower_bound_i64(indptr, n_users + 1, t1);
  if (u0 > n_users) u0 = n_users;
  if (u1 > n_users) u1 = n_users;
}

__global__ void __launch_bounds__(kAccThreads)
ii_block_hist_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                     int64_t n_users, int64_t nnz, int n_items, uint32_t* __restrict__ block_hist, float* __restrict__ block_sq,
                     int* __restrict__ value_flags) {
  extern __shared__ __align__(16) unsigned char ii_smem[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(ii_smem);            // [n_items]
  float* sqs = reinterpret_cast<float*>(ii_smem) + n_items;         // [n_items]
  for (int i = threadIdx.x; i < n_items; i += kAccThreads) { hist[i] = 0u; sqs[i] = 0.f; }
  __syncthreads();
  int64_t u0, u1;
  sort_user_range(indptr, n_users, nnz, blockIdx.x, u0, u1);
  const int64_t e0 = __ldg(indptr + u0), e1 = __ldg(indptr + u1);
  bool odd = false;                                 // a rating that is not an integer in [0, 1024]
  float vmax = 0.f;
  for (int64_t eb = e0 + threadIdx.x; eb < e1; eb += 4 * kAccThreads) {
    int32_t c[4];
    float v[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int64_t e = eb + x * kAccThreads;
      c[x] = e < e1 ? __ldg(indices + e) : -1;
      v[x] = e < e1 ? __ldg(values + e) : 0.f;
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      if (c[x] < 0) continue;
      atomicAdd(hist + c[x], 1u);
      atomicAdd(sqs + c[x], v[x] * v[x]);
      odd |= !(v[x] >= 0.f && v[x] <= 1024.f && v[x] == floorf(v[x]));
      vmax = fmaxf(vmax, v[x]);
    }
  }
  if (__any_sync(kFull, odd) && (threadIdx.x & 31) == 0) atomicOr(value_flags, 1);
  vmax = fmaxf(vmax, __shfl_xor_sync(kFull, vmax, 16));
  vmax = fmaxf(vmax, __shfl_xor_sync(kFull, vmax, 8));
  vmax = fmaxf(vmax, __shfl_xor_sync(kFull, vmax, 4));
  vmax = fmaxf(vmax, __shfl_xor_sync(kFull, vmax, 2));
  vmax = fmaxf(vmax, __shfl_xor_sync(kFull, vmax, 1));
  if ((threadIdx.x & 31) == 0 && !odd) atomicMax(value_flags + 1, static_cast<int>(vmax));
  if (blockIdx.x == 0 && threadIdx.x == 0) value_flags[2] = 1;      // the flags are valid (this kernel ran)
  __syncthreads();
  uint32_t* out = block_hist + static_cast<int64_t>(blockIdx.x) * n_items;
  float* out_sq = block_sq + static_cast<int64_t>(blockIdx.x) * n_items;
  for (int i = threadIdx.x; i < n_items; i += kAccThreads) { out[i] = hist[i]; out_sq[i] = sqs[i]; }
}

// per item: total count and square norm; the per-CTA histograms become per-CTA exclusive prefixes (relative to the
// item's column start)
__global__ void ii_column_prefix_kernel(uint32_t* __restrict__ block_hist, const float* __restrict__ block_sq, int n_items,
                                        unsigned long long* __restrict__ col_cnt, float* __restrict__ sq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  uint32_t run = 0u;
  float s = 0.f;
  constexpr int kBatch = 8;                        // independent loads per round (the loop is latency bound)
  for (int g0 = 0; g0 < kSortCtas; g0 += kBatch) {
    uint32_t c[kBatch];
    float q[kBatch];
#pragma unroll
    for (int x = 0; x < kBatch; ++x) {
      const bool in = g0 + x < kSortCtas;
      c[x] = in ? block_hist[static_cast<int64_t>(g0 + x) * n_items + i] : 0u;
      q[x] = in ? block_sq[static_cast<int64_t>(g0 + x) * n_items + i] : 0.f;
    }
#pragma unroll
    for (int x = 0; x < kBatch; ++x) {
      if (g0 + x < kSortCtas) block_hist[static_cast<int64_t>(g0 + x) * n_items + i] = run;
      run += c[x];
      s += q[x];
    }
  }
  col_cnt[i] = run;
  sq[i] = s > 0.f ? 1.f / sqrtf(s) : 0.f;         // 1 / the item's L2 norm (what the cosine multiplies by)
}

// Exclusive scan over n elements by ONE block of 1024 threads: coalesced chunks of 1024, warp shuffles, a running
// carry.  `load(i)` yields element i, `store(i, exclusive_prefix, value)` consumes it; returns the total (all threads).
template <typename Load, typename Store>
__device__ __forceinline__ long long block_scan_1024(int n, Load load, Store store) {
  __shared__ long long scan_ws[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  long long carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    const long long v = i < n ? load(i) : 0ll;
    long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(kFull, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) scan_ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long y = scan_ws[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long z = __shfl_up_sync(kFull, y, d);
        if (lane >= d) y += z;
      }
      scan_ws[lane] = y;
    }
    __syncthreads();
    if (i < n) store(i, carry + (warp > 0 ? scan_ws[warp - 1] : 0ll) + x - v, v);
    carry += scan_ws[31];
    __syncthreads();
  }
  return carry;
}

// column pointers from the column counts (single block; n_items is catalog sized)
__global__ void __launch_bounds__(1024)
ii_scan_kernel(const unsigned long long* __restrict__ cnt, int n, long long* __restrict__ ptr, unsigned long long* __restrict__ cursor) {
  const long long total = block_scan_1024(
      n, [&](int i) { return static_cast<long long>(cnt[i]); },
      [&](int i, long long off, long long) {
        ptr[i] = off;
        if (cursor != nullptr) cursor[i] = static_cast<unsigned long long>(off);
      });
  if (threadIdx.x == 0) ptr[n] = total;
}

// A warp takes 32 consecutive users: their rows are ONE contiguous element range of the CSR arrays, read coalesced,
// lane per element; the owning user of element e is the number of non-empty rows that start at or before e (row-start
// bitmap of the 32-element window, one redux.or) -- no per-rating search in indptr.  Also writes the packed rating
// words (item | rating << 24) when ii_units_kernel found the ratings small integers.  What is left (0.45 ms at config 3)
// is the price of 19.6M random 8-byte stores: half of the 32-byte sectors leave L2 partially written and come back for a
// read-modify-write (ncu: 250 MB of fills, 236 MB of repeated write-backs); writing the output in L2-sized column slices
// (re-reading the ratings per slice) did not change the time, so a write-combining radix partition would be the next step.
__global__ void __launch_bounds__(kAccThreads)
ii_block_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                        int64_t n_users, int64_t nnz, int n_items, const uint32_t* __restrict__ block_hist,
                        const long long* __restrict__ col_ptr, int2* __restrict__ csc, const int* __restrict__ counters,
                        uint32_t* __restrict__ packed) {
  extern __shared__ __align__(16) unsigned char ii_smem[];
  uint32_t* cursor = reinterpret_cast<uint32_t*>(ii_smem);          // [n_items]: this CTA's next ABSOLUTE slot of every column (nnz < 2^32)
  const uint32_t* mine = block_hist + static_cast<int64_t>(blockIdx.x) * n_items;
  for (int i = threadIdx.x; i < n_items; i += kAccThreads) cursor[i] = static_cast<uint32_t>(col_ptr[i]) + mine[i];
  __syncthreads();
  int64_t u0, u1;
  sort_user_range(indptr, n_users, nnz, blockIdx.x, u0, u1);
  const bool pack = counters[7] != 0, pack_csc = counters[8] != 0;
  uint32_t* csc32 = reinterpret_cast<uint32_t*>(csc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t ub = u0 + warp * 32; ub < u1; ub += kAccThreads) {
    const int64_t u = ub + lane;
    int64_t p0 = 0;
    int len = 0;
    if (u < u1) {
      p0 = __ldg(indptr + u);
      len = static_cast<int>(__ldg(indptr + u + 1) - p0);
    }
    int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    const int excl = incl - len;
    const int64_t base = __shfl_sync(kFull, p0, 0);                 // (lane 0 is a valid user: ub < u1)
    const unsigned nonempty = __ballot_sync(kFull, len > 0);
    const unsigned src = __fns(nonempty, 0, lane + 1);              // lane of the (lane+1)-th non-empty row
    int rbase = 0;                                                  // non-empty rows that start before the window
    for (int e0 = 0; e0 < total; e0 += 128) {                       // four windows of rating reads in flight
      int32_t c[4];
      float v[4];
      int owner[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int w0 = e0 + 32 * x;
        const int d = excl - w0;
        const unsigned starts = __reduce_or_sync(kFull, (len > 0 && d >= 0 && d < 32) ? (1u << d) : 0u);
        const int rank = rbase + __popc(starts & (0xffffffffu >> (31 - lane)));
        rbase += __popc(starts);
        owner[x] = static_cast<int>(__shfl_sync(kFull, src, (rank - 1) & 31));
        const int e = w0 + lane;
        c[x] = e < total ? __ldcs(indices + base + e) : -1;
        v[x] = e < total ? __ldcs(values + base + e) : 0.f;
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        if (c[x] < 0) continue;
        const uint32_t pos = atomicAdd(cursor + c[x], 1u);
        if (pack_csc) csc32[pos] = static_cast<uint32_t>(ub + owner[x]) | (static_cast<uint32_t>(v[x]) << 24);
        else csc[pos] = make_int2(static_cast<int32_t>(ub + owner[x]), __float_as_int(v[x]));
        if (pack) packed[base + e0 + 32 * x + lane] = static_cast<uint32_t>(c[x]) | (static_cast<uint32_t>(v[x]) << 24);
      }
    }
  }
}

__global__ void ii_sqrt_kernel(float* __restrict__ sq, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sq[i] = sq[i] > 0.f ? 1.f / sqrtf(sq[i]) : 0.f;
}

// ---- fallback for catalogs whose histogram does not fit in shared memory: global atomics
__global__ void ii_hist_kernel(const int32_t* __restrict__ indices, const float* __restrict__ values, int64_t nnz,
                               unsigned long long* __restrict__ col_cnt, float* __restrict__ sq) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int32_t c = indices[i];
  atomicAdd(col_cnt + c, 1ull);
  const float v = values[i];
  atomicAdd(sq + c, v * v);
}
__global__ void ii_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  const float* __restrict__ values, int64_t n_users,
                                  unsigned long long* __restrict__ cursor, int2* __restrict__ csc) {
  const int lane = threadIdx.x & 31;
  const int64_t u = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_users) return;
  for (int64_t e = indptr[u] + lane; e < indptr[u + 1]; e += 32) {
    const unsigned long long pos = atomicAdd(cursor + indices[e], 1ull);
    csc[pos] = make_int2(static_cast<int32_t>(u), __float_as_int(values[e]));
  }
}

// ------------------------------------------------------------------ 2. work units
// items ordered by descending floor(log2(degree+1)) bucket (hot columns first: LPT balance)
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
// unit_off[p] = first unit of the p-th item of `order`; hot_of[item] = row of the item in the partial-row buffer
// (-1 for single-unit items); unit_tab[u] = (item, chunk index, units of the item, hot row).  counters[0] = number of
// units, counters[1] = number of hot items, counters[3] = integer accumulators, counters[7] = packed ratings.  Single block.
__global__ void __launch_bounds__(1024)
ii_units_kernel(const unsigned long long* __restrict__ col_cnt, const int32_t* __restrict__ order, int n, int chunk,
                int* __restrict__ unit_off, int* __restrict__ hot_of, int4* __restrict__ unit_tab, int* __restrict__ counters,
                int allow_packed, int allow_packed_users) {
  if (threadIdx.x == 0) {
    // integer accumulators are exact while vmax^2 * (largest degree) stays below 2^31
    // (order[0] lies in the top floor(log2(degree + 1)) bucket: no degree exceeds twice its own + 1)
    const long long dmax = n > 0 ? 2 * (static_cast<long long>(col_cnt[order[0]]) + 1) : 1;
    const long long vm = counters[5];
    counters[3] = (counters[4] == 0 && counters[6] == 1 && vm * vm * dmax < (1ll << 31)) ? 1 : 0;
    // packed ratings: (item | rating << 24) in one word per rating (written by ii_block_scatter_kernel)
    counters[7] = (allow_packed && counters[3] != 0 && vm < 256 && n <= (1 << 24)) ? 1 : 0;
    // ... and (user | rating << 24) in one word per CSC entry while the users fit 24 bits: half the scatter's random stores
    counters[8] = (counters[7] != 0 && allow_packed_users) ? 1 : 0;
  }
  // one scan carries both counts: units in the low word, hot items (more than one unit) in the high word
  const long long total = block_scan_1024(
      n,
      [&](int p) {
        const long long d = static_cast<long long>(col_cnt[order[p]]);
        const long long units = d <= chunk ? 1 : (d + chunk - 1) / chunk;
        return units | (static_cast<long long>(units > 1) << 32);
      },
      [&](int p, long long off, long long v) {
        const int item = order[p];
        const int units = static_cast<int>(v & 0xffffffffll);
        const int ru = static_cast<int>(off & 0xffffffffll), rh = static_cast<int>(off >> 32);
        const int hot = units > 1 ? rh : -1;
        unit_off[p] = ru;
        hot_of[item] = hot;
        for (int c = 0; c < units; ++c) unit_tab[ru + c] = make_int4(item, c, units, hot);
      });
  if (threadIdx.x == 0) {
    counters[0] = static_cast<int>(total & 0xffffffffll);
    counters[1] = static_cast<int>(total >> 32);
    unit_off[n] = counters[0];
  }
}

// Fused planning for catalogs that fit shared memory (the counting-sort path): column pointers, degree-bucket order
// (hot columns first) and the unit table in ONE single-block launch over shared-memory copies of the counts.  The five
// latency-sized launches it replaces (a scan, three bucket kernels, the unit scan: chains of dependent global loads in
// single blocks) took 126 us of a 2.4 ms pass at config 3.
__global__ void __launch_bounds__(1024)
ii_plan_kernel(const unsigned long long* __restrict__ col_cnt, int n, int chunk, long long* __restrict__ col_ptr,
               int4* __restrict__ unit_tab, int* __restrict__ counters, int allow_packed, int allow_packed_users) {
  extern __shared__ __align__(16) unsigned char ii_smem[];
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(ii_smem);               // [n] column degrees (nnz < 2^32 on this path)
  uint16_t* s_order = reinterpret_cast<uint16_t*>(s_cnt + n);           // [n] items by descending degree bucket (n <= 28,672)
  __shared__ int s_bucket[64], s_pos[64];
  const int t = threadIdx.x;
  for (int i = t; i < n; i += 1024) s_cnt[i] = static_cast<uint32_t>(col_cnt[i]);
  if (t < 64) s_bucket[t] = 0;
  __syncthreads();
  // column pointers
  const long long total = block_scan_1024(
      n, [&](int i) { return static_cast<long long>(s_cnt[i]); }, [&](int i, long long off, long long) { col_ptr[i] = off; });
  if (t == 0) col_ptr[n] = total;
  // items ordered by descending floor(log2(degree + 1)) (LPT balance; the order inside a bucket is arbitrary)
  for (int i = t; i < n; i += 1024) atomicAdd(&s_bucket[63 - (31 - __clz(s_cnt[i] + 1u))], 1);
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int b = 0; b < 64; ++b) { s_pos[b] = run; run += s_bucket[b]; }
  }
  __syncthreads();
  for (int i = t; i < n; i += 1024) s_order[atomicAdd(&s_pos[63 - (31 - __clz(s_cnt[i] + 1u))], 1)] = static_cast<uint16_t>(i);
  __syncthreads();
  if (t == 0) {
    // integer accumulators are exact while vmax^2 * (largest degree) stays below 2^31
    // (s_order[0] lies in the top bucket: no degree exceeds twice its own + 1)
    const long long dmax = n > 0 ? 2 * (static_cast<long long>(s_cnt[s_order[0]]) + 1) : 1;
    const long long vm = counters[5];
    counters[3] = (counters[4] == 0 && counters[6] == 1 && vm * vm * dmax < (1ll << 31)) ? 1 : 0;
    counters[7] = (allow_packed && counters[3] != 0 && vm < 256 && n <= (1 << 24)) ? 1 : 0;       // packed rating words
    counters[8] = (counters[7] != 0 && allow_packed_users) ? 1 : 0;                               // packed CSC entries
  }
  // unit table: one scan carries both counts (units in the low word, hot items in the high word)
  const long long tot = block_scan_1024(
      n,
      [&](int p) {
        const long long d = s_cnt[s_order[p]];
        const long long units = d <= chunk ? 1 : (d + chunk - 1) / chunk;
        return units | (static_cast<long long>(units > 1) << 32);
      },
      [&](int p, long long off, long long v) {
        const int item = s_order[p];
        const int units = static_cast<int>(v & 0xffffffffll);
        const int ru = static_cast<int>(off & 0xffffffffll), rh = static_cast<int>(off >> 32);
        const int hot = units > 1 ? rh : -1;
        for (int c = 0; c < units; ++c) unit_tab[ru + c] = make_int4(item, c, units, hot);
      });
  if (t == 0) {
    counters[0] = static_cast<int>(tot & 0xffffffffll);
    counters[1] = static_cast<int>(tot >> 32);
  }
}

__global__ void ii_zero_hot_kernel(float* __restrict__ hot_rows, int* __restrict__ hot_done, const int* __restrict__ counters,
                                   int n_items, int max_hot) {
  const int n_hot = min(counters[1], max_hot);
  const int64_t total = static_cast<int64_t>(n_hot) * n_items;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    hot_rows[i] = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hot; i += gridDim.x * blockDim.x) hot_done[i] = 0;
}

// ------------------------------------------------------------------ 3. accumulate + select
struct IIMerger {          // per-warp running top-128 (same scheme as the scan merge)
  uint64_t* stage;
  int n_new, n_sorted;     // new keys at stage[128 + ...], sorted keys at stage[0 ...]
  int lane;
  uint64_t floor_key;      // keys below the current 128th best are dropped early
  __device__ void init(uint64_t* s, int l) {
    stage = s; n_new = 0; n_sorted = 0; lane = l; floor_key = 0;
    for (int i = lane; i < kCap / 2; i += 32) stage[i] = 0ull;      // (the tree merge reads whole prefixes)
    __syncwarp();
  }
  // Sorts [sorted prefix | new keys] and keeps the best 128.  The sort network is sized to the number of keys that are
  // actually there (a cold item leaves a few dozen keys per warp: a 64-key network is 8x cheaper than the 256-key one).
  template <int E>
  __device__ void flush_n(int n_keys) {
    // gather the n_keys live keys (prefix [0, n_sorted) and new [128, 128 + n_new)) into E registers per lane
    uint64_t v[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int e = (i << 5) | lane;
      v[i] = e < n_sorted ? stage[e] : (e < n_keys ? stage[kCap / 2 + (e - n_sorted)] : 0ull);
    }
    __syncwarp();
    warp_sort_desc<E>(v, lane);
#pragma unroll
    for (int i = 0; i < E; ++i)
      if (i < kEpl / 2) stage[(i << 5) | lane] = v[i];
    if (E < kEpl / 2) {
#pragma unroll
      for (int i = E; i < kEpl / 2; ++i) stage[(i << 5) | lane] = 0ull;
    }
    n_sorted = n_keys < kCap / 2 ? n_keys : kCap / 2;
    uint64_t kth = 0ull;
    if (E >= kEpl / 2) kth = __shfl_sync(kFull, v[kEpl / 2 - 1], 31);
    floor_key = kth;                                 // 0 while fewer than 128 keys are held
    n_new = 0;
    __syncwarp();
  }
  __device__ void flush() {
    const int n_keys = n_sorted + n_new;
    if (n_new == 0) return;
    if (n_keys <= 32) flush_n<1>(n_keys);
    else if (n_keys <= 64) flush_n<2>(n_keys);
    else if (n_keys <= 128) flush_n<4>(n_keys);
    else flush_n<kEpl>(n_keys);
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

// ---- selection of item i's top-k from a Gram-row range held in shared memory.
// Round 1/2a kept a top-128 per warp (32 mergers, 256-key sort networks, a 5-level tree merge): 5.6G warp instructions
// for 405M multiply-adds at config 3, i.e. the pass was bound by the sort networks.  Now:
//   pass A  every thread turns its entries into cosines in place (four per shared-memory read) and keeps its best;
//   bound   T = k-th largest of the thread bests (each is the score of a distinct item, so at least k items reach T):
//           per-warp 32-key sorts, then warp 0 pops from the heads of the sorted lists (one redux.max per pop);
//   pass C  entries >= T go to a small CTA buffer that warp 0 drains into ONE merger; everything that was handled is
//           zeroed (the accumulator is clean for the next unit), an entry that found the buffer full stays and the pass
//           repeats (rare: T admits about k entries).
// (Deriving T from a sixth of the row and computing the cosines in pass C instead was slower: 3.08 vs 2.81 ms.)
constexpr int kSelBuf = 1024;                      // keys of the survivor buffer (aliases the per-warp sorted bests)

struct IISelShared {       // static shared state of the selection
  float thr_c;             // T: at least k items of the row score >= T
  int cnt;                 // survivors appended in this round
  int more;                // some survivors did not fit: repeat pass C
};

template <int kThreads, bool kInt>
__device__ __forceinline__ void ii_select_smem(float* acc, int n, int j0, int i, float inv_i, const float* __restrict__ inv_norm, int k,
                                               uint64_t* buf, IISelShared* sh, IIMerger& mg, float& thr_c, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n4 = (n + 3) >> 2;                     // (entries in [n, range) are never written: they stay zero)
  const uint4* acc4 = reinterpret_cast<const uint4*>(acc);
  const float4* inv4p = reinterpret_cast<const float4*>(inv_norm + j0);     // (j0 is a multiple of 32; padded allocation)
  // pass A.  Integer sums are never negative, so the cosine needs no sign test; the item's own entry is cleared by
  // the one thread that meets its group.
  float best_c = 0.f;
  const int self_q = (i >= j0 && i < j0 + n) ? ((i - j0) >> 2) : -1, self_u = (i - j0) & 3;
  for (int q = tid; q < n4; q += kThreads) {
    const uint4 w = acc4[q];
    if ((w.x | w.y | w.z | w.w) != 0u) {
      const float4 inv4 = __ldg(inv4p + q);
      const unsigned wu[4] = {w.x, w.y, w.z, w.w};
      const float iv[4] = {inv4.x, inv4.y, inv4.z, inv4.w};
      float c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kInt) {
          c[u] = static_cast<float>(wu[u]) * (inv_i * iv[u]);
        } else {
          const float g = __uint_as_float(wu[u]);
          c[u] = g > 0.f ? g * (inv_i * iv[u]) : 0.f;
        }
      }
      if (q == self_q) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (u == self_u) c[u] = 0.f;
      }
      best_c = fmaxf(fmaxf(fmaxf(best_c, c[0]), fmaxf(c[1], c[2])), c[3]);
      reinterpret_cast<float4*>(acc)[q] = make_float4(c[0], c[1], c[2], c[3]);
    }
  }
  // bound: T = k-th largest thread best (scores of k distinct items).  Positive floats order like their bit patterns.
  unsigned v[1] = {__float_as_uint(best_c)};
  warp_sort_desc<1, unsigned>(v, lane);
  unsigned* tops = reinterpret_cast<unsigned*>(buf);                 // [kWarps][32], dead before pass C fills buf
  tops[warp * 32 + lane] = v[0];
  __syncthreads();
  if (warp == 0) {
    int ptr = 0, popped = 0;
    unsigned head = lane < kWarps ? tops[lane * 32] : 0u;
    unsigned m = 0u;
    while (popped < k) {
      m = __reduce_max_sync(kFull, head);
      if (m == 0u) break;
      const bool hit = head == m;                  // equal scores pop together: at least `popped` items reach m
      popped += __popc(__ballot_sync(kFull, hit));
      if (hit) { ++ptr; head = ptr < 32 ? tops[lane * 32 + ptr] : 0u; }
    }
    if (lane == 0) {
      const float t_new = popped >= k ? __uint_as_float(m) : 0.f;
      sh->thr_c = fmaxf(t_new, thr_c);
      sh->cnt = 0;
      sh->more = 0;
    }
  }
  __syncthreads();
  thr_c = sh->thr_c;
  // pass C (one compare per entry: survivors are the positive entries >= T)
  const float lim = fmaxf(thr_c, 1e-37f);
  while (true) {
    for (int q = tid; q < n4; q += kThreads) {
      const float4 w = reinterpret_cast<const float4*>(acc)[q];
      if (w.x != 0.f || w.y != 0.f || w.z != 0.f || w.w != 0.f) {
        float c[4] = {w.x, w.y, w.z, w.w};
        if (c[0] >= lim || c[1] >= lim || c[2] >= lim || c[3] >= lim) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            bool keep = false;
            if (c[u] >= lim) {
              const int pos = atomicAdd(&sh->cnt, 1);
              if (pos < kSelBuf) buf[pos] = make_key(c[u], static_cast<uint32_t>(j0 + q * 4 + u));
              else keep = true;                      // buffer full: stays for the next round
            }
            if (!keep) c[u] = 0.f;
          }
          reinterpret_cast<float4*>(acc)[q] = make_float4(c[0], c[1], c[2], c[3]);
        } else {
          reinterpret_cast<float4*>(acc)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      const int appended = sh->cnt;
      const int cnt = min(appended, kSelBuf);
      for (int x = 0; x < cnt; x += 32) mg.offer(x + lane < cnt ? buf[x + lane] : 0ull);
      if (lane == 0) { sh->cnt = 0; sh->more = appended > kSelBuf; }
    }
    __syncthreads();
    if (sh->more == 0) break;                      // (written again only after the next barrier)
  }
}

// warp 0 writes item i's lists from its merger's sorted top (s0[0..127]): positives first, then zero-similarity items
// by ascending id
__device__ __forceinline__ void ii_emit(const uint64_t* s0, int i, int n_items, int k, int lane, float* __restrict__ out_scores,
                                        int32_t* __restrict__ out_ids) {
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

// One CTA per work unit (item i, chunk of its raters), units handed out by a global counter, hot items first.
// Accumulate: raters are dealt to the warps cyclically (rater e_begin + x goes to warp x % kWarps), so an item with a
// few dozen raters still keeps every warp busy.  Per step a warp holds 32 raters -- lane l has (user row start, length,
// rating) of one, the NEXT step's 32 are already being fetched -- and walks their rows as ONE flattened element range,
// lane per element: the owning rater of element e is the number of rows that end at or before e, i.e. a running base
// plus the population count of the window's row-end bitmap (one redux.or per 32 elements) below the lane.  All 32 lanes
// work on ~20-entry rows and consecutive windows are independent, so several row reads stay in flight per warp.
struct IIRaters {
  int64_t p0;      // start of the rater's row in the rating arrays
  int len;         // row length
  unsigned r;      // the rater's rating of item i (integer mode: the integer; else float bits)
};

template <int kThreads, int kMinBlocks, bool kOneRange>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
ii_accumulate_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                     const uint32_t* __restrict__ packed, const long long* __restrict__ col_ptr, const int2* __restrict__ csc,
                     const float* __restrict__ inv_norm, const int4* __restrict__ unit_tab, const int* __restrict__ counters,
                     int n_items, int range, int chunk, int k, int max_hot, float* __restrict__ hot_rows, int* __restrict__ hot_done,
                     int* __restrict__ work_counter, float* __restrict__ out_scores, int32_t* __restrict__ out_ids) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) unsigned char ii_smem[];
  float* acc = reinterpret_cast<float*>(ii_smem);                                   // [range], zero between units
  uint64_t* buf = reinterpret_cast<uint64_t*>(ii_smem + static_cast<size_t>(range) * 4);   // [kSelBuf]
  uint64_t* mstage = buf + kSelBuf;                                                 // [kCap]: warp 0's merger
  __shared__ int s_unit[2], s_last;
  __shared__ IISelShared s_sel;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_units = counters[0];
  const bool int_mode = counters[3] != 0;          // every rating is a small non-negative integer (set by ii_units_kernel)
  const bool use_packed = counters[7] != 0, packed_csc = counters[8] != 0;
  const uint32_t* csc32 = reinterpret_cast<const uint32_t*>(csc);
  for (int j = tid; j < range; j += kThreads) acc[j] = 0.f;
  if (tid == 0) s_unit[0] = atomicAdd(work_counter, 1);
  __syncthreads();

  for (int it = 0;; ++it) {
    const int unit = s_unit[it & 1];
    if (unit >= n_units) break;
    if (tid == 0) s_unit[(it + 1) & 1] = atomicAdd(work_counter, 1);     // the next unit's ticket travels while this one runs
    const int4 ut = __ldg(unit_tab + unit);
    const int i = ut.x, c_idx = ut.y, n_item_units = ut.z, hot = ut.w;
    const long long e_begin = __ldg(col_ptr + i) + static_cast<long long>(c_idx) * chunk;
    const long long e_end = min(__ldg(col_ptr + i + 1), e_begin + chunk);
    const bool single = n_item_units == 1;
    const float inv_i = __ldg(inv_norm + i);
    IIMerger mg;
    float thr_c = 0.f;
    if (single && warp == 0) mg.init(mstage, lane);

    for (int j0 = 0; j0 < n_items; j0 += range) {
      const int j1 = min(n_items, j0 + range);
      const long long n_raters = e_end - e_begin;
      auto fetch = [&](long long x0) {
        IIRaters f = {0, 0, 0u};
        const long long mine = (x0 + lane) * kWarps + warp;
        if (mine < n_raters) {
          int user;
          if (packed_csc) {                         // read once: evict first, the packed rows should stay in L2
            const uint32_t w = __ldcs(csc32 + e_begin + mine);
            user = static_cast<int>(w & 0xffffffu);
            f.r = w >> 24;
          } else {
            const int2 ur = __ldcs(csc + e_begin + mine);
            user = ur.x;
            f.r = int_mode ? static_cast<unsigned>(__int_as_float(ur.y)) : static_cast<unsigned>(ur.y);
          }
          f.p0 = __ldg(indptr + user);
          f.len = static_cast<int>(__ldg(indptr + user + 1) - f.p0);
        }
        return f;
      };
      IIRaters cur = fetch(0);
      for (long long x0 = 0; x0 * kWarps + warp < n_raters; x0 += 32) {
        const IIRaters nxt = fetch(x0 + 32);
        int incl = cur.len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(kFull, incl, d);
          if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        const int64_t abase = cur.p0 - (incl - cur.len);       // element e of the flattened range lives at abase + e
        int rbase = 0;                                         // rows that end before the window
        // several windows per round: the round's row reads are issued before the first shared-memory atomic needs one
        // (packed words: eight windows, 1 KB of row reads in flight per warp)
        if (use_packed) {
          const uint32_t abase32 = static_cast<uint32_t>(abase);     // (packed words exist only while nnz < 2^31)
          for (int e0 = 0; e0 < total; e0 += 256) {
            uint32_t ww[8];
            unsigned rr[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const int w0 = e0 + 32 * x;
              const int d = incl - w0;
              const unsigned ends = __reduce_or_sync(kFull, (d >= 0 && d < 32) ? (1u << d) : 0u);
              const int r = rbase + __popc(ends & (0xffffffffu >> (31 - lane)));
              rbase += __popc(ends);
              const int e = w0 + lane;
              const uint32_t q = __shfl_sync(kFull, abase32, r & 31) + static_cast<uint32_t>(e);
              rr[x] = __shfl_sync(kFull, cur.r, r & 31);
              ww[x] = e < total ? __ldg(packed + q) : 0u;                 // (rating 0: adds nothing, skipped below)
            }
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const unsigned v = ww[x] >> 24;
              if (kOneRange) {
                if (v != 0u) atomicAdd(reinterpret_cast<unsigned*>(acc) + (ww[x] & 0xffffffu), rr[x] * v);
              } else {
                const int j = static_cast<int>(ww[x] & 0xffffffu) - j0;
                if (static_cast<unsigned>(j) < static_cast<unsigned>(j1 - j0) && v != 0u)
                  atomicAdd(reinterpret_cast<unsigned*>(acc) + j, rr[x] * v);
              }
            }
          }
        } else {
          for (int e0 = 0; e0 < total; e0 += 128) {
            int jj[4];
            unsigned vv[4], rr[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const int w0 = e0 + 32 * x;
              const int d = incl - w0;
              const unsigned ends = __reduce_or_sync(kFull, (d >= 0 && d < 32) ? (1u << d) : 0u);
              const int r = rbase + __popc(ends & (0xffffffffu >> (31 - lane)));
              rbase += __popc(ends);
              const int e = w0 + lane;
              const int64_t q = __shfl_sync(kFull, abase, r & 31) + e;
              rr[x] = __shfl_sync(kFull, cur.r, r & 31);
              jj[x] = -1;
              vv[x] = 0u;
              if (e < total) {
                jj[x] = __ldg(indices + q);
                vv[x] = __float_as_uint(__ldg(values + q));
              }
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              if (jj[x] >= j0 && jj[x] < j1) {
                // integral ratings too large to pack: native integer shared-memory atomics; a float atomicAdd on
                // shared memory is a compare-and-swap loop that spins on popular items
                if (int_mode) atomicAdd(reinterpret_cast<unsigned*>(acc) + (jj[x] - j0), rr[x] * static_cast<unsigned>(__uint_as_float(vv[x])));
                else atomicAdd(acc + (jj[x] - j0), __uint_as_float(rr[x]) * __uint_as_float(vv[x]));
              }
            }
          }
        }
        cur = nxt;
      }
      __syncthreads();
      if (single) {
        if (int_mode) ii_select_smem<kThreads, true>(acc, j1 - j0, j0, i, inv_i, inv_norm, k, buf, &s_sel, mg, thr_c, tid);
        else ii_select_smem<kThreads, false>(acc, j1 - j0, j0, i, inv_i, inv_norm, k, buf, &s_sel, mg, thr_c, tid);
      } else {
        float* row = hot_rows + static_cast<int64_t>(hot) * n_items + j0;
        const bool add = hot >= 0 && hot < max_hot;
        for (int j = tid; j < j1 - j0; j += kThreads) {
          const float v = int_mode ? static_cast<float>(reinterpret_cast<const unsigned*>(acc)[j]) : acc[j];
          if (v != 0.f) {
            if (add) atomicAdd(row + j, v);          // coalesced fire-and-forget reductions (result unused -> RED)
            acc[j] = 0.f;
          }
        }
        __syncthreads();
      }
    }

    if (single) {
      if (warp == 0) {
        mg.flush();
        ii_emit(mstage, i, n_items, k, lane, out_scores, out_ids);
      }
    } else {
      // the unit of a hot item that finishes last selects from the completed global row
      __threadfence();
      __syncthreads();
      if (tid == 0) s_last = (atomicAdd(hot_done + hot, 1) == n_item_units - 1);
      __syncthreads();
      if (s_last) {
        __threadfence();
        if (warp == 0) mg.init(mstage, lane);
        for (int j0 = 0; j0 < n_items; j0 += range) {
          const int j1 = min(n_items, j0 + range);
          const float* row = hot_rows + static_cast<int64_t>(hot) * n_items + j0;
          for (int j = tid; j < j1 - j0; j += kThreads) acc[j] = __ldcg(row + j);
          __syncthreads();
          ii_select_smem<kThreads, false>(acc, j1 - j0, j0, i, inv_i, inv_norm, k, buf, &s_sel, mg, thr_c, tid);
        }
        if (warp == 0) {
          mg.flush();
          ii_emit(mstage, i, n_items, k, lane, out_scores, out_ids);
        }
      }
    }
    __syncthreads();                               // the next ticket is visible; warp 0 is done with its merger stage
  }
}

// raters per work unit: 16384 (BB_II_CHUNK), raised so that the partial-row buffer (<= nnz / chunk hot items x n_items floats) stays
// within 1 GiB
int ii_chunk(int n_items, int64_t nnz) {
  static const long long base = [] { const char* v = getenv("BB_II_CHUNK"); const long long c = v ? atoll(v) : 0; return c >= 256 ? c : 16384; }();
  long long chunk = base;
  const long long budget_rows = (1ll << 30) / (static_cast<long long>(n_items) * 4 + 1);
  if (budget_rows > 0 && nnz / chunk > budget_rows) chunk = (nnz + budget_rows - 1) / budget_rows;
  if (chunk < base) chunk = base;
  if (chunk > 0x3fffffff) chunk = 0x3fffffff;
  return static_cast<int>(chunk);
}
size_t ii_max_units(int n_items, int64_t nnz) {
  return static_cast<size_t>(n_items) + static_cast<size_t>(nnz / ii_chunk(n_items, nnz)) + 1;
}
int ii_max_hot(int n_items, int64_t nnz) {
  const long long m = nnz / ii_chunk(n_items, nnz) + 1;
  return static_cast<int>(m < n_items ? m : n_items);
}

}  // namespace

size_t item_item_workspace_bytes(int64_t n_users, int n_items, int64_t nnz) {
  (void)n_users;
  size_t b = 0;
  b += align_up(static_cast<size_t>(n_items + 1) * 8, 256);   // col_cnt
  b += align_up(static_cast<size_t>(n_items + 1) * 8, 256);   // col_ptr
  b += align_up(static_cast<size_t>(n_items + 1) * 8, 256);   // cursor (fallback path)
  b += align_up(static_cast<size_t>(n_items) * 4, 256);       // sq
  b += align_up(static_cast<size_t>(n_items) * 4, 256);       // order
  b += align_up(static_cast<size_t>(n_items + 1) * 4, 256);   // unit_off
  b += align_up(static_cast<size_t>(n_items) * 4, 256);       // hot_of
  b += align_up(static_cast<size_t>(nnz) * 8, 256);           // csc (user, rating)
  b += align_up(static_cast<size_t>(nnz) * 4, 256);           // packed (item | rating << 24)
  b += align_up(ii_max_units(n_items, nnz) * 16, 256);        // unit table
  b += 1024;                                                  // bucket counters, unit / hot counts, work counter
  if (n_items <= kSmemHistItems / 2 && nnz < 0xffffffffll) b += 2 * align_up(static_cast<size_t>(kSortCtas) * n_items * 4, 256);   // per-CTA histograms + square norms
  const int max_hot = ii_max_hot(n_items, nnz);
  b += align_up(static_cast<size_t>(max_hot) * n_items * 4, 256);   // partial rows of hot items
  b += align_up(static_cast<size_t>(max_hot) * 4, 256);             // hot_done
  return b;
}

cudaError_t launch_item_item(int64_t n_users, int n_items, const int64_t* indptr, const int32_t* indices,
                             const float* values, int k, float* out_scores, int32_t* out_ids, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream, int* n_launches) {
  // nnz lives in indptr[n_users] on the device (one 8-byte copy)
  int64_t nnz = 0;
  cudaError_t e = cudaMemcpyAsync(&nnz, indptr + n_users, sizeof(int64_t), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return e;
  if (item_item_workspace_bytes(n_users, n_items, nnz) > workspace_bytes) return cudaErrorInvalidValue;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  unsigned char* w = static_cast<unsigned char*>(workspace);
  auto take = [&](size_t bytes) { unsigned char* p = w; w += align_up(bytes, 256); return p; };
  unsigned long long* col_cnt = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(n_items + 1) * 8));
  long long* col_ptr = reinterpret_cast<long long*>(take(static_cast<size_t>(n_items + 1) * 8));
  unsigned long long* cursor = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(n_items + 1) * 8));
  float* sq = reinterpret_cast<float*>(take(static_cast<size_t>(n_items) * 4));
  int32_t* order = reinterpret_cast<int32_t*>(take(static_cast<size_t>(n_items) * 4));
  int* unit_off = reinterpret_cast<int*>(take(static_cast<size_t>(n_items + 1) * 4));
  int* hot_of = reinterpret_cast<int*>(take(static_cast<size_t>(n_items) * 4));
  int2* csc = reinterpret_cast<int2*>(take(static_cast<size_t>(nnz) * 8));
  uint32_t* packed = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(nnz) * 4));
  int4* unit_tab = reinterpret_cast<int4*>(take(ii_max_units(n_items, nnz) * 16));
  int* counters = reinterpret_cast<int*>(take(1024));   // [0..63] bucket_cnt, [64..127] bucket_pos, [128] units, [129] hot items, [130] work counter,
                                                        // [131] integer-accumulator mode, [132] non-integral rating seen, [133] largest rating, [134] flags valid
  const bool smem_sort = n_items <= kSmemHistItems / 2 && nnz < 0xffffffffll;   // (count + square-norm histograms side by side; 32-bit cursors)
  uint32_t* block_hist = smem_sort ? reinterpret_cast<uint32_t*>(take(static_cast<size_t>(kSortCtas) * n_items * 4)) : nullptr;
  float* block_sq = smem_sort ? reinterpret_cast<float*>(take(static_cast<size_t>(kSortCtas) * n_items * 4)) : nullptr;
  const int chunk = ii_chunk(n_items, nnz);
  const int max_hot = ii_max_hot(n_items, nnz);
  float* hot_rows = reinterpret_cast<float*>(take(static_cast<size_t>(max_hot) * n_items * 4));
  int* hot_done = reinterpret_cast<int*>(take(static_cast<size_t>(max_hot) * 4));

  cudaMemsetAsync(col_cnt, 0, static_cast<size_t>(n_items + 1) * 8, stream);
  cudaMemsetAsync(sq, 0, static_cast<size_t>(n_items) * 4, stream);
  cudaMemsetAsync(counters, 0, 1024, stream);
  int launches = 0;
  // ---- 1. column counts, norms, column pointers; 2. work units, hot items first; then the CSC scatter (which needs the
  // "ratings are small integers" verdict of the units kernel to write the packed rating words)
  auto plan_units = [&]() {
    ii_bucket_count_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(col_cnt, n_items, counters);
    ii_bucket_scan_kernel<<<1, 32, 0, stream>>>(counters, counters + 64);
    ii_bucket_scatter_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(col_cnt, n_items, counters + 64, order);
    ii_units_kernel<<<1, 1024, 0, stream>>>(col_cnt, order, n_items, chunk, unit_off, hot_of, unit_tab, counters + 128,
                                            nnz < (1ll << 31) ? 1 : 0, n_users <= (1ll << 24) ? 1 : 0);
    ii_zero_hot_kernel<<<sms * 4, 256, 0, stream>>>(hot_rows, hot_done, counters + 128, n_items, max_hot);
    launches += 5;
  };
  if (smem_sort) {
    const size_t hist_smem = static_cast<size_t>(n_items) * 8;
    if (hist_smem > 48 * 1024) {
      e = cudaFuncSetAttribute(ii_block_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hist_smem));
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(ii_block_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hist_smem));
      if (e != cudaSuccess) return e;
    }
    ii_block_hist_kernel<<<kSortCtas, kAccThreads, hist_smem, stream>>>(indptr, indices, values, n_users, nnz, n_items, block_hist, block_sq,
                                                                       counters + 128 + 4);
    ii_column_prefix_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(block_hist, block_sq, n_items, col_cnt, sq);
    {
      const size_t plan_smem = static_cast<size_t>(n_items) * 6 + 16;
      e = cudaFuncSetAttribute(ii_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan_smem));
      if (e != cudaSuccess) return e;
      ii_plan_kernel<<<1, 1024, plan_smem, stream>>>(col_cnt, n_items, chunk, col_ptr, unit_tab, counters + 128,
                                                     nnz < (1ll << 31) ? 1 : 0, n_users <= (1ll << 24) ? 1 : 0);
      ii_zero_hot_kernel<<<sms * 4, 256, 0, stream>>>(hot_rows, hot_done, counters + 128, n_items, max_hot);
      launches += 2;
    }
    ii_block_scatter_kernel<<<kSortCtas, kAccThreads, hist_smem, stream>>>(indptr, indices, values, n_users, nnz, n_items, block_hist,
                                                                          col_ptr, csc, counters + 128, packed);
    launches += 3;
  } else {
    if (nnz > 0) {
      ii_hist_kernel<<<static_cast<int>((nnz + 255) / 256), 256, 0, stream>>>(indices, values, nnz, col_cnt, sq);
      ++launches;
    }
    ii_scan_kernel<<<1, 1024, 0, stream>>>(col_cnt, n_items, col_ptr, cursor);
    ii_sqrt_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(sq, n_items);
    launches += 2;
    if (n_users > 0) {
      ii_scatter_kernel<<<static_cast<int>((n_users * 32 + 255) / 256), 256, 0, stream>>>(indptr, indices, values, n_users,
                                                                                        cursor, csc);
      ++launches;
    }
    plan_units();
  }

  // ---- 3. accumulate + select.  Accumulator range: as many items as fit beside the survivor buffer and warp 0's merger;
  // two CTAs per SM whenever the accumulator leaves room (config 3: 100,864 + 10,240 B per CTA)
  const int max_smem = 227 * 1024;
  const int side_bytes = (kSelBuf + kCap) * 8;
  int range = (max_smem - side_bytes - 1024) / 4;
  if (range > n_items) range = n_items;
  range = (range + 31) / 32 * 32;
  const size_t smem = static_cast<size_t>(range) * 4 + side_bytes;
  const bool two = smem <= 112 * 1024;
#define BB_II_LAUNCH(T, MB, ONE)                                                                                           \
  do {                                                                                                                     \
    e = cudaFuncSetAttribute(ii_accumulate_kernel<T, MB, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                             static_cast<int>(smem));                                                                      \
    if (e != cudaSuccess) return e;                                                                                        \
    ii_accumulate_kernel<T, MB, ONE><<<sms * MB, T, smem, stream>>>(                                                       \
        indptr, indices, values, packed, col_ptr, csc, sq, unit_tab, counters + 128, n_items, range, chunk, k, max_hot,   \
        hot_rows, hot_done, counters + 130, out_scores, out_ids);                                                          \
  } while (0)
  // (measured at config 3: 2 x 512 threads beat 2 x 768 by 25% and 2 x 1024 -- 32 registers, spills -- by more)
  if (two) BB_II_LAUNCH(512, 2, true);
  else if (range >= n_items) BB_II_LAUNCH(1024, 1, true);
  else BB_II_LAUNCH(1024, 1, false);
#undef BB_II_LAUNCH
  ++launches;
  if (n_launches != nullptr) *n_launches = launches;
  return cudaGetLastError();
}

}  // namespace bb
