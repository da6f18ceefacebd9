// Shared device-side pieces of the scoring engine: candidate keys, the warp
// bitonic sorter, per-query candidate buffers with a running threshold, and the
// hard-constraint predicate.  Used by both scan kernels (tcgen05 bf16 and FFMA
// fp32) and by the merge / re-score kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/brickscore.h"

namespace bb {

constexpr int kCap = 256;          // candidate slots per (catalog split, query)
constexpr int kEpl = kCap / 32;    // keys per lane in the warp sorter
constexpr int kCapTc = 512;        // tcgen05 scan: one buffer per (split, query) shared by the four column-slice warps
constexpr int kMaxSel = 128;       // largest k' kept per buffer after a compaction
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- candidate keys
// 64-bit key = (order-preserving score bits << 32) | ~id : larger key == better
// (higher score, then LOWER id).  Key 0 is below every real candidate.
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
constexpr uint32_t kOrdNegInf = 0x007fffffu;   // f2ord(-inf)
__device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
  return (static_cast<uint64_t>(f2ord(score)) << 32) | static_cast<uint32_t>(~id);
}
__device__ __forceinline__ float key_score(uint64_t k) { return ord2f(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ int32_t key_id(uint64_t k) { return static_cast<int32_t>(~static_cast<uint32_t>(k)); }

// ---------------------------------------------------------------- warp bitonic sort (descending)
// 32*EPL keys distributed as element e = i*32 + lane  (i = register index).
template <int EPL, typename K = uint64_t>
__device__ __forceinline__ void warp_sort_desc(K (&v)[EPL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * EPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int ji = j >> 5;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          const int ip = i ^ ji;
          if (ip > i) {
            const bool desc = (((i << 5) & k) == 0);   // k >= 64 here: the bit lives in i
            const K a = v[i], b = v[ip];
            const bool sw = desc ? (a < b) : (a > b);
            v[i] = sw ? b : a;
            v[ip] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          const int e = (i << 5) | lane;
          const K a = v[i];
          const K b = __shfl_xor_sync(kFull, a, j);
          const bool lower = (lane & j) == 0;
          const bool desc = ((e & k) == 0);
          const bool keep_max = (lower == desc);
          const K mx = a > b ? a : b, mn = a > b ? b : a;
          v[i] = keep_max ? mx : mn;
        }
      }
    }
  }
}

// Final phase of the bitonic network: turns a BITONIC sequence of 32*EPL keys into a descending one.
template <int EPL, typename K = uint64_t>
__device__ __forceinline__ void warp_bitonic_merge_desc(K (&v)[EPL], int lane) {
#pragma unroll
  for (int j = 16 * EPL; j > 0; j >>= 1) {
    if (j >= 32) {
      const int ji = j >> 5;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const int ip = i ^ ji;
        if (ip > i) {
          const K a = v[i], b = v[ip];
          const bool sw = a < b;
          v[i] = sw ? b : a;
          v[ip] = sw ? a : b;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const K a = v[i];
        const K b = __shfl_xor_sync(kFull, a, j);
        const bool lower = (lane & j) == 0;
        const K mx = a > b ? a : b, mn = a > b ? b : a;
        v[i] = lower ? mx : mn;
      }
    }
  }
}

// ---------------------------------------------------------------- scan parameters
struct ScanParams {
  // catalog shard
  int64_t n_items;        // local rows
  int64_t id_offset;      // global id of local row 0
  const int2* meta;       // {num_parts, (uint16)year | theme_id << 16}
  // queries
  int32_t n_queries;      // B (rows beyond it are padding)
  int32_t b_pad;          // padded row count (candidate buffer stride)
  int32_t k_pad;          // operand width (multiple of 64)
  int32_t k_sel;          // k' kept at compaction (multiple of 32, <= kMaxSel)
  const bb_constraint* cons;
  int32_t cons_stride;    // 0 = one constraint shared by all rows, 1 = per row
  const int32_t* self_id; // global ids
  const uint32_t* theme_sets;
  int32_t theme_words;
  const uint32_t* allow_bitmaps;
  int64_t allow_words;
  const int64_t* excl_off;
  const int32_t* excl_ids;
  const int32_t* cons_index;      // per-row index into cons[] (distinct constraint sets) or nullptr
  const int32_t* q_excl;          // per-QUERY exclusion-list indices [n_queries x q_excl_n] (-1 none): the user's rated items
  int32_t q_excl_n;               // 0..3
  const uint32_t* class_valid;    // per distinct constraint set: validity bitmap over local rows (mask kernel) or nullptr
  int64_t class_words;            // words per class bitmap
  // candidate buffers: slot = (split * n_sub + sub) * b_pad + row; cand[slot * cap + i], cnt[slot]
  int32_t cap;            // keys per candidate buffer: kCap (exact fp32 scan) or kCapTc (tcgen05 scan)
  int32_t n_sub;          // buffers per (split, row): the tcgen05 epilogue uses one per column half
  uint64_t* cand;
  int32_t* cand_cnt;
  int32_t* pub_cnt;       // per slot: number of SORTED keys at the front of the buffer (published at compaction)
  // per-row shared running threshold: (search epoch << 32) | ordered score bits.  Raised (atomicMax) by every stream
  // of the row on this GPU and -- row-sharded catalogs -- by the other ranks through NVLink peer memory: a rank's
  // local k'-th best is a valid lower bound of the GLOBAL k'-th best, so every shard may prune with it.  Words of
  // another epoch (an earlier search, a rank that runs ahead) are ignored by readers and lose every atomicMax.
  unsigned long long* thr_g;
  // Union thresholds of a row-sharded catalog (collective searches).  Besides thr_g every rank keeps three LEVEL words
  // per row, same (epoch | ordered score) encoding, at thr_g + (1 + l) * lvl_stride + row, l = 0..2: "this rank holds at
  // least k_sel >> (3 - l) valid candidates with score >= value" (raised at every compaction from the stream's sorted
  // prefix).  If `need` ranks each hold >= m candidates >= T and need * m >= k_sel then T is a lower bound of the
  // GLOBAL k_sel-th best: the scan's poller warp reads the peers' words through NVLink peer memory (plain loads, no
  // remote atomics), takes for each level the (8 >> l)-th largest claim over the ranks and raises the LOCAL thr_g.
  // thr_g itself holds bounds that are valid globally, so the max over all ranks' thr_g words is taken as well.
  const unsigned long long* peer_thr[7];   // the same block on the peer GPUs (cudaIpc / peer mappings), read-only here
  int32_t n_peers;
  int64_t lvl_stride;                // 0 = private search (no level words)
  uint32_t epoch;
  // pilot pass (threshold seeding): the epilogue writes, per row, the maximum VALID score of every group of 8 columns of
  // the leading pilot tiles to pilot_max[row * pilot_groups + g] instead of collecting candidates
  float* pilot_max;
  int32_t pilot_groups;
  int32_t pilot_mode;                // 1 = this launch is the pilot pass
  int32_t skip_unit_end;             // 1 = units do not sort / publish / refine at their end (no later unit would use it)
  // exact-fp32 search on the tensor pipe: rows whose tensor-pass result could not be PROVEN complete are re-scanned by the
  // FFMA kernel; row_active[q] != 0 marks them (nullptr = every row is scanned)
  const int32_t* row_active;
  // schedule
  int32_t n_qblocks;      // query blocks
  int32_t n_splits;       // catalog splits (units = n_qblocks * n_splits)
  int32_t n_tiles;        // catalog tiles of tile_items rows covered by this launch (the pilot pass covers the leading ones only)
  int32_t tile_items;
  int32_t n_slot_splits;  // candidate-buffer rows per (sub, query) = n_splits of the main scan
  int32_t n_units;        // unit indices to walk
  // debug
  float* dump;            // dense [b x n_items] score dump (bb_debug_scores) or nullptr
  unsigned long long* stat_candidates;
  int32_t* err_flag;      // set by watchdogs before trapping
  uint32_t* wave_counter; // producers that finished issuing a unit's loads (keeps the CTA groups of a wave in step)
  int32_t debug_flags;    // perf triage (env BB_DEBUG_FLAGS): 1 = epilogue never inserts, 2 = epilogue skips TMEM loads,
                          // 4 = no wave barrier, 8 = no cross-stream refinement at unit end
};

__device__ __forceinline__ void split_tile_range(const ScanParams& P, int split, int& t0, int& t1) {
  const int64_t n = P.n_tiles;
  t0 = static_cast<int>((n * split) / P.n_splits);
  t1 = static_cast<int>((n * (split + 1)) / P.n_splits);
}

// Work unit u -> (query block, candidate-slot row, tile range); units are split-major so that the query blocks of a
// split stream the same catalog region at the same time.
struct UnitInfo {
  int qb, split_slot, t0, t1;
};
__device__ __forceinline__ UnitInfo decode_unit(const ScanParams& P, int u) {
  UnitInfo r;
  r.qb = u % P.n_qblocks;
  r.split_slot = u / P.n_qblocks;
  split_tile_range(P, r.split_slot, r.t0, r.t1);
  return r;
}

// ---------------------------------------------------------------- candidate exchange of a row-sharded catalog
// Every rank exports one block of device memory (cudaIpc / peer access): threshold + level words, arrival counters and
// an inbox.  Query q of a collective search is OWNED by rank q / slice; after its local merge every rank stores its k
// final keys of q straight into the owner's inbox (peer stores over NVLink, fused into merge_emit) and bumps the
// owner's arrival counter; the owner merges the `world` lists of its slice.  Two inbox halves (search parity): a rank
// can be at most one collective search ahead of a peer.
struct CommParams {
  unsigned long long* inbox[8];      // per rank: inbox base, [2 parities][8 senders][slice_cap][k_cap] keys
  unsigned long long* arrivals[8];   // per rank: [2] cumulative arrival counters
  int32_t rank, world;
  int32_t slice;                     // queries per owner = ceil(n_queries / world)
  int32_t slice_cap, k_cap;          // inbox geometry (fixed at creation)
  int32_t parity;
};
__device__ __forceinline__ int64_t inbox_offset(const CommParams& C, int parity, int sender, int local_q) {
  return ((static_cast<int64_t>(parity) * 8 + sender) * C.slice_cap + local_q) * C.k_cap;
}

// ---------------------------------------------------------------- shared threshold words
__device__ __forceinline__ uint32_t thr_load(const ScanParams& P, int row) {
  const unsigned long long v = __ldcg(P.thr_g + row);
  return (static_cast<uint32_t>(v >> 32) == P.epoch) ? static_cast<uint32_t>(v) : kOrdNegInf;
}
__device__ __forceinline__ void thr_raise(const ScanParams& P, int row, uint32_t ord) {
  const unsigned long long v = (static_cast<unsigned long long>(P.epoch) << 32) | ord;
  atomicMax(P.thr_g + row, v);        // local only: peers PULL these words (poller warp of the tcgen05 scan)
}
// level claim l (0..2) of this rank for `row`: at least k_sel >> (3 - l) candidates score >= ord
__device__ __forceinline__ void lvl_raise(const ScanParams& P, int row, int l, uint32_t ord) {
  const unsigned long long v = (static_cast<unsigned long long>(P.epoch) << 32) | ord;
  atomicMax(P.thr_g + (1 + l) * P.lvl_stride + row, v);
}

// ---------------------------------------------------------------- hard-constraint predicate
__device__ __forceinline__ bool bit_test(const uint32_t* words, int64_t bit) {
  return (__ldg(words + (bit >> 5)) >> (bit & 31)) & 1u;
}

__device__ __forceinline__ bool list_contains(const int32_t* ids, int64_t lo, int64_t hi, int32_t gid) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t v = __ldg(ids + mid);
    if (v == gid) return true;
    if (v < gid) lo = mid + 1; else hi = mid;
  }
  return false;
}

// Integer restatement of the reference's constraint SQL (hard_constraint_filter.py:318-480)
// evaluated for one (query row, local item).
__device__ __forceinline__ bool constraint_ok_meta(const ScanParams& P, const bb_constraint& c, int64_t item, const int2 m) {
  if (c.flags & BB_C_ALWAYS_FALSE) return false;
  const int32_t parts = m.x;
  const int32_t year = static_cast<int16_t>(m.y & 0xffff);
  const int32_t theme = m.y >> 16;                 // arithmetic shift: -1 stays -1 (SQL NULL)
  bool ok = (parts >= c.parts_min) & (parts <= c.parts_max) & (year >= c.year_min) & (year <= c.year_max);
  if (!ok) return false;
  if (c.theme_set >= 0) {
    if (theme < 0 || theme >= P.theme_words * 32) return false;
    if (!bit_test(P.theme_sets + static_cast<int64_t>(c.theme_set) * P.theme_words, theme)) return false;
  }
  if (c.allow_bitmap >= 0) {
    if (!bit_test(P.allow_bitmaps + static_cast<int64_t>(c.allow_bitmap) * P.allow_words, item)) return false;
  }
  const int32_t gid = static_cast<int32_t>(item + P.id_offset);
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const int32_t l = c.excl_list[e];
    if (l >= 0 && list_contains(P.excl_ids, __ldg(P.excl_off + l), __ldg(P.excl_off + l + 1), gid)) return false;
  }
  return true;
}
__device__ __forceinline__ bool constraint_ok(const ScanParams& P, const bb_constraint& c, int64_t item) {
  if (c.flags & BB_C_ALWAYS_FALSE) return false;
  return constraint_ok_meta(P, c, item, __ldg(P.meta + item));
}

// The query's own exclusion lists (already-rated items, recommendation_system.py:441-451).
__device__ __forceinline__ bool query_excluded(const ScanParams& P, int row, int64_t item) {
  const int32_t gid = static_cast<int32_t>(item + P.id_offset);
  for (int e = 0; e < P.q_excl_n; ++e) {
    const int32_t l = __ldg(P.q_excl + static_cast<int64_t>(row) * P.q_excl_n + e);
    if (l >= 0 && list_contains(P.excl_ids, __ldg(P.excl_off + l), __ldg(P.excl_off + l + 1), gid)) return true;
  }
  return false;
}

// Validity of local `item` for query `row` (bounds, self exclusion, constraints).
static __device__ __noinline__ bool item_ok(const ScanParams& P, int row, int64_t item) {
  if (item >= P.n_items) return false;
  if (P.self_id != nullptr && __ldg(P.self_id + row) == static_cast<int32_t>(item + P.id_offset)) return false;
  if (P.q_excl_n > 0 && query_excluded(P, row, item)) return false;
  if (P.cons == nullptr) return true;
  const int64_t ci = P.cons_index != nullptr ? __ldg(P.cons_index + row) : static_cast<int64_t>(row) * P.cons_stride;
  const bb_constraint c = P.cons[ci];
  return constraint_ok(P, c, item);
}

// ---------------------------------------------------------------- per-row candidate buffer
// One thread owns a (query row, stream): `thr` is a lower bound of the row's k'-th best
// valid score, `cnt` the number of keys in its buffer.  A score passes when
// score >= thr; passing + valid candidates are appended; when a buffer could
// overflow the warp sorts it and keeps the best k_sel, which raises thr.
struct RowState {
  float thr;
  int cnt;
  uint64_t* buf;
  int64_t slot;           // (split * n_sub + sub) * b_pad + row
  int sorted;             // keys at the front of buf that are already in descending order (from the last compaction)
};

__device__ __forceinline__ int32_t ld_acquire_s32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Warp-cooperative compaction of ONE row's buffer (all 32 lanes call with the
// same buf / cnt).  Returns the new count; *new_thr gets the k_sel-th best score
// (or -inf when fewer than k_sel candidates exist).  The sort network is sized to cnt.
template <int EPL>
__device__ __forceinline__ uint64_t warp_compact_n(uint64_t* buf, int cnt, int k_sel, int lane) {
  uint64_t v[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int e = (i << 5) | lane;
    v[i] = (e < cnt) ? __ldcg(buf + e) : 0ull;
  }
  warp_sort_desc<EPL>(v, lane);
  const int kw = k_sel >> 5;
  uint64_t kth = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    if (i < kw) {
      __stcg(buf + ((i << 5) | lane), v[i]);
      if (i == kw - 1) kth = v[i];
    }
  }
  return __shfl_sync(kFull, kth, 31);
}

// Tail sort + merge: the first kMaxSel keys are already sorted (previous compaction), only the <= 128 newly
// appended keys need sorting; one compare against the reversed tail yields the best kMaxSel of the union as a
// bitonic sequence, one merge phase orders them.  ~3-4x fewer compare-exchanges than re-sorting all 256 keys.
template <int TE>
__device__ __forceinline__ uint64_t warp_compact_merge(uint64_t* buf, int cnt, int lane) {
  constexpr int PE = kMaxSel / 32;
  uint64_t p[PE], t[PE];
#pragma unroll
  for (int i = 0; i < PE; ++i) p[i] = __ldcg(buf + ((i << 5) | lane));
  {
    uint64_t tt[TE];
#pragma unroll
    for (int i = 0; i < TE; ++i) {
      const int e = kMaxSel + ((i << 5) | lane);
      tt[i] = (e < cnt) ? __ldcg(buf + e) : 0ull;
    }
    warp_sort_desc<TE>(tt, lane);
#pragma unroll
    for (int i = 0; i < PE; ++i) t[i] = (i < TE) ? tt[i < TE ? i : 0] : 0ull;
  }
  // c[e] = max(p[e], t[L-1-e]): element L-1-e lives in register PE-1-i of lane 31-lane
#pragma unroll
  for (int i = 0; i < PE; ++i) {
    const uint64_t r = __shfl_sync(kFull, t[PE - 1 - i], 31 - lane);
    p[i] = p[i] > r ? p[i] : r;
  }
  warp_bitonic_merge_desc<PE>(p, lane);
#pragma unroll
  for (int i = 0; i < PE; ++i) __stcg(buf + ((i << 5) | lane), p[i]);
  return __shfl_sync(kFull, p[PE - 1], 31);
}

__device__ __forceinline__ int warp_compact(uint64_t* buf, int cnt, int k_sel, int sorted, int lane, float* new_thr) {
  uint64_t kth;
  if (k_sel == kMaxSel && sorted == kMaxSel && cnt > kMaxSel) {
    const int tail = cnt - kMaxSel;
    if (tail <= 32) kth = warp_compact_merge<1>(buf, cnt, lane);
    else if (tail <= 64) kth = warp_compact_merge<2>(buf, cnt, lane);
    else if (tail <= 128) kth = warp_compact_merge<4>(buf, cnt, lane);
    else if (tail <= 256) kth = warp_compact_merge<8>(buf, cnt, lane);
    else kth = warp_compact_n<kCapTc / 32>(buf, cnt, k_sel, lane);
  } else if (cnt <= 32) kth = warp_compact_n<1>(buf, cnt, k_sel, lane);
  else if (cnt <= 64) kth = warp_compact_n<2>(buf, cnt, k_sel, lane);
  else if (cnt <= 128) kth = warp_compact_n<4>(buf, cnt, k_sel, lane);
  else if (cnt <= kCap) kth = warp_compact_n<kEpl>(buf, cnt, k_sel, lane);
  else kth = warp_compact_n<kCapTc / 32>(buf, cnt, k_sel, lane);
  *new_thr = (cnt >= k_sel) ? key_score(kth) : -INFINITY;
  __syncwarp();
  return cnt < k_sel ? cnt : k_sel;
}

// Cross-stream threshold refinement.  A query row is scanned by many independent streams (catalog splits x
// column halves), each keeping its own sorted top-k' prefix.  If `need` streams each hold at least r keys
// >= T then at least need*r >= k' valid items score >= T, so T is a valid lower bound of the row's k'-th best.
// With n published streams the useful levels are the smallest r = 2^l with ceil(k'/r) <= n and the next one;
// T_r = the ceil(k'/r)-th largest of the streams' r-th best keys and the bound is max_r T_r (r = k' is each
// stream's own threshold).  This lets ~74 streams per row share a threshold close to the global one instead of
// the max of their local ones.
constexpr int kRefineSlots = 128;
// best[3] = bound of the row's k'-th best (0 = none).  Collective searches (P.lvl_stride > 0) also derive the rank's
// LEVEL claims best[l], l = 0..2 ("this rank holds >= k' >> (3 - l) candidates >= best[l]"): if j streams each hold >= r
// keys >= T the rank holds >= j * r of them, so level m takes the ceil(m / r)-th largest of the streams' r-th best keys.
template <int kPer>
__device__ __forceinline__ void warp_refine_threshold_n(const ScanParams& P, int row, int lane, uint32_t (&best)[4]) {
  best[0] = best[1] = best[2] = best[3] = 0u;
  const int n_slots = P.n_slot_splits * P.n_sub;
  if (n_slots <= 1) return;
  constexpr int kSlots = kPer * 32;               // streams examined (sort network size)
  const int stride = (n_slots + kSlots - 1) / kSlots;
  const int n_use = (n_slots + stride - 1) / stride;
  int c[kPer];
  int64_t slot[kPer];
  int n_pub = 0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int s = (lane + 32 * i) * stride;
    c[i] = 0;
    slot[i] = 0;
    if (lane + 32 * i < n_use && s < n_slots) {
      slot[i] = static_cast<int64_t>(s) * P.b_pad + row;
      c[i] = __ldcg(P.pub_cnt + slot[i]);
    }
    n_pub += __popc(__ballot_sync(kFull, c[i] > 0));
  }
  __threadfence();                                  // keys read below are at least as new as the counts read above
  if (n_pub == 0) return;
  int r0 = 2;
  while (r0 < P.k_sel && (P.k_sel + r0 - 1) / r0 > n_pub) r0 <<= 1;
  // private searches: the two useful levels of the k' bound; collective searches start at the smallest r that can
  // serve the k'/8 claim
  int r_lo = r0;
  const bool levels = P.lvl_stride > 0;
  if (levels) {
    const int m0 = P.k_sel >> 3;
    r_lo = 1;
    while (r_lo < m0 && (m0 + r_lo - 1) / r_lo > n_pub) r_lo <<= 1;
  }
#pragma unroll 1
  for (int r = r_lo; r <= 2 * r0 && r <= P.k_sel; r <<= 1) {
    uint32_t v[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i)
      v[i] = (c[i] >= r) ? static_cast<uint32_t>(__ldcg(P.cand + slot[i] * P.cap + (r - 1)) >> 32) : 0u;
    warp_sort_desc<kPer, uint32_t>(v, lane);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (l < 3 && !levels) continue;
      if (l == 3 && (r < r0 || r >= P.k_sel)) continue;   // (the r = k' level is the stream's own threshold)
      const int m = P.k_sel >> (3 - l);
      const int e = (m + r - 1) / r - 1;              // element e lives in register e>>5 of lane e&31
      if (e >= kSlots) continue;
      uint32_t t = 0u;
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if ((e >> 5) == i) t = v[i];
      t = __shfl_sync(kFull, t, e & 31);
      best[l] = t > best[l] ? t : best[l];
    }
  }
}

// (the sort network follows the number of streams of a row: 23 at the 8-GPU shard size, 60 on one GPU)
__device__ __forceinline__ void warp_refine_threshold(const ScanParams& P, int row, int lane, uint32_t (&best)[4]) {
  const int n_slots = P.n_slot_splits * P.n_sub;
  if (n_slots <= 32) warp_refine_threshold_n<1>(P, row, lane, best);
  else if (n_slots <= 64) warp_refine_threshold_n<2>(P, row, lane, best);
  else warp_refine_threshold_n<kRefineSlots / 32>(P, row, lane, best);
}

// Compacts the buffers of the rows in `need0` (a warp-uniform lane mask; lane l holds row l's RowState): sort /
// merge, publish the sorted prefix, optionally refine across streams, raise the shared threshold.
__device__ __forceinline__ int warp_compact_rows(const ScanParams& P, RowState& st, int row, int lane, unsigned need0,
                                                 unsigned refine_mask) {
  if (need0 == 0u) return 0;
  // phase 1: sort / merge every flagged row's buffer (warp-cooperative, one row at a time)
  float my_nt = -INFINITY;
  int my_nc = st.cnt;
  for (unsigned need = need0; need; need &= need - 1) {
    const int src = __ffs(need) - 1;
    uint64_t* b = reinterpret_cast<uint64_t*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(st.buf), src));
    const int c = __shfl_sync(kFull, st.cnt, src);
    const int c_sorted = __shfl_sync(kFull, st.sorted, src);
    float nt;
    const int nc = warp_compact(b, c, P.k_sel, c_sorted, lane, &nt);
    if (lane == src) {
      my_nt = nt;
      my_nc = nc;
    }
  }
  // phase 2: ONE fence for all rows (sorted keys are visible before the counts that advertise them), then every
  // flagged lane publishes its own count
  __threadfence();
  const bool mine = (need0 >> lane) & 1u;
  if (mine) {
    atomicMax(P.pub_cnt + st.slot, my_nc);
    st.cnt = my_nc;
    st.sorted = my_nc;
  }
  __syncwarp();
  // phase 3: cross-stream refinement for the rows that asked for it, then raise the shared threshold (and, in
  // collective searches, this rank's level claims: the stream's own sorted prefix and the refined ones)
  uint32_t my_ref[4] = {0u, 0u, 0u, 0u};
  for (unsigned need = need0 & refine_mask; need; need &= need - 1) {
    const int src = __ffs(need) - 1;
    const int src_row = __shfl_sync(kFull, row, src);
    uint32_t refined[4];
    warp_refine_threshold(P, src_row, lane, refined);
    if (lane == src) {
#pragma unroll
      for (int l = 0; l < 4; ++l) my_ref[l] = refined[l];
    }
  }
  if (mine) {
    uint32_t o = f2ord(my_nt);
    o = my_ref[3] > o ? my_ref[3] : o;
    if (o > f2ord(st.thr)) {
      st.thr = ord2f(o);
      thr_raise(P, row, o);
    }
    if (P.lvl_stride > 0) {
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const int m = P.k_sel >> (3 - l);
        uint32_t claim = my_ref[l];
        if (my_nc >= m) {
          const uint32_t own = static_cast<uint32_t>(__ldcg(st.buf + (m - 1)) >> 32);
          claim = own > claim ? own : claim;
        }
        if (claim > kOrdNegInf) lvl_raise(P, row, l, claim);
      }
    }
  }
  return __popc(need0);
}

// After a group of <= 32 appends per lane: compact every row of this warp whose
// buffer could overflow on the next group, publish its sorted prefix, and raise the
// row's shared threshold (own k'-th best and the cross-stream refinement).
// `force_mask`: lanes whose buffer is compacted regardless of fill (unit end: leave a sorted, published prefix).
// `refine_mask`: lanes whose compaction also runs the cross-stream refinement (warp-uniform value).
// `headroom`: appends a lane may make before the next call (the buffer must not overflow in between).
__device__ __forceinline__ int warp_compact_if_needed(const ScanParams& P, RowState& st, int row, int lane,
                                                      unsigned force_mask = 0u, unsigned refine_mask = kFull, int headroom = 32) {
  __syncwarp();
  const unsigned need0 = __ballot_sync(kFull, st.cnt > kCap - headroom) | force_mask;
  return warp_compact_rows(P, st, row, lane, need0, refine_mask);
}

__device__ __forceinline__ void row_append(const ScanParams& P, RowState& st, float score, int64_t item) {
  __stcg(st.buf + st.cnt, make_key(score, static_cast<uint32_t>(item + P.id_offset)));
  ++st.cnt;
}

}  // namespace bb
