// K1 -- tcgen05 bf16 scan: TMA-fed UMMA (fp32 accumulate in TMEM) with the fused
// hybrid blend (folded into the operands) + hard-constraint mask + running-threshold
// top-k epilogue.  Score tiles never leave the SM: only candidates that beat the
// per-query threshold are written (8 B each) to L2-resident candidate buffers.
//
// Decomposition (query-stationary):
//   * a CTA group = 1 CTA (cta_group::1) or a 2-CTA cluster on one TPC (cta_group::2).
//   * the group owns 128*CG query rows; their whole bf16 operand (<= 7 k-blocks of
//     64 columns, 16 KB each, 128B-swizzled, K-major) stays RESIDENT in shared memory
//     for a unit, so L2->SM traffic is the catalog stream only.
//   * the catalog streams through a TMA ring of [256/CG rows x 64 cols] stages; one
//     UMMA tile is 128*CG queries x 256 items; two 256-column TMEM accumulators are
//     double buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
//   * work unit = (query block, catalog split); units are dealt round-robin to groups.
//
// Warp roles (640 threads): 0 = TMA producer, 1 = MMA issuer (leader CTA),
// 2 = TMEM allocator, 3 = union-threshold poller (collective searches of a row-sharded catalog: reads the peers' level
// words over NVLink peer memory; idle otherwise), 4..19 = epilogue (TMEM lane quarter = warp % 4; the four
// warps of a quarter -- a row group -- split the tile's 256 columns and share ONE candidate
// stream per (catalog split, query row): row state in shared memory, one named barrier per tile).
//
// Roofline: tensor core.  2*128*256*k_pad flop per CTA tile; catalog bytes
// 256*k_pad*2 per group tile.
#include <cuda.h>
#include <cuda_bf16.h>

#pragma once
#include "bb_common.cuh"
#include "bb_kernels.h"

namespace bb {
namespace {

constexpr int kBlockK = 64;                 // bf16 columns per k-block (= one 128 B swizzle row)
constexpr int kAkbBytes = 128 * 128;        // resident query k-block: 128 rows x 128 B
constexpr int kUmmaK = 16;
constexpr int kAccCols = kTcTile;           // fp32 accumulator columns per tile
// Epilogue cycle breakdown (bb_info.epilogue_cycles): compiled in only with -DBB_EPILOGUE_PROFILE=1
// (BB_EPILOGUE_PROFILE=1 python -m brickbrain_rec_engine_b200.build) -- the clock reads cost ~2% of the scan.
#ifndef BB_EPILOGUE_PROFILE
#define BB_EPILOGUE_PROFILE 0
#endif
__device__ __forceinline__ long long epi_clock() {
#if BB_EPILOGUE_PROFILE
  return clock64();
#else
  return 0;
#endif
}
constexpr int kAccStages = 512 / kAccCols;  // TMEM accumulator ring: how far the MMA issuer may run ahead of the epilogue
constexpr int kTmemCols = 512;              // two accumulators
constexpr int kColWarps = 4;                // epilogue warps per TMEM lane quarter (row group), each scans a column slice of every tile
constexpr int kEpiWarps = 4 * kColWarps;
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kChunksPerWarp = (kAccCols / 32) / (kEpiWarps / 4);
// the four column-slice warps of a row group append to ONE buffer per row; it is checked once per tile, and a tile
// can add at most kTcTile keys to a row
constexpr int kCompactLine = kCapTc - kTcTile;   // a row whose count crosses this line is compacted at the tile's end
static_assert(kCompactLine >= kMaxSel, "candidate buffer too small: a compacted row (<= kMaxSel keys) must be below the line");
static_assert(kTcSubStreams == 1, "one candidate stream per (split, row)");
constexpr int kColMetaBytes = kEpiWarps * 32 * 16;   // per-warp packed metadata of the 32 columns of a chunk
constexpr int kRowStateBytes = 3 * 128 * 4 + 128;   // per row: count, sorted prefix, threshold; + compaction flags [3 tiles][4 row groups] + poller word
constexpr int kThemeBytes = 32 * 128 * 4;         // per-row allowed-theme bitmaps, word-major [32 words][128 rows]
constexpr long long kWatchdogCycles = 6000000000ll;   // ~3 s: never hang the GPU, trap instead
constexpr long long kWaveTimeoutCycles = 40000000ll;  // ~20 ms: the grid is not co-resident, give the wave barrier up
constexpr int kPublishTiles = 64;                     // periodic publication of what a stream learnt (power of two)
constexpr int kPublishMinNew = 12;                    // ... for rows with at least this many keys since their last compaction

// kRowCons: the batch carries per-row constraint descriptors that are prefiltered in the epilogue (needs the
// per-row allowed-theme bitmaps in shared memory); otherwise that space is a sixth TMA stage.
template <int CG, bool kRowCons> struct Cfg {
  static constexpr int kBRows = kTcTile / CG;            // catalog rows this CTA loads per stage
  static constexpr int kBStageBytes = kBRows * 128;
  static constexpr int kStages = (CG == 2) ? (kRowCons ? 80 : 96) * 1024 / kBStageBytes
                                           : (kRowCons ? 64 : 96) * 1024 / kBStageBytes;
  static constexpr int kThemeSmem = kRowCons ? kThemeBytes : 0;
  static constexpr int kSmemBytes =
      kTcMaxKBlocks * kAkbBytes + kStages * kBStageBytes + kColMetaBytes + kThemeSmem + 512 + kRowStateBytes + 1024;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // default semantics (release at CTA scope), as CUTLASS' ClusterBarrier::arrive: a cluster-scope release would
  // fence every outstanding global candidate store of the warp (ERRBAR) once per tile
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
template <bool kHint = false>
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  if constexpr (kHint) {
    // suspend-time hint (ns): the (many) epilogue warps sleep in hardware instead of burning issue slots and power
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  return ok;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
template <bool kHint = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* err_flag, int code) {
  if (mbar_try_wait<kHint>(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait<kHint>(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) {
      if (err_flag != nullptr) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}

__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint32_t bar, uint32_t dst, int32_t c0, int32_t c1) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
  } else {
    // bar is the LEADER CTA's barrier in the shared::cluster window: both CTAs' bytes
    // complete one transaction barrier that the single MMA issuer waits on.
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
  }
}

// peer memory over NVLink: system-scope relaxed load (never served from this SM's L1)
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// tcgen05.commit: arrives on `bar` (same smem offset in every CTA of the group) once all
// previously issued MMAs of this thread have completed.
template <int CG> __device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (rows of 64 bf16 = 128 B,
// 8-row groups 1024 B apart).  Field layout: start[0,14) | LBO[16,30) | SBO[32,46) |
// version[46,48)=1 | layout[61,64)=2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------ the kernel
// bar.sync on a named barrier: the four column-slice warps of a row group (128 threads)
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// Compaction of the rows in `need` (bit = row of the row group `ew`), split between the group's four warps: warp
// `half` takes every fourth flagged row.  Row state lives in shared memory; lane l of EVERY warp of the group
// addresses row l's buffer (`buf`, `slot`, `q`), so a row's pointers are one shuffle away.  Called between two
// barriers of the row group: nobody appends meanwhile.
__device__ __forceinline__ int group_compact(const ScanParams& P, unsigned need, unsigned refine_mask, int half, int* s_cnt,
                                             int* s_sorted, float* s_thr, int ew, uint64_t* buf, int64_t slot, int q, int lane) {
  unsigned mine = 0u;
  int i = 0;
  for (unsigned m = need; m; m &= m - 1, ++i)
    if ((i & 3) == half) mine |= m & (0u - m);
  if (mine == 0u) return 0;
  const int row_in_cta = ew * 32 + lane;
  RowState st;
  st.buf = buf;
  st.slot = slot;
  st.cnt = s_cnt[row_in_cta];
  st.sorted = s_sorted[row_in_cta];
  st.thr = s_thr[row_in_cta];
  const int n = warp_compact_rows(P, st, q, lane, mine, refine_mask & mine);
  if ((mine >> lane) & 1u) {
    s_cnt[row_in_cta] = st.cnt;
    s_sorted[row_in_cta] = st.sorted;
    s_thr[row_in_cta] = st.thr;
  }
  return n;
}

// Union threshold of one query row over the ranks of a row-sharded catalog (see ScanParams::lvl_stride): word 0 of every
// rank is a bound that is already valid globally (max over ranks), level word l (1..3 in memory, claim "this rank
// holds >= k' >> (3 - l) candidates >= value") contributes the (8 >> l)-th largest claim over the ranks.  Words of
// another search epoch count as "no claim".  Raises the LOCAL threshold word only.
__device__ __forceinline__ void poll_union_threshold(const ScanParams& P, int row) {
  const uint32_t ep = P.epoch;
  const int n_ranks = P.n_peers + 1;
  uint32_t best = 0u, own3 = 0u;
#pragma unroll
  for (int w = 0; w < 4; ++w) {               // w = 0: thr_g, w = 1..3: level l = w - 1
    uint32_t v[8];
    {
      const unsigned long long x = __ldcg(P.thr_g + w * P.lvl_stride + row);
      v[0] = (static_cast<uint32_t>(x >> 32) == ep) ? static_cast<uint32_t>(x) : 0u;
    }
#pragma unroll
    for (int p = 0; p < 7; ++p) {             // (constant indices: a run-time index would copy the kernel parameters to local memory)
      v[1 + p] = 0u;
      if (p < P.n_peers) {
        const unsigned long long x = ld_relaxed_sys_u64(P.peer_thr[p] + w * P.lvl_stride + row);
        v[1 + p] = (static_cast<uint32_t>(x >> 32) == ep) ? static_cast<uint32_t>(x) : 0u;
      }
    }
    if (w == 0) {
      own3 = v[0];
#pragma unroll
      for (int i = 0; i < 8; ++i) best = v[i] > best ? v[i] : best;
    } else {
      const int need = 8 >> (w - 1);          // ranks that must hold k' / need candidates each
      if (need <= n_ranks) {
        // need-th largest of the claims: the value with exactly need - 1 values ranked before it
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int before = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) before += (v[j] > v[i]) || (v[j] == v[i] && j < i);
          if (before == need - 1) best = v[i] > best ? v[i] : best;
        }
      }
    }
  }
  if (best > own3 && best > kOrdNegInf) atomicMax(P.thr_g + row, (static_cast<unsigned long long>(ep) << 32) | best);
}

// kSplit: exact-fp32 search on the tensor pipe.  Operands are split into two bf16 terms, x = hi + lo + r with
// |r| <= 2^-18 |x|, and a tile accumulates q_hi.e_hi + q_lo.e_hi + q_hi.e_lo in fp32 (the dropped terms are below
// 1.2e-5 |q||e|).  q_hi stays resident like the bf16 operand; per k-block the ring carries THREE slots -- e_hi, q_lo,
// e_lo -- so a tile streams 3x the bytes for 3x the MMAs (same bytes per MMA as the bf16 scan).  The epilogue is
// unchanged; the merge kernel re-scores the survivors in exact fp32 and PROVES the top-k complete or flags the row.
// kPilot: the threshold-seeding pass (group maxima instead of candidates); a separate instantiation keeps its code and
// registers out of the main scan.
template <int CG, bool kRowCons, bool kSplit, bool kPilot>
__global__ void __launch_bounds__(kThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_e,
               const __grid_constant__ CUtensorMap tmap_ql, const __grid_constant__ CUtensorMap tmap_el,
               const ScanParams P, const int n_groups) {
  using C = Cfg<CG, kRowCons>;
  extern __shared__ unsigned char smem_raw[];
  // 1024 B alignment (128B swizzle atoms) by pointer arithmetic on the __shared__ array, so that the compiler keeps
  // the shared address space (integer round-trips degrade every later access to generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + kTcMaxKBlocks * kAkbBytes;
  uint4* smem_colmeta = reinterpret_cast<uint4*>(smem_b + C::kStages * C::kBStageBytes);
  uint32_t* smem_theme = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(smem_colmeta) + kColMetaBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(smem_theme) + C::kThemeSmem);
  uint64_t* bar_full = bars;                      // [kStages]  TMA -> MMA     (leader's are used)
  uint64_t* bar_empty = bars + C::kStages;        // [kStages]  MMA -> TMA     (per CTA)
  uint64_t* bar_a_full = bars + 2 * C::kStages;   //            TMA -> MMA     (leader)
  uint64_t* bar_a_empty = bar_a_full + 1;         //            MMA -> TMA     (per CTA)
  uint64_t* bar_acc_full = bar_a_full + 2;                     // [kAccStages]  MMA -> epilogue (per CTA)
  uint64_t* bar_acc_empty = bar_acc_full + kAccStages;         // [kAccStages]  epilogue -> MMA (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + kAccStages);
  int* s_cnt = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(bars) + 512);   // [128] keys in the row's buffer
  int* s_sorted = s_cnt + 128;                                                       // [128] sorted prefix length
  float* s_thr = reinterpret_cast<float*>(s_sorted + 128);                           // [128] row threshold (after compactions)
  unsigned* s_flag = reinterpret_cast<unsigned*>(s_thr + 128);                       // [3][4] rows that crossed the compaction line
  volatile int* s_poll = reinterpret_cast<volatile int*>(s_flag + 12);               // first query row of this CTA's current unit (-2 not yet, -1 done)
  unsigned* s_pflag = s_flag + 16;                                                   // [3][4] rows picked for a periodic publication

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int group = blockIdx.x / CG;
  const int n_kb = P.k_pad / kBlockK;
  const int n_units = P.n_units;
  int32_t* err = P.err_flag;

  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(bar_full + s), 1);
      mbar_init(smem_u32(bar_empty + s), 1);
    }
    mbar_init(smem_u32(bar_a_full), 1);
    mbar_init(smem_u32(bar_a_empty), 1);
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(smem_u32(bar_acc_full + a), 1);
      mbar_init(smem_u32(bar_acc_empty + a), CG * kEpiWarps);
    }
    fence_mbar_init();
    *s_poll = -2;
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_e);
    if constexpr (kSplit) {
      prefetch_tmap(&tmap_ql);
      prefetch_tmap(&tmap_el);
    }
  }
  if (warp == 2) tmem_alloc<CG>(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const uint32_t a_full_leader = (CG == 2) ? mapa_u32(smem_u32(bar_a_full), 0) : smem_u32(bar_a_full);
      int wave = 0;
      for (int u = group; u < n_units; u += n_groups, ++wave) {
        const UnitInfo ui = decode_unit(P, u);
        const int qb = ui.qb, t0 = ui.t0, t1 = ui.t1;
        const int row0 = qb * (kTcRowsPerCta * CG) + static_cast<int>(cta_rank) * kTcRowsPerCta;
        // wave barrier: the query blocks that stream the same catalog split share it through L2 only while they
        // stay in step; do not start wave w before every CTA issued all loads of wave w-1.  This is a performance
        // device only: it assumes the whole grid is co-resident (one CTA per SM), which does not hold when another
        // kernel occupies SMs (a second search on another stream, MPS).  A wait that exceeds kWaveTimeout therefore
        // BREAKS the barrier for the rest of the launch (the counter jumps past every target) instead of trapping.
        if (wave > 0 && P.wave_counter != nullptr && !(P.debug_flags & 4)) {
          const uint32_t target = static_cast<uint32_t>(min(wave * n_groups, n_units)) * CG;
          const long long t0w = clock64();
          while (*reinterpret_cast<volatile uint32_t*>(P.wave_counter) < target) {
            __nanosleep(200);
            if (clock64() - t0w > kWaveTimeoutCycles) {
              atomicMax(P.wave_counter, 0x40000000u);
              break;
            }
          }
        }
        // resident query operand: wait until the previous unit's MMAs stopped reading it
        mbar_wait(smem_u32(bar_a_empty), a_phase ^ 1, err, 101);
        a_phase ^= 1;
        if (leader) mbar_arrive_expect_tx(smem_u32(bar_a_full), static_cast<uint32_t>(CG * n_kb * kAkbBytes));
        for (int kb = 0; kb < n_kb; ++kb)
          tma_load_2d<CG>(&tmap_q, a_full_leader, smem_u32(smem_a + kb * kAkbBytes), kb * kBlockK, row0);
        for (int t = t0; t < t1; ++t) {
          const int item0 = t * kTcTile + static_cast<int>(cta_rank) * C::kBRows;
          for (int kb = 0; kb < n_kb; ++kb) {
            // slots of a k-block: e_hi [, q_lo, e_lo]
#pragma unroll
            for (int part = 0; part < (kSplit ? 3 : 1); ++part) {
              mbar_wait(smem_u32(bar_empty + stage), phase ^ 1, err, 102);
              const uint32_t full_leader = (CG == 2) ? mapa_u32(smem_u32(bar_full + stage), 0) : smem_u32(bar_full + stage);
              const uint32_t dst = smem_u32(smem_b + stage * C::kBStageBytes);
              if (part == 1) {
                if (leader) mbar_arrive_expect_tx(smem_u32(bar_full + stage), static_cast<uint32_t>(CG * kAkbBytes));
                tma_load_2d<CG>(&tmap_ql, full_leader, dst, kb * kBlockK, row0);
              } else {
                if (leader) mbar_arrive_expect_tx(smem_u32(bar_full + stage), static_cast<uint32_t>(CG * C::kBStageBytes));
                tma_load_2d<CG>(part == 0 ? &tmap_e : &tmap_el, full_leader, dst, kb * kBlockK, item0);
              }
              if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
        if (P.wave_counter != nullptr) atomicAdd(P.wave_counter, 1u);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA) ==============================
    if (leader) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_phase = 0;
      constexpr uint32_t idesc = make_idesc(128 * CG, kTcTile);
      long long t_acc_wait = 0, t_full_wait = 0;     // pipeline triage: who makes the MMA issuer wait
      const long long t_begin = clock64();
      for (int u = group; u < n_units; u += n_groups) {
        const UnitInfo ui = decode_unit(P, u);
        const int t0 = ui.t0, t1 = ui.t1;
        const long long tu0 = epi_clock();
        mbar_wait(smem_u32(bar_a_full), a_phase, err, 103);
        a_phase ^= 1;
        const long long tu1 = epi_clock();
        for (int t = t0; t < t1; ++t) {
          long long tw = clock64();
          mbar_wait(smem_u32(bar_acc_empty + acc), acc_phase ^ 1, err, 104);
          t_acc_wait += clock64() - tw;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kAccCols;
          for (int kb = 0; kb < n_kb; ++kb) {
            if constexpr (!kSplit) {
              tw = clock64();
              mbar_wait(smem_u32(bar_full + stage), phase, err, 105);
              t_full_wait += clock64() - tw;
              tc_fence_after();
              if (lane == 0) {
                const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + kb * kAkbBytes));
                const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + stage * C::kBStageBytes));
#pragma unroll
                for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                  // advance 16 bf16 = 32 B inside the swizzle atom: +2 in the (>>4) start-address field
                  umma_bf16<CG>(d_tmem, a_desc + static_cast<uint64_t>(ks * 2), b_desc + static_cast<uint64_t>(ks * 2), idesc,
                                static_cast<uint32_t>((kb | ks) != 0));
                }
                umma_commit<CG>(smem_u32(bar_empty + stage));      // stage reusable once these MMAs retire
              }
              __syncwarp();
              if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            } else {
              // q_hi.e_hi, q_lo.e_hi (the e_hi slot serves both), q_hi.e_lo
              uint32_t slot[3];
#pragma unroll
              for (int part = 0; part < 3; ++part) {
                tw = clock64();
                mbar_wait(smem_u32(bar_full + stage), phase, err, 105);
                t_full_wait += clock64() - tw;
                tc_fence_after();
                slot[part] = stage;
                if (lane == 0) {
                  const uint64_t a_desc = make_smem_desc(part == 1 ? smem_u32(smem_b + slot[1] * C::kBStageBytes)
                                                                   : smem_u32(smem_a + kb * kAkbBytes));
                  const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + slot[part == 2 ? 2 : 0] * C::kBStageBytes));
#pragma unroll
                  for (int ks = 0; ks < kBlockK / kUmmaK; ++ks)
                    umma_bf16<CG>(d_tmem, a_desc + static_cast<uint64_t>(ks * 2), b_desc + static_cast<uint64_t>(ks * 2), idesc,
                                  static_cast<uint32_t>((kb | ks | part) != 0));
                  if (part == 1) {                                   // e_hi and q_lo are free once both products retired
                    umma_commit<CG>(smem_u32(bar_empty + slot[0]));
                    umma_commit<CG>(smem_u32(bar_empty + slot[1]));
                  } else if (part == 2) {
                    umma_commit<CG>(smem_u32(bar_empty + slot[2]));
                  }
                }
                __syncwarp();
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
              }
            }
          }
          if (lane == 0) umma_commit<CG>(smem_u32(bar_acc_full + acc));   // accumulator ready
          __syncwarp();
          if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) umma_commit<CG>(smem_u32(bar_a_empty));            // query operand no longer read
        __syncwarp();
#if BB_EPILOGUE_PROFILE
        if (lane == 0 && P.stat_candidates != nullptr) {      // per scheduling round: unit duration and its operand wait
          const int round = min((u - group) / n_groups, 15);
          atomicAdd(P.stat_candidates + 32 + round, static_cast<unsigned long long>(clock64() - tu0));
          atomicAdd(P.stat_candidates + 48 + round, static_cast<unsigned long long>(tu1 - tu0));
        }
#endif
      }
      if (lane == 0 && P.stat_candidates != nullptr) {
        atomicAdd(P.stat_candidates + 4, static_cast<unsigned long long>(t_acc_wait));
        atomicAdd(P.stat_candidates + 5, static_cast<unsigned long long>(t_full_wait));
        atomicAdd(P.stat_candidates + 6, static_cast<unsigned long long>(clock64() - t_begin));
      }
    }
  } else if (warp == 3) {
    // ============================== union-threshold poller (collective searches) ==============================
    if (!kPilot && P.n_peers > 0 && P.lvl_stride > 0 && P.dump == nullptr) {
      while (true) {
        const int base = *s_poll;
        if (base == -1) break;
        if (base >= 0) {
#pragma unroll 1
          for (int pass = 0; pass < kTcRowsPerCta / 32; ++pass) {
            const int row = base + pass * 32 + lane;
            if (row < P.n_queries) poll_union_threshold(P, row);
          }
        }
        // ~40 us between polls (a unit lasts hundreds), in short naps so that the warp leaves promptly at the end
#pragma unroll 1
        for (int nap = 0; nap < 20 && *s_poll != -1; ++nap) __nanosleep(2000);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================== epilogue: mask + running-threshold top-k ==============================
    const int ew = warp & 3;                        // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;               // which column slice of every tile this warp scans
    const int row_in_cta = ew * 32 + lane;
    uint4* colmeta_w = smem_colmeta + (warp - 4) * 32;
    uint32_t* theme_row = smem_theme + row_in_cta;  // this row's allowed-theme words at [w * 128]
    uint32_t acc = 0, acc_phase = 0;
    const uint32_t acc_empty_leader_base = (CG == 2) ? mapa_u32(smem_u32(bar_acc_empty), 0) : smem_u32(bar_acc_empty);
    const bool use_cons = kRowCons && P.cons != nullptr && P.class_valid == nullptr;   // per-row prefilter (no class bitmaps)
    unsigned stat_appends = 0, stat_slow = 0, stat_compact = 0, stat_wasted = 0;   // per-thread / per-warp counters (debug statistics)
    long long cyc_wait = 0, cyc_slow = 0, cyc_compact = 0, cyc_end = 0, cyc_hold = 0, cyc_ld = 0, t_full = 0;
    const long long cyc_begin = epi_clock();
    for (int u = group; u < n_units; u += n_groups) {
      const UnitInfo ui = decode_unit(P, u);
      const int qb = ui.qb, split = ui.split_slot, t0 = ui.t0, t1 = ui.t1;
      const int q = qb * (kTcRowsPerCta * CG) + static_cast<int>(cta_rank) * kTcRowsPerCta + row_in_cta;
      if (warp == 4 && lane == 0) *s_poll = q;      // (row_in_cta == 0 here: the first row of this CTA's block)
      const bool q_valid = q < P.n_queries;
      // the row's cheap predicates (integer ranges + allowed-theme bitmap) live in registers / smem so that a
      // score that beats the threshold is masked without touching global memory
      int pmin = INT32_MIN, ymin = -32768;
      unsigned prange = 0xffffffffu, yrange = 0xffffu;
      bool has_theme = false, dead = !q_valid;
      bool needs_full = false;                       // allow bitmap / exclusion lists / wide theme sets -> item_ok()
      int64_t self_local = -1;                       // local row excluded for this query (-1 none)
      if (q_valid && P.self_id != nullptr) {
        const int64_t sl = static_cast<int64_t>(__ldg(P.self_id + q)) - P.id_offset;
        if (sl >= 0 && sl < P.n_items) self_local = sl;
      }
      // The query's own exclusion lists (already-rated items).  ONE list per query (the common case) is walked like
      // a merge: the warp visits its columns in increasing order, so the lane keeps the next listed id (as a local
      // row) and a cursor; listed columns are removed from the pass bitmap BEFORE the vote and cost the slow path
      // nothing.  More than one list per query falls back to a binary search per passing candidate.
      // (the per-row-descriptor variant is short of registers: it keeps the per-candidate search)
      constexpr bool kWalk = !kRowCons;
      bool excl_only = false;
      int ex_next = INT32_MAX, ex_i = 0;             // next listed LOCAL row >= the columns seen so far; its index in the list
      if (kWalk && q_valid && P.q_excl_n == 1) {
        const int32_t l = __ldg(P.q_excl + q);
        if (l >= 0) {
          const int64_t lo = __ldg(P.excl_off + l), hi = __ldg(P.excl_off + l + 1);
          const int64_t first = P.id_offset + static_cast<int64_t>(t0) * kTcTile;
          int64_t a = lo, b = hi;                    // lower bound of the unit's first row
          while (a < b) {
            const int64_t mid = (a + b) >> 1;
            if (static_cast<int64_t>(__ldg(P.excl_ids + mid)) < first) a = mid + 1; else b = mid;
          }
          ex_i = static_cast<int>(a - lo);
          if (a < hi) {
            const int64_t r = static_cast<int64_t>(__ldg(P.excl_ids + a)) - P.id_offset;
            ex_next = r < INT32_MAX ? static_cast<int>(r) : INT32_MAX;
          }
        }
      } else if (q_valid && P.q_excl_n > (kWalk ? 1 : 0)) {
        for (int e = 0; e < P.q_excl_n; ++e) excl_only |= __ldg(P.q_excl + static_cast<int64_t>(q) * P.q_excl_n + e) >= 0;
      }
      const uint32_t* class_row = nullptr;           // this row's constraint-set validity bitmap (class mode)
      if (q_valid && P.cons != nullptr) {
        const int64_t ci = P.cons_index != nullptr ? __ldg(P.cons_index + q) : static_cast<int64_t>(q) * P.cons_stride;
        const bb_constraint c = P.cons[ci];
        if (P.class_valid != nullptr) class_row = P.class_valid + ci * P.class_words;
        needs_full = P.class_valid != nullptr ? false : c.allow_bitmap >= 0 || c.excl_list[0] >= 0 || c.excl_list[1] >= 0 || c.excl_list[2] >= 0 ||
                     (c.theme_set >= 0 && P.theme_words > 32);   // (class bitmaps already include all of these)
        if ((c.flags & BB_C_ALWAYS_FALSE) || c.parts_max < c.parts_min || c.year_max < c.year_min) dead = true;
        pmin = c.parts_min;
        prange = static_cast<unsigned>(c.parts_max) - static_cast<unsigned>(c.parts_min);
        ymin = c.year_min;
        yrange = static_cast<unsigned>(static_cast<int>(c.year_max) - static_cast<int>(c.year_min));
        if (use_cons && c.theme_set >= 0 && P.theme_words <= 32) {
          has_theme = true;
          const uint32_t* src = P.theme_sets + static_cast<int64_t>(c.theme_set) * P.theme_words;
          for (int w = 0; w < 32; ++w) theme_row[w * 128] = (w < P.theme_words) ? __ldg(src + w) : 0u;   // both halves write the same words
        }
      }
      if (needs_full) excl_only = false;             // item_ok() covers the query's lists too
      // ---- one candidate stream per (split, row), shared by the row group's four column-slice warps: count,
      //      sorted-prefix length and threshold live in shared memory, appends take their slot with a shared-memory
      //      atomic; the four warps meet at a named barrier once per tile, where rows that crossed the compaction
      //      line are compacted (split between the four warps) before anyone appends again
      const int64_t slot = static_cast<int64_t>(split) * P.b_pad + q;
      uint64_t* const buf = P.cand + slot * kCapTc;
      const int bar_id = 1 + ew;
      if (half == 0) {
        s_cnt[row_in_cta] = 0;
        s_sorted[row_in_cta] = 0;
        s_thr[row_in_cta] = dead ? INFINITY : -INFINITY;
        if (lane < 3) s_flag[lane * 4 + ew] = 0u;
      }
      named_bar_sync(bar_id, kColWarps * 32);
      float thr = dead ? INFINITY : -INFINITY;
      const bool share_thr = !dead && P.dump == nullptr;
      uint32_t thr_next = share_thr ? thr_load(P, q) : kOrdNegInf;   // shared threshold, fetched one tile ahead
      int tslot = 0;                                 // tile index mod 3: which flag word this tile's appends raise
      for (int t = t0; t < t1; ++t) {
        thr = fmaxf(thr, s_thr[row_in_cta]);
        if (share_thr) {
          thr = fmaxf(thr, ord2f(thr_next));
          thr_next = thr_load(P, q);
        }
        {
          const long long c0 = epi_clock();
          mbar_wait<true>(smem_u32(bar_acc_full + acc), acc_phase, err, 106);
          t_full = epi_clock();
          cyc_wait += t_full - c0;
        }
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * kAccCols;
        const int64_t item0 = static_cast<int64_t>(t) * kTcTile;
#pragma unroll 1
        for (int chunk = half * kChunksPerWarp; chunk < (half + 1) * kChunksPerWarp; ++chunk) {
          if (P.debug_flags & 2) {                 // triage mode: never touch TMEM, just recycle the accumulator
            if (chunk == (half + 1) * kChunksPerWarp - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(acc_empty_leader_base + acc * 8);
                else mbar_arrive_local(acc_empty_leader_base + acc * 8);
              }
            }
            continue;
          }
          uint32_t r[32];
          const long long t_ld = epi_clock();
          tmem_ld32(taddr + chunk * 32, r);
          // this lane also "owns" one column of the chunk: fetch and unpack its metadata while the TMEM load is
          // in flight (only batches with per-row constraints pay for it)
          const int64_t col_item = item0 + chunk * 32 + lane;
          uint4 cm = make_uint4(0u, 0u, 0u, 0u);
          uint32_t class_word = 0xffffffffu;
          if (class_row != nullptr) class_word = __ldg(class_row + ((item0 + chunk * 32) >> 5));
          if (use_cons) {
            int2 mt = make_int2(0, 0);
            if (col_item < P.n_items) mt = __ldg(P.meta + col_item);
            const int c_theme = mt.y >> 16;
            cm.x = static_cast<unsigned>(mt.x);                                                     // num_parts
            cm.y = static_cast<unsigned>(static_cast<int>(static_cast<int16_t>(mt.y & 0xffff)));   // year
            cm.z = static_cast<unsigned>(c_theme >= 0 ? min(c_theme >> 5, 31) : 0) * 128u;          // theme word offset
            cm.w = (c_theme >= 0 && c_theme < 1024) ? (1u << (c_theme & 31)) : 0u;                  // theme bit (0 = NULL)
          }
          tmem_ld_wait();
          cyc_ld += epi_clock() - t_ld;
          if (chunk == (half + 1) * kChunksPerWarp - 1) cyc_hold += epi_clock() - t_full;
          if (chunk == (half + 1) * kChunksPerWarp - 1) {
            // the tile's last slice is in registers: hand the accumulator back to the MMA issuer BEFORE
            // processing it, so slow-path work and compactions overlap the next tile's MMAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(acc_empty_leader_base + acc * 8);
              else mbar_arrive_local(acc_empty_leader_base + acc * 8);
            }
          }
          // columns of this chunk that the row's exclusion list names (rare: a cursor compare per chunk)
          unsigned ex_mask = 0u;
          if constexpr (kWalk) {
            const int64_t c_lo = item0 + chunk * 32;
            while (static_cast<int64_t>(ex_next) < c_lo + 32) {
              if (static_cast<int64_t>(ex_next) >= c_lo) ex_mask |= 1u << static_cast<int>(ex_next - c_lo);
              const int32_t l = __ldg(P.q_excl + q);
              const int64_t lo = __ldg(P.excl_off + l), hi = __ldg(P.excl_off + l + 1);
              ++ex_i;
              ex_next = INT32_MAX;
              if (lo + ex_i < hi) {
                const int64_t r = static_cast<int64_t>(__ldg(P.excl_ids + lo + ex_i)) - P.id_offset;
                if (r < INT32_MAX) ex_next = static_cast<int>(r);
              }
            }
          }
          if (P.dump != nullptr) {
            if (q_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int64_t item = item0 + chunk * 32 + j;
                if (item < P.n_items) P.dump[static_cast<int64_t>(q) * P.n_items + item] = __uint_as_float(r[j]);
              }
            }
            continue;
          }
          if constexpr (kPilot) {
            // pilot pass: no candidates, only the row's best VALID score of every 8 columns (a group maximum is the score
            // of one distinct valid item, so the m-th largest of them bounds the row's m-th best from below)
            unsigned vm = class_word & ~ex_mask & ~__ballot_sync(kFull, col_item >= P.n_items);
            const int64_t c0 = item0 + chunk * 32;
            if (self_local >= c0 && self_local < c0 + 32) vm &= ~(1u << static_cast<int>(self_local - c0));
            if (dead) vm = 0u;
            float g[4];
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              float m = -INFINITY;
#pragma unroll
              for (int j = 0; j < 8; ++j) m = fmaxf(m, ((vm >> (gi * 8 + j)) & 1u) ? __uint_as_float(r[gi * 8 + j]) : -INFINITY);
              g[gi] = m;
            }
            if (q_valid)
              *reinterpret_cast<float4*>(P.pilot_max + static_cast<int64_t>(q) * P.pilot_groups + (c0 >> 3)) = make_float4(g[0], g[1], g[2], g[3]);
            continue;
          }
          // per row: the columns whose score passes the row's threshold and whose item the row's constraint set
          // admits (class mode: one prefetched bitmap word masks the 32 columns; a passing score of a masked-out
          // item must not drag the warp into the slow path -- two thirds of the triggers were of that kind)
          unsigned pm = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) pm |= (__uint_as_float(r[j]) >= thr) ? (1u << j) : 0u;
          pm &= class_word & ~ex_mask & ~__ballot_sync(kFull, col_item >= P.n_items);
          if (__any_sync(kFull, pm != 0u) && !(P.debug_flags & 1)) {
            // ---- slow path, entered only when some row of this warp has a live candidate
            ++stat_slow;
            const long long c0 = epi_clock();
            if (use_cons) {
              // column owners publish their item's unpacked metadata (prefetched above) in shared memory;
              // every row then checks only ITS OWN passing columns (trip count = max passes of one row)
              __syncwarp();
              colmeta_w[lane] = cm;
              __syncwarp();
              unsigned todo = pm, vm = 0;
              while (todo != 0u) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint4 c4 = colmeta_w[j];
                bool ok = (c4.x - static_cast<unsigned>(pmin) <= prange) && (c4.y - static_cast<unsigned>(ymin) <= yrange);
                if (has_theme) ok = ok && (theme_row[c4.z] & c4.w) != 0u;
                vm |= ok ? (1u << j) : 0u;
              }
              pm = vm;
              __syncwarp();
            }
            if (!__any_sync(kFull, pm != 0u)) ++stat_wasted;
            else {
              while (__any_sync(kFull, pm != 0u)) {
                if (pm != 0u) {
                  const int j = __ffs(pm) - 1;
                  pm &= pm - 1;
                  // r[j] with a run-time j: 5-level select tree (registers cannot be indexed dynamically)
                  uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                  for (int x = 0; x < 16; ++x) s16[x] = (j & 16) ? r[16 + x] : r[x];
#pragma unroll
                  for (int x = 0; x < 8; ++x) s8[x] = (j & 8) ? s16[8 + x] : s16[x];
#pragma unroll
                  for (int x = 0; x < 4; ++x) s4[x] = (j & 4) ? s8[4 + x] : s8[x];
#pragma unroll
                  for (int x = 0; x < 2; ++x) s2[x] = (j & 2) ? s4[2 + x] : s4[x];
                  const float sc = __uint_as_float((j & 1) ? s2[1] : s2[0]);
                  const int64_t item = item0 + chunk * 32 + j;
                  // cheap predicates were applied above; only rows with allow bitmaps / exclusion lists pay the call
                  if (item != self_local && (!needs_full || item_ok(P, q, item)) && (!excl_only || !query_excluded(P, q, item))) {
                    const int pos = atomicAdd(s_cnt + row_in_cta, 1);
                    __stcg(buf + pos, make_key(sc, static_cast<uint32_t>(item + P.id_offset)));
                    if (pos == kCompactLine) atomicOr(s_flag + tslot * 4 + ew, 1u << lane);   // (exactly one append crosses the line)
                    ++stat_appends;
                  }
                }
              }
            }
            cyc_slow += epi_clock() - c0;
          }
        }
        if (!kPilot && P.dump == nullptr && !(P.debug_flags & 3)) {
          // ---- tile end: every append of this tile by the row group's four warps is complete and visible
          // (the periodic mask is taken by ONE warp of the row group before the barrier -- it may miss this tile's last
          // appends of the other warps, which is harmless -- so that all four warps act on the same rows)
          if (half == 0) {
            unsigned pm_rows = 0u;
            if (((t - t0) & (kPublishTiles - 1)) == kPublishTiles - 1 && t + 8 < t1)
              pm_rows = __ballot_sync(kFull, !dead && s_cnt[row_in_cta] - s_sorted[row_in_cta] >= kPublishMinNew);
            if (lane == 0) s_pflag[tslot * 4 + ew] = pm_rows;
          }
          named_bar_sync(bar_id, kColWarps * 32);
          unsigned need = s_flag[tslot * 4 + ew];                // rows that crossed the line in this tile (written before the barrier)
          const int pslot = tslot == 0 ? 2 : tslot - 1;
          if (half == 0 && lane == 0) s_flag[pslot * 4 + ew] = 0u;   // recycle the flag word of tile t-1 (next written in tile t+2)
          // Periodic publication: a stream whose threshold was seeded appends a few dozen keys per unit and never
          // reaches the compaction line, so what it learnt used to stay private until the unit's end (~0.8 ms, a tenth
          // of a shard scan on 8 GPUs) while ~18 streams of the row ran elsewhere.  Every kPublishTiles tiles the rows
          // that gathered enough new keys are compacted, published and refined across streams (and ranks).
          const unsigned periodic = s_pflag[tslot * 4 + ew];
          need |= periodic;
          if (need != 0u) {
            const long long c0 = epi_clock();
            stat_compact += group_compact(P, need, (P.debug_flags & 16) ? 0u : periodic, half, s_cnt, s_sorted, s_thr, ew, buf, slot, q, lane);
            cyc_compact += epi_clock() - c0;
            named_bar_sync(bar_id, kColWarps * 32);              // new counts / thresholds are visible before the next append
          }
        }
        if (++tslot == 3) tslot = 0;
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
      if (!kPilot && P.dump == nullptr) {
        // unit end: leave a sorted prefix, publish it and refine the row's shared threshold, so that the units
        // that scan this row later start from everything this stream has learnt
        // (streams that learnt almost nothing in this unit skip it: their few keys still reach the final merge)
        const long long c0 = epi_clock();
        if (P.debug_flags & 3) named_bar_sync(bar_id, kColWarps * 32);
        const unsigned fresh = P.skip_unit_end ? 0u : __ballot_sync(kFull, !dead && s_cnt[row_in_cta] >= 8);
        stat_compact += group_compact(P, fresh, (P.debug_flags & 8) ? 0u : kFull, half, s_cnt, s_sorted, s_thr, ew, buf, slot, q, lane);
        named_bar_sync(bar_id, kColWarps * 32);
        cyc_end += epi_clock() - c0;
        if (half == 0) {
          const int c = s_cnt[row_in_cta];
          P.cand_cnt[slot] = c;
          if (P.stat_candidates != nullptr && c > 0) atomicAdd(P.stat_candidates, static_cast<unsigned long long>(c));
        }
        named_bar_sync(bar_id, kColWarps * 32);                  // the next unit re-initializes the row state
      }
    }
    if (warp == 4 && lane == 0) *s_poll = -1;       // releases the poller warp
    if (!kPilot && P.stat_candidates != nullptr && P.dump == nullptr) {
      const unsigned a = __reduce_add_sync(kFull, stat_appends);
      if (lane == 0) {
        atomicAdd(P.stat_candidates + 1, static_cast<unsigned long long>(a));
        atomicAdd(P.stat_candidates + 2, static_cast<unsigned long long>(stat_slow));
        atomicAdd(P.stat_candidates + 3, static_cast<unsigned long long>(stat_compact));
        atomicAdd(P.stat_candidates + 13, static_cast<unsigned long long>(stat_wasted));
        atomicAdd(P.stat_candidates + 14, static_cast<unsigned long long>(cyc_hold));
        atomicAdd(P.stat_candidates + 15, static_cast<unsigned long long>(cyc_ld));
        atomicAdd(P.stat_candidates + 8, static_cast<unsigned long long>(cyc_wait));
        atomicAdd(P.stat_candidates + 9, static_cast<unsigned long long>(cyc_slow));
        atomicAdd(P.stat_candidates + 10, static_cast<unsigned long long>(cyc_compact));
        atomicAdd(P.stat_candidates + 11, static_cast<unsigned long long>(cyc_end));
        atomicAdd(P.stat_candidates + 12, static_cast<unsigned long long>(epi_clock() - cyc_begin));
      }
    }
  }

  // ============================== teardown ==============================
  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, kTmemCols);
}

}  // namespace

// One variant of the kernel, launched as clusters of CG CTAs.  The variants are instantiated in separate translation units
// (bb_scan_tc_*.cu) so that they compile in parallel: one instantiation is ~70k SASS instructions.
template <int CG, bool kRowCons, bool kSplit, bool kPilot>
cudaError_t launch_scan_tc_variant(const ScanParams& P, const void* tmap_q, const void* tmap_e, const void* tmap_ql, const void* tmap_el,
                                   int n_groups, cudaStream_t stream) {
  const CUtensorMap* tq = static_cast<const CUtensorMap*>(tmap_q);
  const CUtensorMap* te = static_cast<const CUtensorMap*>(tmap_e);
  const CUtensorMap* tql = kSplit ? static_cast<const CUtensorMap*>(tmap_ql) : tq;
  const CUtensorMap* tel = kSplit ? static_cast<const CUtensorMap*>(tmap_el) : te;
  auto kernel = scan_tc_kernel<CG, kRowCons, kSplit, kPilot>;
  constexpr int smem = Cfg<CG, kRowCons>::kSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kThreads);
  cfg.stream = stream;
  cfg.gridDim = dim3(n_groups * CG);
  cfg.dynamicSmemBytes = smem;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, *tq, *te, *tql, *tel, P, n_groups);
}

}  // namespace bb
