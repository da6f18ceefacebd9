// C ABI of libbrickscore.so (include/brickscore.h): handle, resident catalog,
// per-stream workspaces, scheduling of the scan units, and the entry points.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/brickscore.h"
#include "bb_common.cuh"
#include "bb_kernels.h"

using namespace bb;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p != nullptr) {
      cudaError_t e = cudaFree(p);
      p = nullptr;
      cap = 0;
      if (e != cudaSuccess) return e;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct Workspace {
  std::mutex mu;
  DevBuf q_f32, q_bf16, q_lo, cand, cand_cnt, pub_cnt, thr_g, out_scores, out_ids, misc, stage, dump, shared_valid, pilot, row_flags;
  alignas(64) CUtensorMap tmap_ql;
  const void* tmap_ql_ptr = nullptr;
  int tmap_ql_rows = 0;
  // pinned mirrors: host inputs of a call are packed here and cross PCIe in ONE copy; host outputs come back in one
  unsigned char* h_stage = nullptr;
  size_t h_stage_cap = 0;
  unsigned char* h_out = nullptr;
  size_t h_out_cap = 0;
  cudaEvent_t flag_event = nullptr;   // deferred check of the device error flag after a device-output search
  bool flag_pending = false;
  cudaEvent_t stage_event = nullptr;  // the packed host->device copy out of h_stage
  bool stage_busy = false;
  // large pageable host inputs (a rating matrix, a catalog-sized batch) cross PCIe through two pinned chunk buffers:
  // the host memcpy of chunk i+1 overlaps the DMA of chunk i (a cudaMemcpyAsync from pageable memory ran at ~3 GB/s)
  unsigned char* h_chunk[2] = {nullptr, nullptr};
  cudaEvent_t chunk_event[2] = {nullptr, nullptr};
  bool chunk_busy[2] = {false, false};
  alignas(64) CUtensorMap tmap_q;
  const void* tmap_q_ptr = nullptr;
  int tmap_q_rows = 0;
  int32_t* host_flag = nullptr;   // pinned
  std::vector<bb_constraint> dedup_cons;   // host-side de-duplication of per-query descriptors (class mode)
  std::vector<int32_t> dedup_idx;
  // per-stream control block on the device (concurrent searches on other streams must not share these):
  // bytes [0,128) statistics, [128,132) wave-barrier counter, [192,196) error / watchdog flag
  unsigned char* ctl = nullptr;
  unsigned long long* stats() const { return reinterpret_cast<unsigned long long*>(ctl); }
  uint32_t* wave_counter() const { return reinterpret_cast<uint32_t*>(ctl + 128); }
  int32_t* err_flag() const { return reinterpret_cast<int32_t*>(ctl + 192); }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct bb_handle {
  bb_config cfg{};
  int k_pad = 0;
  int sm_count = 0;
  int cta_group = 2;
  bool loaded = false;
  float* cat_f32 = nullptr;
  __nv_bfloat16* cat_bf16 = nullptr;
  __nv_bfloat16* cat_lo = nullptr;        // second bf16 term of the catalog operand (exact-fp32 search on the tensor pipe)
  unsigned* cat_norm_max = nullptr;       // largest operand-row norm (float bits)
  alignas(64) CUtensorMap tmap_el;
  bool tmap_el_ok = false;
  std::atomic<long long> last_flagged{0};
  int2* meta = nullptr;
  int64_t* excl_off = nullptr;
  int32_t* excl_ids = nullptr;
  int64_t n_excl = 0;
  uint32_t* theme_sets = nullptr;
  int n_theme_sets = 0;
  int theme_words = 32;
  uint32_t* allow = nullptr;
  bool allow_owned = false;
  int n_allow = 0;
  int64_t allow_words = 0;
  alignas(64) CUtensorMap tmap_e;
  bool tmap_e_ok = false;
  EncodeTiledFn encode = nullptr;
  std::mutex mu;
  std::map<cudaStream_t, Workspace*> ws;
  std::atomic<long long> launches{0};
  std::atomic<long long> last_units{0};
  Workspace* last_ws = nullptr;   // most recently created / used workspace (statistics in bb_get_info)
  // exported communication block (row-sharded catalogs): [4 x rows_cap] threshold + level words | [2] arrival counters
  // (256 B) | inbox [2][8][slice_cap][k_cap] keys
  unsigned char* comm = nullptr;
  size_t comm_bytes = 0;
  int comm_rows = 0;                 // rows_cap (max_batch padded)
  int comm_slice_cap = 0, comm_k_cap = 0;
  size_t comm_arrivals_off = 0, comm_inbox_off = 0;
  unsigned char* peer_comm[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // by rank (own = comm)
  bool peer_ipc[8] = {false, false, false, false, false, false, false, false};
  int comm_rank = 0, comm_world = 1;
  unsigned long long comm_expected[2] = {0ull, 0ull};   // cumulative arrivals this rank waits for, per parity
  bool pending_scatter = false;      // bb_sharded_scatter issued, bb_sharded_collect outstanding
  int pending_slice = 0, pending_parity = 0, pending_mine = 0, pending_k = 0;
  bool pending_l2 = false;
  unsigned long long pending_expected = 0ull;   // arrivals the pending scatter's owner-side merge waits for
  unsigned long long* thr_shared() const { return reinterpret_cast<unsigned long long*>(comm); }
  std::atomic<uint32_t> epoch{0};          // private searches
  std::atomic<uint32_t> shared_epoch{0};   // collective searches (BB_Q_SHARE_THRESHOLDS): same sequence on all ranks
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_free;
  double scan_ms = 0.0;
  long long scan_launches = 0;
  mutable std::mutex err_mu;
  mutable std::string err;
};

namespace {

int fail(const bb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h != nullptr) {
    std::lock_guard<std::mutex> g(h->err_mu);
    h->err = buf;
  } else {
    g_create_error = buf;
  }
  return code;
}

#define BB_CUDA(h, expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(h, _e == cudaErrorMemoryAllocation ? BB_ERR_NOMEM : BB_ERR_CUDA, "%s: %s", #expr, \
                  cudaGetErrorString(_e));                                                    \
  } while (0)

bool is_device_ptr(const void* p) {
  if (p == nullptr) return false;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// Returns a device view of `p` (`bytes` long): p itself when it is device memory, else a staged copy inside the
// workspace's staging buffer.  Small host inputs are first packed into the workspace's PINNED mirror and cross PCIe in
// one cudaMemcpyAsync at flush() (a request with eight small arrays used to pay eight pageable copies); large ones
// (catalog loads) are copied directly.  Usage: account() every pointer, prepare(), get() in the same order, flush().
struct Stager {
  bb_handle* h;
  Workspace* w;
  cudaStream_t stream;
  size_t off = 0, need = 0;          // mirror-packed region [0, need_mirror); direct copies land behind it
  size_t off_direct = 0, need_mirror = 0;
  unsigned char kind[16] = {};       // 0 pageable host (mirror), 1 device, 2 pinned host and large (direct DMA)
  int n_acc = 0, n_get = 0;
  bool pinned = false, pending = false;
  static constexpr size_t kPinnedLimit = 8u << 20;
  static constexpr size_t kDirectMin = 64u << 10;     // a caller's pinned array of this size skips the mirror
  static unsigned char classify(const void* p, size_t bytes) {
    if (p == nullptr) return 0;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 1;
    return (a.type == cudaMemoryTypeHost && bytes >= kDirectMin) ? 2 : 0;
  }
  void account(const void* p, size_t bytes) {
    const unsigned char c = classify(p, bytes);
    if (n_acc < 16) kind[n_acc] = c;
    ++n_acc;
    if (p == nullptr || c == 1) return;
    need += align256(bytes);
    if (c == 0) need_mirror += align256(bytes);
  }
  cudaError_t prepare() {
    cudaError_t e = w->stage.ensure(need + 256);
    if (e != cudaSuccess) return e;
    off_direct = need_mirror;
    if (need_mirror == 0 || need_mirror > kPinnedLimit) return cudaSuccess;
    if (w->stage_event != nullptr && w->stage_busy) {        // the previous call's packed copy must have left the mirror
      e = cudaEventSynchronize(w->stage_event);
      w->stage_busy = false;
      if (e != cudaSuccess) return e;
    }
    if (w->h_stage_cap < need_mirror) {
      if (w->h_stage != nullptr) cudaFreeHost(w->h_stage);
      w->h_stage = nullptr;
      w->h_stage_cap = 0;
      const size_t want = need_mirror + need_mirror / 2 + 4096;
      if (cudaMallocHost(reinterpret_cast<void**>(&w->h_stage), want) != cudaSuccess) {
        cudaGetLastError();
        w->h_stage = nullptr;
        return cudaSuccess;                                  // fall back to direct copies
      }
      w->h_stage_cap = want;
    }
    pinned = true;
    return cudaSuccess;
  }
  cudaError_t get(const void* p, size_t bytes, const void** out) {
    const int i = n_get++;
    if (p == nullptr) { *out = nullptr; return cudaSuccess; }
    const unsigned char c = i < 16 ? kind[i] : classify(p, bytes);
    if (c == 1) { *out = p; return cudaSuccess; }
    if (c == 2 && i < 16) {
      // the caller's buffer is pinned already: one DMA straight from it (packing it into the mirror first cost a host
      // memcpy of the whole array -- 0.7 ms for a 7 MB query batch at config 2)
      void* dst = static_cast<unsigned char*>(w->stage.p) + off_direct;
      *out = dst;
      off_direct += align256(bytes);
      return cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, stream);
    }
    void* dst = static_cast<unsigned char*>(w->stage.p) + off;
    *out = dst;
    if (pinned) {
      std::memcpy(w->h_stage + off, p, bytes);
      pending = true;
      off += align256(bytes);
      return cudaSuccess;
    }
    off += align256(bytes);
    return big_h2d(dst, p, bytes);
  }
  static constexpr size_t kChunk = 8u << 20;
  // one core copies pageable memory at ~3-5 GB/s on the GPU hosts: eight cores keep the PCIe DMA of the previous chunk busy
  static void parallel_memcpy(unsigned char* dst, const unsigned char* src, size_t n) {
    constexpr int kCopyThreads = 8;
    const size_t part = (n / kCopyThreads + 4095) & ~static_cast<size_t>(4095);
    if (n < (1u << 20) || part == 0) { std::memcpy(dst, src, n); return; }
    std::thread helpers[kCopyThreads - 1];
    int started = 0;
    for (int t = 1; t < kCopyThreads; ++t) {
      const size_t o = part * t;
      if (o >= n) break;
      const size_t len = n - o < part ? n - o : part;
      try {
        helpers[started] = std::thread([=] { std::memcpy(dst + o, src + o, len); });
        ++started;
      } catch (...) {
        std::memcpy(dst + o, src + o, len);         // no thread to be had: copy inline
      }
    }
    std::memcpy(dst, src, part < n ? part : n);
    for (int t = 0; t < started; ++t) helpers[t].join();
  }
  cudaError_t big_h2d(void* dst, const void* p, size_t bytes) {
    cudaPointerAttributes a{};
    const bool registered = cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost;
    if (!registered) cudaGetLastError();
    if (registered || bytes <= kChunk) return cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, stream);
    for (int b = 0; b < 2; ++b) {
      if (w->h_chunk[b] == nullptr && cudaMallocHost(reinterpret_cast<void**>(&w->h_chunk[b]), kChunk) != cudaSuccess) {
        cudaGetLastError();
        w->h_chunk[b] = nullptr;
        return cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, stream);
      }
      if (w->chunk_event[b] == nullptr) {
        const cudaError_t e = cudaEventCreateWithFlags(&w->chunk_event[b], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
      }
    }
    int b = 0;
    for (size_t o = 0; o < bytes; o += kChunk, b ^= 1) {
      const size_t n = bytes - o < kChunk ? bytes - o : kChunk;
      if (w->chunk_busy[b]) {
        const cudaError_t e = cudaEventSynchronize(w->chunk_event[b]);
        w->chunk_busy[b] = false;
        if (e != cudaSuccess) return e;
      }
      parallel_memcpy(w->h_chunk[b], static_cast<const unsigned char*>(p) + o, n);
      cudaError_t e = cudaMemcpyAsync(static_cast<unsigned char*>(dst) + o, w->h_chunk[b], n, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) return e;
      e = cudaEventRecord(w->chunk_event[b], stream);
      if (e != cudaSuccess) return e;
      w->chunk_busy[b] = true;
    }
    return cudaSuccess;
  }
  cudaError_t flush() {
    if (!pending) return cudaSuccess;
    pending = false;
    cudaError_t e = cudaMemcpyAsync(w->stage.p, w->h_stage, off, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    if (w->stage_event == nullptr) {
      e = cudaEventCreateWithFlags(&w->stage_event, cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    w->stage_busy = true;
    return cudaEventRecord(w->stage_event, stream);
  }
};

Workspace* get_ws(bb_handle* h, cudaStream_t s) {
  std::lock_guard<std::mutex> g(h->mu);
  auto it = h->ws.find(s);
  if (it != h->ws.end()) return it->second;
  Workspace* w = new Workspace();
  if (cudaMallocHost(reinterpret_cast<void**>(&w->host_flag), 64) != cudaSuccess) {
    cudaGetLastError();
    w->host_flag = nullptr;
  }
  if (cudaMalloc(reinterpret_cast<void**>(&w->ctl), 1024) != cudaSuccess || cudaMemset(w->ctl, 0, 1024) != cudaSuccess) {
    cudaGetLastError();
    if (w->host_flag != nullptr) cudaFreeHost(w->host_flag);
    delete w;
    return nullptr;
  }
  h->last_ws = w;
  h->ws[s] = w;
  return w;
}

int make_tmap(bb_handle* h, CUtensorMap* out, const void* base, uint64_t rows, int k_pad, int box_rows) {
  if (h->encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return fail(h, BB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_pad), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_pad) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, BB_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  return BB_OK;
}

struct Plan {
  bool split;                                    // exact-fp32 search ran on the tensor pipe (split-bf16 operands)
  int n_qblocks, n_tiles, n_splits, n_groups, b_pad, tile_items, n_units;
  int pilot_tiles, pilot_splits, pilot_groups;   // threshold-seeding pilot pass over the leading tiles (0 = none)
};

Plan make_plan(const bb_handle* h, int n_queries, bool tc) {
  Plan p{};
  const int rows = tc ? kTcRowsPerCta * h->cta_group : kFp32QRows;
  p.tile_items = tc ? kTcTile : kFp32Tile;
  p.b_pad = (n_queries + 255) / 256 * 256;
  p.n_qblocks = (n_queries + rows - 1) / rows;
  // every padded row must belong to a block so its candidate count is written
  p.n_qblocks = p.b_pad / rows;
  p.n_tiles = static_cast<int>((h->cfg.n_items + p.tile_items - 1) / p.tile_items);
  if (p.n_tiles < 1) p.n_tiles = 1;
  const int hw_groups = tc ? h->sm_count / h->cta_group : h->sm_count * 2;
  // Catalog splits.  Every unit pays a fixed price (cold candidate streams, a sorted + published prefix at its end),
  // so units should be long (>= ~64k catalog rows) but there must be enough of them to fill the machine evenly: pick the
  // split count near that target whose static round-robin schedule keeps the most CTA groups busy.
  const long long mem_cap = (5ll << 29) / (static_cast<long long>(p.b_pad) * (tc ? kCapTc : kCap) * 8 * (tc ? kTcSubStreams : 1));   // 2.5 GiB
  long long s_cap = p.n_tiles;
  if (s_cap > 64) s_cap = 64;
  if (s_cap > mem_cap) s_cap = mem_cap > 0 ? mem_cap : 1;
  long long s_lo = p.n_tiles / (65536 / p.tile_items);      // ~64k catalog rows per unit
  const long long s_fill = (hw_groups + p.n_qblocks - 1) / p.n_qblocks;    // enough units to occupy every group once
  if (s_lo < s_fill) s_lo = s_fill;
  if (s_lo < 1) s_lo = 1;
  if (s_lo > s_cap) s_lo = s_cap;
  long long splits = s_lo;
  const long long s_len = p.n_tiles / (65536 / p.tile_items);
  if (tc && s_len < s_fill) {
    // Small catalog (fewer than ~64k rows per unit even with the fewest splits that fill the machine; 25k rows = 99
    // tiles): every unit pays a fixed price F of ~8 tile times (query operand load, pipeline fill, candidate-stream start
    // and end), so minimize rounds x (tiles per unit + F) -- ONE round of 64 long units beats two rounds of 144 short ones.
    double best = 1e300;
    const double fixed = 8.0;
    for (long long sc = (s_fill + 1) / 2 > 1 ? (s_fill + 1) / 2 : 1; sc <= s_cap && sc <= s_fill + 16; ++sc) {
      const long long units = sc * p.n_qblocks;
      const long long rounds = (units + hw_groups - 1) / hw_groups;
      const double cost = static_cast<double>(rounds) * (static_cast<double>((p.n_tiles + sc - 1) / sc) + fixed);
      if (cost < best - 1e-9) {
        best = cost;
        splits = sc;
      }
    }
  } else {
    // Large catalog: units of ~64k rows; among the split counts near that target pick the one whose static round-robin
    // schedule keeps the most CTA groups busy.
    double best = -1.0;
    const long long w_lo = (s_lo - 8 > s_fill) ? s_lo - 8 : (s_fill < s_lo ? s_fill : s_lo);
    for (long long sc = w_lo; sc <= s_cap && sc <= s_lo + 16; ++sc) {
      const long long units = sc * p.n_qblocks;
      const long long rounds = (units + hw_groups - 1) / hw_groups;
      // occupancy of the static schedule, with a mild preference for the target unit length
      const double eff = static_cast<double>(units) / static_cast<double>(rounds * hw_groups) -
                         0.001 * static_cast<double>(sc > s_lo ? sc - s_lo : s_lo - sc);
      if (eff > best + 1e-9) {
        best = eff;
        splits = sc;
      }
    }
  }
  if (const char* ov = getenv("BB_SPLITS")) {      // experimentation override
    const long long v = atoll(ov);
    if (v >= 1 && v <= s_cap) splits = v;
  }
  p.n_splits = static_cast<int>(splits);
  const long long units = static_cast<long long>(p.n_qblocks) * p.n_splits;
  p.n_groups = static_cast<int>(units < hw_groups ? units : hw_groups);
  p.n_units = static_cast<int>(units);
  // Pilot pass (threshold seeding): a short first launch over the leading tiles that only records, per query row, the
  // best valid score of every 8 columns; the (k' + E)-th largest of them seeds the row's threshold, so no stream of the
  // main scan starts cold (a cold stream appends ~every score of its first tiles and sorts its buffer three or four
  // times: on a 25k-row catalog that was 90% of the scan).  The pilot tiles are scanned twice; keep them <= 1/4 of the
  // catalog and <= 8192 rows (1024 group maxima per row).
  p.pilot_tiles = p.pilot_splits = p.pilot_groups = 0;
  if (tc && p.n_tiles >= 32 && h->cta_group == 2 && getenv("BB_NO_PILOT") == nullptr) {
    int pt = p.n_tiles / 4;
    if (pt > 8192 / p.tile_items) pt = 8192 / p.tile_items;
    int ps = hw_groups / p.n_qblocks;
    if (ps < 1) ps = 1;
    if (ps > pt) ps = pt;
    pt = pt / ps * ps;
    p.pilot_tiles = pt;
    p.pilot_splits = ps;
    const long long pu = static_cast<long long>(ps) * p.n_qblocks;
    p.pilot_groups = static_cast<int>(pu < hw_groups ? pu : hw_groups);
  }
  return p;
}

int device_guard(bb_handle* h) {
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) return fail(h, BB_ERR_CUDA, "cudaSetDevice(%d): %s", h->cfg.device, cudaGetErrorString(e));
  return BB_OK;
}

// A search with device outputs returns without synchronizing: its error flag (q_item outside the shard, watchdog)
// is copied to pinned memory behind the kernels and examined at the NEXT call on the same stream's workspace.
int check_deferred_flag(bb_handle* h, Workspace* w) {
  if (!w->flag_pending || w->host_flag == nullptr) return BB_OK;
  if (cudaEventQuery(w->flag_event) != cudaSuccess) {
    cudaGetLastError();
    return BB_OK;                      // still running: examined later
  }
  w->flag_pending = false;
  const int32_t flag = w->host_flag[1];
  if (flag == 1) return fail(h, BB_ERR_INVALID, "previous search on this stream: q_item / id outside this shard's id range (results invalid)");
  if (flag != 0) return fail(h, BB_ERR_CUDA, "previous search on this stream: device watchdog code %d", flag);
  return BB_OK;
}

int defer_flag_check(bb_handle* h, Workspace* w, cudaStream_t stream) {
  if (w->host_flag == nullptr) return BB_OK;
  if (w->flag_event == nullptr) BB_CUDA(h, cudaEventCreateWithFlags(&w->flag_event, cudaEventDisableTiming));
  BB_CUDA(h, cudaMemcpyAsync(w->host_flag + 1, w->err_flag(), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  BB_CUDA(h, cudaEventRecord(w->flag_event, stream));
  w->flag_pending = true;
  return BB_OK;
}

// Shared front half of bb_search / bb_search_sharded / bb_debug_scores: stage inputs, build the query operands, plan the
// units, (pilot pass + threshold seeds,) run the scan.  On return `P` describes the candidate state.
// `collective`: search of a row-sharded catalog whose ranks share thresholds through the communication block.
int run_scan(bb_handle* h, Workspace* w, const bb_query_batch* q, int k_sel, float* dump, cudaStream_t stream,
             ScanParams& P, Plan& plan, bool collective, const std::function<int()>* before_scan = nullptr) {
  const int B = q->n_queries;
  const int prec = q->precision;
  // Exact fp32 on the tensor pipe: split-bf16 scan + exact re-score + proof of completeness (rows that cannot be proven
  // fall back to the FFMA scan).  Needs the split catalog plane; small batches keep the HBM-streaming path, collective
  // searches and score dumps the FFMA tile scan.
  const bool split = prec == BB_PREC_FP32 && h->cat_lo != nullptr && h->cat_bf16 != nullptr && dump == nullptr && !collective && h->cta_group == 2 &&
                     B > kSmallBatch && h->k_pad / 64 <= kTcMaxKBlocks && getenv("BB_FP32_FFMA") == nullptr;
  const bool tc = prec != BB_PREC_FP32 || split;
  if (tc && h->cat_bf16 == nullptr) return fail(h, BB_ERR_STATE, "bf16 precision requested but BB_CFG_BUILD_BF16 not set");
  if ((!tc || prec == BB_PREC_BF16_RESCORE) && h->cat_f32 == nullptr)
    return fail(h, BB_ERR_STATE, "fp32 operands requested but BB_CFG_KEEP_FP32 not set");
  if (tc && h->k_pad / 64 > kTcMaxKBlocks)
    return fail(h, BB_ERR_UNSUPPORTED, "dim+cf_dim=%d exceeds the resident-operand limit of %d columns", h->k_pad,
                kTcMaxKBlocks * 64);
  if (q->q_emb == nullptr && q->q_item == nullptr && q->q_cf == nullptr)
    return fail(h, BB_ERR_INVALID, "query batch has neither embeddings, item ids nor CF factors");
  if ((q->q_emb != nullptr || q->q_item != nullptr) && h->cfg.dim == 0)
    return fail(h, BB_ERR_INVALID, "content query given but catalog has dim=0");
  if (q->q_cf != nullptr && h->cfg.cf_dim == 0) return fail(h, BB_ERR_INVALID, "CF query given but catalog has cf_dim=0");
  if (q->constraint_index == nullptr && q->n_constraints != 0 && q->n_constraints != 1 && q->n_constraints != B)
    return fail(h, BB_ERR_INVALID, "n_constraints must be 0, 1 or n_queries (or any count with constraint_index)");
  if (q->constraint_index != nullptr && q->n_constraints <= 0)
    return fail(h, BB_ERR_INVALID, "constraint_index given but n_constraints <= 0");
  if (q->constraint_index != nullptr && !is_device_ptr(q->constraint_index)) {
    for (int i = 0; i < B; ++i)
      if (q->constraint_index[i] < 0 || q->constraint_index[i] >= q->n_constraints)
        return fail(h, BB_ERR_INVALID, "constraint_index[%d] = %d out of range", i, q->constraint_index[i]);
  }
  if (q->n_constraints != 0 && q->constraints == nullptr) return fail(h, BB_ERR_INVALID, "constraints is NULL");
  if (q->n_allow < 0 || (q->n_allow > 0 && q->allow_bitmaps == nullptr)) return fail(h, BB_ERR_INVALID, "bad per-call allow bitmaps");
  if (q->n_excl_per_query < 0 || q->n_excl_per_query > 3 || (q->n_excl_per_query > 0 && q->excl_lists == nullptr))
    return fail(h, BB_ERR_INVALID, "n_excl_per_query must be in [0, 3] (with excl_lists)");
  const int n_allow = q->allow_bitmaps != nullptr ? q->n_allow : h->n_allow;
  const int n_qexcl = q->n_excl_per_query;
  if (n_qexcl > 0 && !is_device_ptr(q->excl_lists)) {
    for (int64_t i = 0; i < static_cast<int64_t>(B) * n_qexcl; ++i)
      if (q->excl_lists[i] >= h->n_excl) return fail(h, BB_ERR_INVALID, "excl_lists[%lld] = %d out of range", static_cast<long long>(i), q->excl_lists[i]);
  }

  plan = make_plan(h, B, tc);
  plan.split = split;
  // exact fp32 search of a small batch (the reference's one-query-per-call pattern): stream the catalog once at HBM
  // speed (bb_scan_small.cu) instead of running the 128-row FFMA tile scan on padded rows
  const bool small = !tc && dump == nullptr && B <= kSmallBatch && static_cast<size_t>(kSmallBatch) * h->k_pad * 4 <= 200 * 1024 &&
                     getenv("BB_NO_SMALL") == nullptr;
  if (small) {
    plan.n_splits = small_scan_splits(h->cfg.n_items);
    plan.n_units = plan.n_splits * B;
  }
  const int D = h->cfg.dim, F = h->cfg.cf_dim, KP = h->k_pad;

  // ---- validate descriptor table indices on the host when the descriptors are host memory
  if (q->n_constraints > 0 && !is_device_ptr(q->constraints)) {
    for (int i = 0; i < q->n_constraints; ++i) {
      const bb_constraint& c = q->constraints[i];
      if (c.theme_set >= h->n_theme_sets) return fail(h, BB_ERR_INVALID, "constraint %d: theme_set %d out of range", i, c.theme_set);
      if (c.allow_bitmap >= n_allow) return fail(h, BB_ERR_INVALID, "constraint %d: allow_bitmap %d out of range", i, c.allow_bitmap);
      for (int e = 0; e < 3; ++e)
        if (c.excl_list[e] >= h->n_excl) return fail(h, BB_ERR_INVALID, "constraint %d: excl_list %d out of range", i, c.excl_list[e]);
    }
  }

  // ---- per-query descriptors in host memory: queries that share a constraint set (popular filters) are folded
  //      into distinct sets + a per-query index, which selects the class-bitmap path of the scan
  const bb_constraint* q_cons = q->constraints;
  const int32_t* q_cidx = q->constraint_index;
  int n_cons = q->n_constraints;
  if (tc && dump == nullptr && q_cidx == nullptr && n_cons == B && B > 1 && !is_device_ptr(q_cons)) {
    std::unordered_map<std::string, int32_t> seen;
    w->dedup_cons.clear();
    w->dedup_idx.resize(B);
    bool few = true;
    for (int i = 0; i < B && few; ++i) {
      std::string key(reinterpret_cast<const char*>(q_cons + i), sizeof(bb_constraint));
      auto it = seen.find(key);
      if (it == seen.end()) {
        if (w->dedup_cons.size() >= 64) { few = false; break; }
        it = seen.emplace(std::move(key), static_cast<int32_t>(w->dedup_cons.size())).first;
        w->dedup_cons.push_back(q_cons[i]);
      }
      w->dedup_idx[i] = it->second;
    }
    if (few) {
      q_cons = w->dedup_cons.data();
      q_cidx = w->dedup_idx.data();
      n_cons = static_cast<int>(w->dedup_cons.size());
    }
  }

  // ---- stage host inputs (packed into one pinned buffer, one copy)
  const int64_t a_words = (h->cfg.n_items + 31) / 32;
  const size_t allow_bytes = q->allow_bitmaps != nullptr ? sizeof(uint32_t) * static_cast<size_t>(q->n_allow) * a_words : 0;
  Stager st{h, w, stream};
  st.account(q->q_emb, sizeof(float) * B * D);
  st.account(q->q_item, sizeof(int32_t) * B);
  st.account(q->q_cf, sizeof(float) * B * F);
  st.account(q->w_content, sizeof(float) * B);
  st.account(q->w_cf, sizeof(float) * B);
  st.account(q_cons, sizeof(bb_constraint) * n_cons);
  st.account(q->self_id, sizeof(int32_t) * B);
  st.account(q_cidx, sizeof(int32_t) * B);
  st.account(q->allow_bitmaps, allow_bytes);
  st.account(q->excl_lists, sizeof(int32_t) * B * n_qexcl);
  BB_CUDA(h, st.prepare());
  const void *d_emb, *d_item, *d_cf, *d_wc, *d_wf, *d_cons, *d_self, *d_cidx, *d_allow, *d_qexcl;
  BB_CUDA(h, st.get(q->q_emb, sizeof(float) * B * D, &d_emb));
  BB_CUDA(h, st.get(q->q_item, sizeof(int32_t) * B, &d_item));
  BB_CUDA(h, st.get(q->q_cf, sizeof(float) * B * F, &d_cf));
  BB_CUDA(h, st.get(q->w_content, sizeof(float) * B, &d_wc));
  BB_CUDA(h, st.get(q->w_cf, sizeof(float) * B, &d_wf));
  BB_CUDA(h, st.get(q_cons, sizeof(bb_constraint) * n_cons, &d_cons));
  BB_CUDA(h, st.get(q->self_id, sizeof(int32_t) * B, &d_self));
  BB_CUDA(h, st.get(q_cidx, sizeof(int32_t) * B, &d_cidx));
  BB_CUDA(h, st.get(q->allow_bitmaps, allow_bytes, &d_allow));
  BB_CUDA(h, st.get(q->excl_lists, sizeof(int32_t) * B * n_qexcl, &d_qexcl));
  BB_CUDA(h, st.flush());

  // ---- control block (statistics, wave counters, error flag) and query operands (padding rows are written as zeros
  //      by the kernel)
  BB_CUDA(h, cudaMemsetAsync(w->ctl, 0, 512, stream));
  const size_t qf_bytes = sizeof(float) * static_cast<size_t>(plan.b_pad) * KP;
  const size_t qb_bytes = sizeof(__nv_bfloat16) * static_cast<size_t>(plan.b_pad) * KP;
  BB_CUDA(h, w->q_f32.ensure(qf_bytes));
  BB_CUDA(h, w->q_bf16.ensure(qb_bytes));
  if (split) BB_CUDA(h, w->q_lo.ensure(qb_bytes));
  BB_CUDA(h, launch_prep_queries(static_cast<const float*>(d_emb), static_cast<const int32_t*>(d_item),
                                 static_cast<const float*>(d_cf), static_cast<const float*>(d_wc),
                                 static_cast<const float*>(d_wf), q->w_content_all, q->w_cf_all, B, plan.b_pad, D, F, KP,
                                 (q->flags & BB_Q_EMB_PRENORMALIZED) ? BB_METRIC_IP : h->cfg.metric, h->cat_f32, h->cat_bf16,
                                 h->cfg.id_offset, h->cfg.n_items, static_cast<float*>(w->q_f32.p), w->q_bf16.p,
                                 split ? w->q_lo.p : nullptr, w->err_flag(), stream));
  h->launches++;

  // ---- candidate state
  std::memset(&P, 0, sizeof(P));
  P.n_items = h->cfg.n_items;
  P.id_offset = h->cfg.id_offset;
  P.meta = h->meta;
  P.n_queries = B;
  P.b_pad = plan.b_pad;
  P.k_pad = KP;
  P.k_sel = k_sel;
  P.cons = static_cast<const bb_constraint*>(d_cons);
  P.cons_stride = (q_cidx == nullptr && n_cons == B && B > 1) ? 1 : 0;
  P.cons_index = static_cast<const int32_t*>(d_cidx);
  P.self_id = static_cast<const int32_t*>(d_self);
  P.q_excl = static_cast<const int32_t*>(d_qexcl);
  P.q_excl_n = n_qexcl;
  P.theme_sets = h->theme_sets;
  P.theme_words = h->theme_words;
  P.allow_bitmaps = q->allow_bitmaps != nullptr ? static_cast<const uint32_t*>(d_allow) : h->allow;
  P.allow_words = q->allow_bitmaps != nullptr ? a_words : h->allow_words;
  P.excl_off = h->excl_off;
  P.excl_ids = h->excl_ids;
  P.n_qblocks = plan.n_qblocks;
  P.n_splits = plan.n_splits;
  P.n_tiles = plan.n_tiles;
  P.n_sub = tc ? kTcSubStreams : 1;
  P.n_slot_splits = plan.n_splits;
  P.n_units = small ? plan.n_units : plan.n_qblocks * plan.n_splits;
  P.tile_items = plan.tile_items;
  P.dump = dump;
  P.stat_candidates = w->stats();
  P.err_flag = w->err_flag();
  h->last_ws = w;
  {
    const char* dbg = getenv("BB_DEBUG_FLAGS");
    P.debug_flags = dbg != nullptr ? atoi(dbg) : 0;
  }
  if (dump == nullptr) {
    size_t slots = static_cast<size_t>(P.n_slot_splits) * P.n_sub * plan.b_pad;
    P.cap = tc ? kCapTc : kCap;
    size_t cand_bytes = slots * static_cast<size_t>(P.cap) * sizeof(uint64_t), slots_max = slots;
    if (split) {                                   // the FFMA fallback pass re-uses the buffers with its own layout
      const Plan fp = make_plan(h, B, false);
      const size_t s2 = static_cast<size_t>(fp.n_splits) * fp.b_pad;
      if (s2 * kCap * sizeof(uint64_t) > cand_bytes) cand_bytes = s2 * kCap * sizeof(uint64_t);
      if (s2 > slots_max) slots_max = s2;
    }
    BB_CUDA(h, w->cand.ensure(cand_bytes));
    BB_CUDA(h, w->cand_cnt.ensure(slots_max * sizeof(int32_t)));
    if (!small) {
      BB_CUDA(h, w->pub_cnt.ensure(slots_max * sizeof(int32_t)));
      BB_CUDA(h, cudaMemsetAsync(w->pub_cnt.p, 0, slots * sizeof(int32_t), stream));
    }
    P.pub_cnt = small ? nullptr : static_cast<int32_t*>(w->pub_cnt.p);
    P.cand = static_cast<uint64_t*>(w->cand.p);
    P.cand_cnt = static_cast<int32_t*>(w->cand_cnt.p);
    // collective search of a row-sharded catalog: the threshold words live in the exported communication block and
    // carry the collective sequence number; every other search uses the workspace's private array
    if (collective) {
      P.epoch = ++h->shared_epoch;
      P.thr_g = h->thr_shared();
      P.lvl_stride = h->comm_rows;
      P.n_peers = 0;
      for (int r = 0; r < h->comm_world; ++r)
        if (r != h->comm_rank) P.peer_thr[P.n_peers++] = reinterpret_cast<const unsigned long long*>(h->peer_comm[r]);
    } else {
      const size_t before = w->thr_g.cap;
      BB_CUDA(h, w->thr_g.ensure(static_cast<size_t>(plan.b_pad) * sizeof(unsigned long long)));
      if (w->thr_g.cap != before) BB_CUDA(h, cudaMemsetAsync(w->thr_g.p, 0, w->thr_g.cap, stream));   // epoch 0 = never current
      P.epoch = ++h->epoch;
      P.thr_g = static_cast<unsigned long long*>(w->thr_g.p);
      P.n_peers = 0;
      P.lvl_stride = 0;
    }
    P.wave_counter = w->wave_counter();
    // A unit ends by sorting its candidate buffer, publishing the sorted prefix and refining the row's threshold across
    // streams -- for the units that scan the row LATER.  With a single scheduling round (or two, when the pilot pass
    // already seeded every threshold) there is nobody to learn from it: the merge kernel takes unsorted buffers.
    const long long rounds = (static_cast<long long>(P.n_units) + plan.n_groups - 1) / (plan.n_groups > 0 ? plan.n_groups : 1);
    const bool seeded = tc && plan.pilot_tiles > 0;
    P.skip_unit_end = (!collective && (rounds <= 1 || (seeded && rounds <= 2)) && getenv("BB_UNIT_END") == nullptr) ? 1 : 0;
  }

  // ---- few distinct constraint sets (one shared by the batch, or an explicit constraint_index): evaluate each
  //      set once per item with the mask kernel (K3); the scan epilogue then masks a 32-item chunk of a row with
  //      one word of its set's bitmap
  // (a single query with its own descriptor is a batch that shares one set: without the bitmap it took the per-row
  // variant, which has no pilot seeds -- 0.24 ms per call instead of 0.13 ms at config 1 with a valid-set filter)
  const int n_classes = n_cons == 1 ? 1 : (q_cidx != nullptr ? n_cons : 0);
  if (tc && dump == nullptr && n_classes >= 1 && n_classes <= 64) {
    const int64_t words = (h->cfg.n_items + 31) / 32;
    const size_t bm_bytes = static_cast<size_t>(n_classes) * words * sizeof(uint32_t);
    BB_CUDA(h, w->shared_valid.ensure(bm_bytes + 8 * static_cast<size_t>(n_classes) + 64));
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(
        (reinterpret_cast<uintptr_t>(static_cast<unsigned char*>(w->shared_valid.p) + bm_bytes) + 7) & ~static_cast<uintptr_t>(7));
    BB_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * n_classes, stream));
    BB_CUDA(h, launch_mask(P, n_classes, static_cast<uint32_t*>(w->shared_valid.p), cnt, stream));
    h->launches += (n_classes + 31) / 32;
    P.class_valid = static_cast<const uint32_t*>(w->shared_valid.p);
    P.class_words = words;
  }

  // ---- tensor maps
  if (tc) {
    if (w->tmap_q_ptr != w->q_bf16.p || w->tmap_q_rows != plan.b_pad) {
      int rc = make_tmap(h, &w->tmap_q, w->q_bf16.p, static_cast<uint64_t>(plan.b_pad), KP, kTcRowsPerCta);
      if (rc != BB_OK) return rc;
      w->tmap_q_ptr = w->q_bf16.p;
      w->tmap_q_rows = plan.b_pad;
    }
    if (!h->tmap_e_ok) {
      int rc = make_tmap(h, &h->tmap_e, h->cat_bf16, static_cast<uint64_t>(h->cfg.n_items), KP, kTcTile / h->cta_group);
      if (rc != BB_OK) return rc;
      h->tmap_e_ok = true;
    }
    if (split) {
      if (w->tmap_ql_ptr != w->q_lo.p || w->tmap_ql_rows != plan.b_pad) {
        int rc = make_tmap(h, &w->tmap_ql, w->q_lo.p, static_cast<uint64_t>(plan.b_pad), KP, kTcRowsPerCta);
        if (rc != BB_OK) return rc;
        w->tmap_ql_ptr = w->q_lo.p;
        w->tmap_ql_rows = plan.b_pad;
      }
      if (!h->tmap_el_ok) {
        std::lock_guard<std::mutex> g(h->mu);
        if (!h->tmap_el_ok) {
          int rc = make_tmap(h, &h->tmap_el, h->cat_lo, static_cast<uint64_t>(h->cfg.n_items), KP, kTcTile / h->cta_group);
          if (rc != BB_OK) return rc;
          h->tmap_el_ok = true;
        }
      }
    }
  }
  const void* t_ql = split ? static_cast<const void*>(&w->tmap_ql) : nullptr;
  const void* t_el = split ? static_cast<const void*>(&h->tmap_el) : nullptr;

  // ---- pilot pass + threshold seeds (tcgen05 scan; rows whose validity is a bitmap word: no constraints or class mode)
  const bool pilot = tc && dump == nullptr && plan.pilot_tiles > 0 && (P.cons == nullptr || P.class_valid != nullptr);
  if (pilot) {
    const int groups = plan.pilot_tiles * (kTcTile / 8);
    BB_CUDA(h, w->pilot.ensure(static_cast<size_t>(plan.b_pad) * groups * sizeof(float)));
    ScanParams PP = P;
    PP.pilot_mode = 1;
    PP.pilot_max = static_cast<float*>(w->pilot.p);
    PP.pilot_groups = groups;
    PP.n_tiles = plan.pilot_tiles;
    PP.n_splits = plan.pilot_splits;
    PP.n_units = plan.n_qblocks * plan.pilot_splits;
    PP.stat_candidates = nullptr;
    PP.wave_counter = w->wave_counter() + 1;
    PP.n_peers = 0;
    BB_CUDA(h, launch_scan_tc(PP, &w->tmap_q, &h->tmap_e, h->cta_group, plan.pilot_groups, stream, t_ql, t_el));
    const int64_t pilot_items = static_cast<int64_t>(plan.pilot_tiles) * kTcTile;
    BB_CUDA(h, launch_seed_thresholds(PP, pilot_items < h->cfg.n_items ? pilot_items : h->cfg.n_items, stream));
    h->launches += 2;
  }

  // ---- (pipelined collective search) work that must precede this scan but not its prologue: the owner-side merge of
  //      the PREVIOUS search waits for the peers' keys here, after this search's operands, masks and seeds are done
  if (before_scan != nullptr) {
    const int hrc = (*before_scan)();
    if (hrc != BB_OK) return hrc;
  }

  // ---- scan
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  if (h->profiling) {
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->prof_free.empty()) {
      ev = h->prof_free.back();
      h->prof_free.pop_back();
    } else {
      BB_CUDA(h, cudaEventCreate(&ev.first));
      BB_CUDA(h, cudaEventCreate(&ev.second));
    }
  }
  if (ev.first != nullptr) BB_CUDA(h, cudaEventRecord(ev.first, stream));
  if (tc) {
    BB_CUDA(h, launch_scan_tc(P, &w->tmap_q, &h->tmap_e, h->cta_group, plan.n_groups, stream, t_ql, t_el));
  } else if (small) {
    int bt = 1;
    while (bt < B) bt <<= 1;
    BB_CUDA(h, w->misc.ensure(static_cast<size_t>(bt) * static_cast<size_t>(h->cfg.n_items) * sizeof(float)));
    BB_CUDA(h, launch_scan_small(P, static_cast<const float*>(w->q_f32.p), h->cat_f32, static_cast<float*>(w->misc.p),
                                 h->sm_count, stream));
    h->launches++;
  } else {
    BB_CUDA(h, launch_scan_fp32(P, static_cast<const float*>(w->q_f32.p), h->cat_f32, stream));
  }
  if (ev.first != nullptr) {
    BB_CUDA(h, cudaEventRecord(ev.second, stream));
    std::lock_guard<std::mutex> g(h->mu);
    h->prof_pending.push_back(ev);
  }
  h->launches++;
  h->last_units = P.n_units;
  return BB_OK;
}

// Second half of the exact-fp32 search on the tensor pipe: the merge kernel flagged the rows whose top-k could not be
// proven complete (ties or near-ties at the boundary of the over-fetch); those rows -- usually none -- are scanned by the
// FFMA kernel (blocks without a flagged row leave immediately) and merged again over the same outputs.  No host
// synchronization: both launches are unconditional.
int run_ffma_rows(bb_handle* h, Workspace* w, const ScanParams& P1, int k, bool as_l2, float* d_scores, int32_t* d_ids,
                  cudaStream_t stream) {
  const Plan fp = make_plan(h, P1.n_queries, false);
  ScanParams P = P1;
  P.class_valid = nullptr;
  P.class_words = 0;
  P.k_sel = (k + 31) / 32 * 32;
  P.cap = kCap;
  P.n_sub = 1;
  P.n_qblocks = fp.n_qblocks;
  P.n_splits = fp.n_splits;
  P.n_tiles = fp.n_tiles;
  P.tile_items = fp.tile_items;
  P.n_slot_splits = fp.n_splits;
  P.n_units = fp.n_qblocks * fp.n_splits;
  P.b_pad = fp.b_pad;
  P.pilot_mode = 0;
  P.pilot_max = nullptr;
  P.skip_unit_end = 0;
  P.lvl_stride = 0;
  P.n_peers = 0;
  P.epoch = ++h->epoch;
  P.row_active = static_cast<const int32_t*>(w->row_flags.p);
  const size_t slots = static_cast<size_t>(P.n_slot_splits) * P.b_pad;
  BB_CUDA(h, cudaMemsetAsync(w->pub_cnt.p, 0, slots * sizeof(int32_t), stream));
  BB_CUDA(h, launch_scan_fp32(P, static_cast<const float*>(w->q_f32.p), h->cat_f32, stream));
  BB_CUDA(h, launch_merge_emit(P, k, 0, static_cast<const float*>(w->q_f32.p), h->cat_f32, as_l2, d_scores, d_ids, nullptr, stream));
  h->launches += 2;
  return BB_OK;
}

int check_device_flag(bb_handle* h, Workspace* w, cudaStream_t stream) {
  int32_t flag = 0;
  if (w->host_flag != nullptr) {
    BB_CUDA(h, cudaMemcpyAsync(w->host_flag, w->err_flag(), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    BB_CUDA(h, cudaStreamSynchronize(stream));
    flag = *w->host_flag;
  } else {
    BB_CUDA(h, cudaMemcpyAsync(&flag, w->err_flag(), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    BB_CUDA(h, cudaStreamSynchronize(stream));
  }
  if (flag == 1) return fail(h, BB_ERR_INVALID, "q_item / id outside this shard's id range");
  if (flag != 0) return fail(h, BB_ERR_CUDA, "device watchdog code %d", flag);
  return BB_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

const char* bb_version(void) { return "brickscore 0.1.0 (sm_100a)"; }

const char* bb_last_error(const bb_handle* h) {
  if (h == nullptr) return g_create_error.c_str();
  std::lock_guard<std::mutex> g(h->err_mu);
  thread_local std::string copy;
  copy = h->err;
  return copy.c_str();
}

int bb_create(bb_handle** out, const bb_config* cfg) {
  if (out == nullptr || cfg == nullptr) return fail(nullptr, BB_ERR_INVALID, "bb_create: NULL argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(bb_config)) return fail(nullptr, BB_ERR_INVALID, "bb_config.struct_size mismatch");
  if (cfg->n_items < 0 || cfg->dim < 0 || cfg->cf_dim < 0) return fail(nullptr, BB_ERR_INVALID, "negative shape");
  if (cfg->n_items > 0 && cfg->dim + cfg->cf_dim <= 0) return fail(nullptr, BB_ERR_INVALID, "dim + cf_dim must be > 0");
  if (cfg->k_max < 0 || cfg->k_max > kMaxSel) return fail(nullptr, BB_ERR_UNSUPPORTED, "k_max must be in [0, %d]", kMaxSel);
  if (cfg->n_items + cfg->id_offset > 0x7fffffffll) return fail(nullptr, BB_ERR_UNSUPPORTED, "ids must fit int32");
  if (cfg->theme_bits < 0 || cfg->theme_bits % 32 != 0) return fail(nullptr, BB_ERR_INVALID, "theme_bits must be a multiple of 32");
  if (cfg->metric != BB_METRIC_COSINE && cfg->metric != BB_METRIC_IP) return fail(nullptr, BB_ERR_INVALID, "unknown metric");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(nullptr, BB_ERR_CUDA, "no CUDA device (%s); brickscore has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, BB_ERR_INVALID, "device %d out of range", cfg->device);
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return fail(nullptr, BB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop{};
  e = cudaGetDeviceProperties(&prop, cfg->device);
  if (e != cudaSuccess) return fail(nullptr, BB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fail(nullptr, BB_ERR_UNSUPPORTED, "device is sm_%d%d; brickscore is built for sm_100a only", prop.major, prop.minor);

  bb_handle* h = new bb_handle();
  h->cfg = *cfg;
  if (h->cfg.flags == 0) h->cfg.flags = BB_CFG_KEEP_FP32 | BB_CFG_BUILD_BF16;
  if (h->cfg.theme_bits == 0) h->cfg.theme_bits = 1024;
  h->theme_words = h->cfg.theme_bits / 32;
  h->k_pad = (cfg->dim + cfg->cf_dim + 63) / 64 * 64;
  if (h->k_pad == 0) h->k_pad = 64;
  h->sm_count = prop.multiProcessorCount;
  const char* cg = getenv("BB_CTA_GROUP");
  h->cta_group = (cg != nullptr && atoi(cg) == 1) ? 1 : 2;
  auto bail = [&](int code) { bb_destroy(h); return code; };
  const size_t rows = static_cast<size_t>(cfg->n_items);
  if (rows > 0) {
    if (h->cfg.flags & BB_CFG_KEEP_FP32) {
      e = cudaMalloc(reinterpret_cast<void**>(&h->cat_f32), rows * h->k_pad * sizeof(float));
      if (e != cudaSuccess) { fail(nullptr, BB_ERR_NOMEM, "fp32 catalog alloc: %s", cudaGetErrorString(e)); return bail(BB_ERR_NOMEM); }
    }
    if (h->cfg.flags & BB_CFG_BUILD_BF16) {
      e = cudaMalloc(reinterpret_cast<void**>(&h->cat_bf16), rows * h->k_pad * sizeof(__nv_bfloat16));
      if (e != cudaSuccess) { fail(nullptr, BB_ERR_NOMEM, "bf16 catalog alloc: %s", cudaGetErrorString(e)); return bail(BB_ERR_NOMEM); }
    }
    // split plane: explicit flag, or by default when the catalog is small enough that +2 B per element is nothing
    const bool want_split = (h->cfg.flags & BB_CFG_BUILD_SPLIT) != 0 ||
                            (cfg->flags == 0 && rows * h->k_pad * sizeof(__nv_bfloat16) <= (static_cast<size_t>(4) << 30));
    if (want_split && (h->cfg.flags & BB_CFG_KEEP_FP32) && (h->cfg.flags & BB_CFG_BUILD_BF16)) {
      e = cudaMalloc(reinterpret_cast<void**>(&h->cat_lo), rows * h->k_pad * sizeof(__nv_bfloat16));
      if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->cat_norm_max), 256);
      if (e == cudaSuccess) e = cudaMemset(h->cat_norm_max, 0, 256);
      if (e != cudaSuccess) { fail(nullptr, BB_ERR_NOMEM, "split catalog alloc: %s", cudaGetErrorString(e)); return bail(BB_ERR_NOMEM); }
    }
    e = cudaMalloc(reinterpret_cast<void**>(&h->meta), rows * sizeof(int2));
    if (e != cudaSuccess) { fail(nullptr, BB_ERR_NOMEM, "metadata alloc: %s", cudaGetErrorString(e)); return bail(BB_ERR_NOMEM); }
  }
  if (cfg->max_batch > 0) {
    // exported communication block of a row-sharded catalog (see CommParams / ScanParams::lvl_stride)
    h->comm_rows = (cfg->max_batch + 255) / 256 * 256;
    h->comm_slice_cap = (h->comm_rows + 1) / 2;
    h->comm_k_cap = cfg->k_max > 0 ? cfg->k_max : kMaxSel;
    const size_t thr_bytes = align256(static_cast<size_t>(4) * h->comm_rows * sizeof(unsigned long long));
    h->comm_arrivals_off = thr_bytes;
    h->comm_inbox_off = thr_bytes + 256;
    h->comm_bytes = h->comm_inbox_off + static_cast<size_t>(2) * 8 * h->comm_slice_cap * h->comm_k_cap * sizeof(unsigned long long);
    e = cudaMalloc(reinterpret_cast<void**>(&h->comm), h->comm_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->comm, 0, h->comm_inbox_off);
    if (e != cudaSuccess) { fail(nullptr, BB_ERR_NOMEM, "communication block alloc: %s", cudaGetErrorString(e)); return bail(BB_ERR_NOMEM); }
    h->peer_comm[0] = h->comm;
  }
  *out = h;
  return BB_OK;
}

namespace {
void comm_close_peers(bb_handle* h) {
  for (int r = 0; r < 8; ++r) {
    if (h->peer_ipc[r] && h->peer_comm[r] != nullptr) cudaIpcCloseMemHandle(h->peer_comm[r]);
    h->peer_ipc[r] = false;
    h->peer_comm[r] = nullptr;
  }
  cudaGetLastError();
}
int comm_reset(bb_handle* h, int rank, int world) {
  // a (re-)import restarts the collective sequence on every rank
  h->comm_rank = rank;
  h->comm_world = world;
  h->shared_epoch = 0;
  h->comm_expected[0] = h->comm_expected[1] = 0ull;
  h->peer_comm[rank] = h->comm;
  BB_CUDA(h, cudaMemset(h->comm, 0, h->comm_inbox_off));
  return BB_OK;
}
}  // namespace

int bb_comm_export(bb_handle* h, void* out_handle) {
  if (h == nullptr || out_handle == nullptr) return BB_ERR_INVALID;
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (h->comm == nullptr) return fail(h, BB_ERR_STATE, "sharded search needs bb_config.max_batch > 0");
  static_assert(sizeof(cudaIpcMemHandle_t) == BB_IPC_HANDLE_BYTES, "ipc handle size");
  cudaIpcMemHandle_t hd;
  BB_CUDA(h, cudaIpcGetMemHandle(&hd, h->comm));
  std::memcpy(out_handle, &hd, sizeof(hd));
  return BB_OK;
}

int bb_comm_import(bb_handle* h, int32_t rank, int32_t world, const void* handles) {
  if (h == nullptr) return BB_ERR_INVALID;
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (world < 1 || world > 8 || rank < 0 || rank >= world || (world > 1 && handles == nullptr))
    return fail(h, BB_ERR_INVALID, "bb_comm_import: world must be in [1, 8] and rank in [0, world)");
  if (h->comm == nullptr) return fail(h, BB_ERR_STATE, "sharded search needs bb_config.max_batch > 0");
  BB_CUDA(h, cudaDeviceSynchronize());
  comm_close_peers(h);
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const unsigned char*>(handles) + static_cast<size_t>(r) * sizeof(hd), sizeof(hd));
    void* p = nullptr;
    BB_CUDA(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peer_comm[r] = static_cast<unsigned char*>(p);
    h->peer_ipc[r] = true;
  }
  return comm_reset(h, rank, world);
}

int bb_comm_link_local(bb_handle* h, int32_t rank, int32_t world, bb_handle* const* handles) {
  if (h == nullptr) return BB_ERR_INVALID;
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (world < 1 || world > 8 || rank < 0 || rank >= world || handles == nullptr)
    return fail(h, BB_ERR_INVALID, "bb_comm_link_local: world must be in [1, 8] and rank in [0, world)");
  if (h->comm == nullptr) return fail(h, BB_ERR_STATE, "sharded search needs bb_config.max_batch > 0");
  BB_CUDA(h, cudaDeviceSynchronize());
  comm_close_peers(h);
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    const bb_handle* o = handles[r];
    if (o == nullptr || o->comm == nullptr) return fail(h, BB_ERR_INVALID, "peer %d has no communication block", r);
    if (o->comm_rows != h->comm_rows || o->comm_k_cap != h->comm_k_cap)
      return fail(h, BB_ERR_INVALID, "peer %d was created with another max_batch / k_max", r);
    if (o->cfg.device != h->cfg.device) {
      int can = 0;
      BB_CUDA(h, cudaDeviceCanAccessPeer(&can, h->cfg.device, o->cfg.device));
      if (!can) return fail(h, BB_ERR_UNSUPPORTED, "device %d cannot access device %d", h->cfg.device, o->cfg.device);
      cudaError_t e = cudaDeviceEnablePeerAccess(o->cfg.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(h, BB_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
      cudaGetLastError();
    }
    h->peer_comm[r] = o->comm;
  }
  return comm_reset(h, rank, world);
}

void bb_destroy(bb_handle* h) {
  if (h == nullptr) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (auto& kv : h->ws) {
    Workspace* w = kv.second;
    for (DevBuf* b : {&w->q_f32, &w->q_bf16, &w->cand, &w->cand_cnt, &w->pub_cnt, &w->thr_g, &w->out_scores, &w->out_ids, &w->misc,
                      &w->stage, &w->dump, &w->shared_valid, &w->q_lo, &w->row_flags})
      b->release();
    w->pilot.release();
    if (w->host_flag != nullptr) cudaFreeHost(w->host_flag);
    if (w->h_stage != nullptr) cudaFreeHost(w->h_stage);
    if (w->h_out != nullptr) cudaFreeHost(w->h_out);
    for (int b = 0; b < 2; ++b) {
      if (w->h_chunk[b] != nullptr) cudaFreeHost(w->h_chunk[b]);
      if (w->chunk_event[b] != nullptr) cudaEventDestroy(w->chunk_event[b]);
    }
    if (w->flag_event != nullptr) cudaEventDestroy(w->flag_event);
    if (w->stage_event != nullptr) cudaEventDestroy(w->stage_event);
    if (w->ctl != nullptr) cudaFree(w->ctl);
    delete w;
  }
  comm_close_peers(h);
  if (h->comm) cudaFree(h->comm);
  if (h->cat_f32) cudaFree(h->cat_f32);
  if (h->cat_bf16) cudaFree(h->cat_bf16);
  if (h->cat_lo) cudaFree(h->cat_lo);
  if (h->cat_norm_max) cudaFree(h->cat_norm_max);
  if (h->meta) cudaFree(h->meta);
  if (h->excl_off) cudaFree(h->excl_off);
  if (h->excl_ids) cudaFree(h->excl_ids);
  if (h->theme_sets) cudaFree(h->theme_sets);
  if (h->allow && h->allow_owned) cudaFree(h->allow);
  for (auto* v : {&h->prof_pending, &h->prof_free})
    for (auto& ev : *v) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  cudaGetLastError();
  delete h;
}

int bb_get_info(const bb_handle* hc, bb_info* info) {
  bb_handle* h = const_cast<bb_handle*>(hc);
  if (h == nullptr || info == nullptr) return BB_ERR_INVALID;
  if (info->struct_size != sizeof(bb_info)) return fail(h, BB_ERR_INVALID, "bb_info.struct_size mismatch");
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  info->sm_count = h->sm_count;
  info->k_pad = h->k_pad;
  info->cta_group = h->cta_group;
  info->launches = h->launches.load();
  info->last_units = h->last_units.load();
  unsigned long long c[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (h->last_ws != nullptr) BB_CUDA(h, cudaMemcpy(c, h->last_ws->stats(), sizeof(c), cudaMemcpyDeviceToHost));
  info->last_candidates = static_cast<int64_t>(c[0]);
  info->last_appends = static_cast<int64_t>(c[1]);
  info->last_slow_chunks = static_cast<int64_t>(c[2]);
  info->last_compactions = static_cast<int64_t>(c[3]);
  info->mma_wait_epilogue_cycles = static_cast<int64_t>(c[4]);
  info->mma_wait_tma_cycles = static_cast<int64_t>(c[5]);
  info->mma_total_cycles = static_cast<int64_t>(c[6]);
  for (int i = 0; i < 5; ++i) info->epilogue_cycles[i] = static_cast<int64_t>(c[8 + i]);
  info->epilogue_cycles[5] = static_cast<int64_t>(c[14]);
  info->epilogue_cycles[6] = static_cast<int64_t>(c[15]);
  info->last_wasted_chunks = static_cast<int64_t>(c[13]);
  {
    int32_t nf = 0;
    if (h->last_ws != nullptr) BB_CUDA(h, cudaMemcpy(&nf, h->last_ws->ctl + 196, sizeof(nf), cudaMemcpyDeviceToHost));
    info->last_flagged_rows = nf;
  }
  {
    unsigned long long r[32] = {0};
    if (h->last_ws != nullptr) BB_CUDA(h, cudaMemcpy(r, h->last_ws->stats() + 32, sizeof(r), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; ++i) {
      info->round_cycles[i] = static_cast<int64_t>(r[i]);
      info->round_wait_cycles[i] = static_cast<int64_t>(r[16 + i]);
    }
  }
  const double rows = static_cast<double>(h->cfg.n_items);
  info->bytes_fp32 = h->cat_f32 ? rows * h->k_pad * 4.0 : 0.0;
  info->bytes_bf16 = h->cat_bf16 ? rows * h->k_pad * 2.0 : 0.0;
  {
    std::lock_guard<std::mutex> g(h->mu);
    for (auto& ev : h->prof_pending) {
      float ms = 0.f;
      if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
        h->scan_ms += ms;
        h->scan_launches++;
      }
      h->prof_free.push_back(ev);
    }
    h->prof_pending.clear();
    cudaGetLastError();
  }
  info->scan_ms = h->scan_ms;
  info->scan_launches = h->scan_launches;
  return BB_OK;
}

int bb_set_profiling(bb_handle* h, int enable) {
  if (h == nullptr) return BB_ERR_INVALID;
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  std::lock_guard<std::mutex> g(h->mu);
  for (auto& ev : h->prof_pending) {
    cudaEventSynchronize(ev.second);
    h->prof_free.push_back(ev);
  }
  h->prof_pending.clear();
  h->profiling = enable != 0;
  h->scan_ms = 0.0;
  h->scan_launches = 0;
  return BB_OK;
}

int bb_load_catalog(bb_handle* h, int64_t row_begin, int64_t n_rows, const float* emb, const float* item_factors,
                    const int32_t* num_parts, const int16_t* year, const int16_t* theme_id, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (row_begin < 0 || n_rows < 0 || row_begin + n_rows > h->cfg.n_items) return fail(h, BB_ERR_INVALID, "row range outside catalog");
  // factor-only reload (re-trained CF factors, recommendation_system.py:402-407): emb == NULL keeps the stored
  // embedding columns and metadata of rows that were loaded before
  const bool factors_only = h->cfg.dim > 0 && emb == nullptr && item_factors != nullptr && h->loaded;
  if (h->cfg.dim > 0 && emb == nullptr && !factors_only) return fail(h, BB_ERR_INVALID, "emb is NULL but dim=%d (only a factor reload of a loaded catalog may omit it)", h->cfg.dim);
  if (h->cfg.cf_dim > 0 && item_factors == nullptr) return fail(h, BB_ERR_INVALID, "item_factors is NULL but cf_dim=%d", h->cfg.cf_dim);
  const int D = h->cfg.dim, F = h->cfg.cf_dim, KP = h->k_pad;
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  const int64_t chunk = 1 << 18;   // staging granularity for host inputs
  for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
    const int64_t n = (n_rows - r0 < chunk) ? (n_rows - r0) : chunk;
    Stager st{h, w, stream};
    const float* e_p = emb ? emb + r0 * D : nullptr;
    const float* f_p = item_factors ? item_factors + r0 * F : nullptr;
    const int32_t* p_p = num_parts ? num_parts + r0 : nullptr;
    const int16_t* y_p = year ? year + r0 : nullptr;
    const int16_t* t_p = theme_id ? theme_id + r0 : nullptr;
    st.account(e_p, sizeof(float) * n * D);
    st.account(f_p, sizeof(float) * n * F);
    st.account(p_p, sizeof(int32_t) * n);
    st.account(y_p, sizeof(int16_t) * n);
    st.account(t_p, sizeof(int16_t) * n);
    if (st.need > 0) {
      if (w->stage.cap < st.need + 256) BB_CUDA(h, cudaStreamSynchronize(stream));   // previous chunk still reading it
      BB_CUDA(h, st.prepare());
    }
    const void *de, *df, *dp, *dy, *dt;
    BB_CUDA(h, st.get(e_p, sizeof(float) * n * D, &de));
    BB_CUDA(h, st.get(f_p, sizeof(float) * n * F, &df));
    BB_CUDA(h, st.get(p_p, sizeof(int32_t) * n, &dp));
    BB_CUDA(h, st.get(y_p, sizeof(int16_t) * n, &dy));
    BB_CUDA(h, st.get(t_p, sizeof(int16_t) * n, &dt));
    BB_CUDA(h, st.flush());
    const int64_t row = row_begin + r0;
    BB_CUDA(h, launch_prep_catalog(static_cast<const float*>(de), static_cast<const float*>(df), n, D, F, KP,
                                   h->cfg.metric, h->cat_f32 ? h->cat_f32 + row * KP : nullptr,
                                   h->cat_bf16 ? h->cat_bf16 + row * KP : nullptr, h->cat_lo ? h->cat_lo + row * KP : nullptr,
                                   h->cat_norm_max, stream));
    if (!factors_only || num_parts != nullptr || year != nullptr || theme_id != nullptr)
      BB_CUDA(h, launch_pack_meta(static_cast<const int32_t*>(dp), static_cast<const int16_t*>(dy),
                                  static_cast<const int16_t*>(dt), n, h->meta + row, stream));
    h->launches += 2;
    if (st.need > 0) BB_CUDA(h, cudaStreamSynchronize(stream));   // staging buffer is reused by the next chunk
  }
  h->loaded = true;
  return BB_OK;
}

int bb_fetch_embeddings(bb_handle* h, int32_t n, const int32_t* ids, float* out, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (!h->loaded) return fail(h, BB_ERR_STATE, "catalog not loaded");
  if (n <= 0 || ids == nullptr || out == nullptr) return fail(h, BB_ERR_INVALID, "bad arguments");
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  Stager st{h, w, stream};
  st.account(ids, sizeof(int32_t) * n);
  BB_CUDA(h, st.prepare());
  const void* d_ids;
  BB_CUDA(h, st.get(ids, sizeof(int32_t) * n, &d_ids));
  BB_CUDA(h, st.flush());
  const bool out_dev = is_device_ptr(out);
  float* d_out = out;
  if (!out_dev) {
    BB_CUDA(h, w->misc.ensure(sizeof(float) * static_cast<size_t>(n) * h->cfg.dim));
    d_out = static_cast<float*>(w->misc.p);
  }
  BB_CUDA(h, cudaMemsetAsync(w->err_flag(), 0, sizeof(int32_t), stream));
  BB_CUDA(h, launch_gather_rows(h->cat_f32, h->cat_bf16, h->k_pad, h->cfg.dim, h->cfg.id_offset, h->cfg.n_items,
                                static_cast<const int32_t*>(d_ids), n, d_out, w->err_flag(), stream));
  h->launches++;
  if (!out_dev) {
    BB_CUDA(h, cudaMemcpyAsync(out, d_out, sizeof(float) * static_cast<size_t>(n) * h->cfg.dim, cudaMemcpyDeviceToHost, stream));
    return check_device_flag(h, w, stream);
  }
  return BB_OK;
}

static int replace_table(bb_handle* h, void** slot, const void* src, size_t bytes) {
  if (*slot != nullptr) {
    BB_CUDA(h, cudaDeviceSynchronize());
    BB_CUDA(h, cudaFree(*slot));
    *slot = nullptr;
  }
  if (bytes == 0) return BB_OK;
  BB_CUDA(h, cudaMalloc(slot, bytes));
  BB_CUDA(h, cudaMemcpy(*slot, src, bytes, cudaMemcpyDefault));
  return BB_OK;
}

int bb_set_exclusion_lists(bb_handle* h, int64_t n_lists, const int64_t* offsets, const int32_t* item_ids) {
  if (h == nullptr) return BB_ERR_INVALID;
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (n_lists < 0 || (n_lists > 0 && offsets == nullptr)) return fail(h, BB_ERR_INVALID, "bad exclusion lists");
  int64_t total = 0;
  if (n_lists > 0) {
    if (is_device_ptr(offsets)) BB_CUDA(h, cudaMemcpy(&total, offsets + n_lists, sizeof(int64_t), cudaMemcpyDeviceToHost));
    else total = offsets[n_lists];
  }
  if (total > 0 && item_ids == nullptr) return fail(h, BB_ERR_INVALID, "item_ids is NULL");
  rc = replace_table(h, reinterpret_cast<void**>(&h->excl_off), offsets, n_lists > 0 ? sizeof(int64_t) * (n_lists + 1) : 0);
  if (rc != BB_OK) return rc;
  rc = replace_table(h, reinterpret_cast<void**>(&h->excl_ids), item_ids, sizeof(int32_t) * total);
  if (rc != BB_OK) return rc;
  h->n_excl = n_lists;
  return BB_OK;
}

int bb_set_theme_sets(bb_handle* h, int32_t n_sets, const uint32_t* bitmaps) {
  if (h == nullptr) return BB_ERR_INVALID;
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (n_sets < 0 || (n_sets > 0 && bitmaps == nullptr)) return fail(h, BB_ERR_INVALID, "bad theme sets");
  rc = replace_table(h, reinterpret_cast<void**>(&h->theme_sets), bitmaps, sizeof(uint32_t) * static_cast<size_t>(n_sets) * h->theme_words);
  if (rc != BB_OK) return rc;
  h->n_theme_sets = n_sets;
  return BB_OK;
}

int bb_set_allow_bitmaps(bb_handle* h, int32_t n_bitmaps, const uint32_t* bitmaps) {
  if (h == nullptr) return BB_ERR_INVALID;
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (n_bitmaps < 0 || (n_bitmaps > 0 && bitmaps == nullptr)) return fail(h, BB_ERR_INVALID, "bad allow bitmaps");
  const int64_t words = (h->cfg.n_items + 31) / 32;
  if (h->allow != nullptr && !h->allow_owned) h->allow = nullptr;
  rc = replace_table(h, reinterpret_cast<void**>(&h->allow), bitmaps, sizeof(uint32_t) * static_cast<size_t>(n_bitmaps) * words);
  if (rc != BB_OK) return rc;
  h->allow_owned = true;
  h->n_allow = n_bitmaps;
  h->allow_words = words;
  return BB_OK;
}

namespace {
int search_checks(bb_handle* h, const bb_query_batch* q, int32_t k, const void* out_scores, const void* out_ids) {
  if (q == nullptr || out_scores == nullptr || out_ids == nullptr) return fail(h, BB_ERR_INVALID, "NULL argument");
  if (q->struct_size != sizeof(bb_query_batch)) return fail(h, BB_ERR_INVALID, "bb_query_batch.struct_size mismatch");
  if (!h->loaded) return fail(h, BB_ERR_STATE, "catalog not loaded");
  if (q->n_queries <= 0) return fail(h, BB_ERR_INVALID, "n_queries must be > 0");
  if (h->cfg.max_batch > 0 && q->n_queries > h->cfg.max_batch) return fail(h, BB_ERR_INVALID, "n_queries %d > max_batch %d", q->n_queries, h->cfg.max_batch);
  if (k <= 0 || k > kMaxSel || (h->cfg.k_max > 0 && k > h->cfg.k_max)) return fail(h, BB_ERR_INVALID, "k=%d outside [1, %d]", k, h->cfg.k_max > 0 ? h->cfg.k_max : kMaxSel);
  if (q->precision < BB_PREC_FP32 || q->precision > BB_PREC_BF16_RESCORE) return fail(h, BB_ERR_INVALID, "unknown precision");
  return BB_OK;
}
// k' kept per candidate buffer: a multiple of 32; the re-scored path over-fetches.
int k_select(int k, bool rescore) {
  int k_sel = (k + 31) / 32 * 32;
  if (rescore) {
    k_sel = (k + k / 4 + 8 + 31) / 32 * 32;
    if (k_sel > kMaxSel) k_sel = kMaxSel;
  }
  return k_sel;
}
// candidates whose score is recomputed in exact fp32: the over-fetch itself (k + 25% + 8), not its rounding up to 32
int k_rescore(int k, int k_sel) {
  const int r = k + k / 4 + 8;
  return r < k_sel ? r : k_sel;
}
// Results to the caller: device outputs are written in place (nothing to do); host outputs come back in ONE packed
// copy ([scores | ids] through the workspace's pinned mirror) followed by the error-flag check.
struct OutPack {
  float* d_scores;
  int32_t* d_ids;
  bool host;
};
int out_prepare(bb_handle* h, Workspace* w, size_t n, float* out_scores, int32_t* out_ids, OutPack& o) {
  const bool s_dev = is_device_ptr(out_scores), i_dev = is_device_ptr(out_ids);
  if (s_dev != i_dev) return fail(h, BB_ERR_INVALID, "out_scores and out_ids must both be host or both be device memory");
  o.host = !s_dev;
  o.d_scores = out_scores;
  o.d_ids = out_ids;
  if (o.host) {
    BB_CUDA(h, w->out_scores.ensure(8 * n + 256));
    o.d_scores = static_cast<float*>(w->out_scores.p);
    o.d_ids = reinterpret_cast<int32_t*>(o.d_scores + n);
    if (w->h_out_cap < 8 * n) {
      if (w->h_out != nullptr) cudaFreeHost(w->h_out);
      w->h_out = nullptr;
      w->h_out_cap = 0;
      BB_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&w->h_out), 8 * n + 4096));
      w->h_out_cap = 8 * n + 4096;
    }
  }
  return BB_OK;
}
int out_finish(bb_handle* h, Workspace* w, size_t n, float* out_scores, int32_t* out_ids, const OutPack& o, cudaStream_t stream) {
  if (!o.host) return defer_flag_check(h, w, stream);
  if (4 * n >= Stager::kDirectMin && Stager::classify(out_scores, 4 * n) == 2 && Stager::classify(out_ids, 4 * n) == 2) {
    // the caller's result buffers are pinned: two DMAs straight into them instead of the mirror + two host memcpys
    BB_CUDA(h, cudaMemcpyAsync(out_scores, o.d_scores, 4 * n, cudaMemcpyDeviceToHost, stream));
    BB_CUDA(h, cudaMemcpyAsync(out_ids, o.d_ids, 4 * n, cudaMemcpyDeviceToHost, stream));
    return check_device_flag(h, w, stream);           // synchronizes the stream
  }
  BB_CUDA(h, cudaMemcpyAsync(w->h_out, o.d_scores, 8 * n, cudaMemcpyDeviceToHost, stream));
  int rc = check_device_flag(h, w, stream);           // synchronizes the stream
  if (rc != BB_OK) return rc;
  std::memcpy(out_scores, w->h_out, 4 * n);
  std::memcpy(out_ids, w->h_out + 4 * n, 4 * n);
  return BB_OK;
}
}  // namespace

int bb_search(bb_handle* h, const bb_query_batch* q, int32_t k, float* out_scores, int32_t* out_ids, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  rc = search_checks(h, q, k, out_scores, out_ids);
  if (rc != BB_OK) return rc;
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  rc = check_deferred_flag(h, w);
  if (rc != BB_OK) return rc;
  const bool rescore = q->precision == BB_PREC_BF16_RESCORE;
  const int b_pad = (q->n_queries + 255) / 256 * 256;
  // BB_Q_SHARE_THRESHOLDS: the ranks of a row-sharded catalog prune with each other's bounds; results stay local
  const bool collective = (q->flags & BB_Q_SHARE_THRESHOLDS) != 0 && h->comm_world > 1 && b_pad <= h->comm_rows;
  // (the exact-fp32 search on the tensor pipe over-fetches like the re-scored bf16 search; run_scan decides with the same test)
  const bool will_split = q->precision == BB_PREC_FP32 && h->cat_lo != nullptr && h->cat_bf16 != nullptr && !collective && h->cta_group == 2 &&
                          q->n_queries > kSmallBatch && h->k_pad / 64 <= kTcMaxKBlocks && getenv("BB_FP32_FFMA") == nullptr;
  const int k_sel = k_select(k, rescore || will_split);
  ScanParams P;
  Plan plan;
  rc = run_scan(h, w, q, k_sel, nullptr, stream, P, plan, collective);
  if (rc != BB_OK) return rc;
  const size_t n = static_cast<size_t>(q->n_queries) * k;
  OutPack o;
  rc = out_prepare(h, w, n, out_scores, out_ids, o);
  if (rc != BB_OK) return rc;
  const bool as_l2 = (q->flags & BB_Q_SCORES_AS_L2) != 0;
  if (plan.split) {
    // exact fp32 on the tensor pipe: re-score + proof; flagged rows fall back to the FFMA scan
    BB_CUDA(h, w->row_flags.ensure(sizeof(int32_t) * static_cast<size_t>(plan.b_pad)));
    int32_t* n_flagged = reinterpret_cast<int32_t*>(w->ctl + 196);
    BB_CUDA(h, launch_merge_emit(P, k, k_rescore(k, k_sel), static_cast<const float*>(w->q_f32.p), h->cat_f32, as_l2, o.d_scores, o.d_ids,
                                 nullptr, stream, static_cast<int32_t*>(w->row_flags.p), n_flagged, h->cat_norm_max));
    h->launches++;
    rc = run_ffma_rows(h, w, P, k, as_l2, o.d_scores, o.d_ids, stream);
    if (rc != BB_OK) return rc;
    return out_finish(h, w, n, out_scores, out_ids, o, stream);
  }
  BB_CUDA(h, launch_merge_emit(P, k, rescore ? k_rescore(k, k_sel) : 0, static_cast<const float*>(w->q_f32.p), h->cat_f32,
                               as_l2, o.d_scores, o.d_ids, nullptr, stream));
  h->launches++;
  return out_finish(h, w, n, out_scores, out_ids, o, stream);
}

namespace {
void comm_params(const bb_handle* h, int slice, int parity, CommParams& C) {
  C = CommParams{};
  for (int r = 0; r < h->comm_world; ++r) {
    C.inbox[r] = reinterpret_cast<unsigned long long*>(h->peer_comm[r] + h->comm_inbox_off);
    C.arrivals[r] = reinterpret_cast<unsigned long long*>(h->peer_comm[r] + h->comm_arrivals_off);
  }
  C.rank = h->comm_rank;
  C.world = h->comm_world;
  C.slice = slice;
  C.slice_cap = h->comm_slice_cap;
  C.k_cap = h->comm_k_cap;
  C.parity = parity;
}
}  // namespace

// First half of a collective search: scan this shard against the union thresholds, merge, and deliver this rank's k
// best keys of every query to the query's owner rank.  Enqueues only (never synchronizes).
namespace {
// launches the owner-side merge of the pending scatter into out_* (device or host memory)
int launch_pending_collect(bb_handle* h, Workspace* w, float* out_scores, int32_t* out_ids, cudaStream_t stream, bool device_only) {
  h->pending_scatter = false;
  const int n_mine = h->pending_mine, k = h->pending_k;
  const size_t n = static_cast<size_t>(n_mine > 0 ? n_mine : 1) * k;
  OutPack o;
  int rc = out_prepare(h, w, n, out_scores, out_ids, o);
  if (rc != BB_OK) return rc;
  if (device_only && o.host) return fail(h, BB_ERR_INVALID, "a pipelined collect needs device output buffers");
  CommParams C;
  comm_params(h, h->pending_slice, h->pending_parity, C);
  BB_CUDA(h, launch_merge_inbox(C, n_mine, k, h->pending_expected, h->pending_l2, o.d_scores, o.d_ids, w->err_flag(), stream));
  h->launches++;
  if (device_only) return BB_OK;
  if (n_mine == 0) return o.host ? check_device_flag(h, w, stream) : defer_flag_check(h, w, stream);
  return out_finish(h, w, static_cast<size_t>(n_mine) * k, out_scores, out_ids, o, stream);
}

int sharded_scatter_impl(bb_handle* h, const bb_query_batch* q, int32_t k, float* prev_scores, int32_t* prev_ids, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  rc = search_checks(h, q, k, q, q);
  if (rc != BB_OK) return rc;
  if (h->comm == nullptr || h->comm_world <= 1) return fail(h, BB_ERR_STATE, "sharded search needs bb_config.max_batch > 0 and bb_comm_import / bb_comm_link_local with world > 1");
  const bool chain = prev_scores != nullptr || prev_ids != nullptr;
  if (chain && (prev_scores == nullptr || prev_ids == nullptr || !h->pending_scatter))
    return fail(h, BB_ERR_STATE, "bb_sharded_scatter_after needs both output buffers of a scatter that was not collected yet");
  if (!chain && h->pending_scatter) return fail(h, BB_ERR_STATE, "bb_sharded_scatter: the previous scatter was not collected");
  const int B = q->n_queries, world = h->comm_world;
  const int b_pad = (B + 255) / 256 * 256;
  const int slice = (B + world - 1) / world;
  if (b_pad > h->comm_rows || slice > h->comm_slice_cap || k > h->comm_k_cap)
    return fail(h, BB_ERR_INVALID, "batch %d / k %d exceed the communication block (max_batch %d, k_max %d)", B, k, h->cfg.max_batch, h->comm_k_cap);
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  rc = check_deferred_flag(h, w);
  if (rc != BB_OK) return rc;
  const bool rescore = q->precision == BB_PREC_BF16_RESCORE;
  ScanParams P;
  Plan plan;
  const int k_sel = k_select(k, rescore);
  // pipelined: the previous search's owner-side merge (which waits for the slowest peer's keys) is launched after THIS
  // search's operand preparation, masks, pilot and seeds -- work that does not depend on the peers -- and before its scan;
  // the two inbox halves stay safe because the merge still precedes this search's scan and scatter in stream order
  const std::function<int()> collect_prev = [&]() { return launch_pending_collect(h, w, prev_scores, prev_ids, stream, true); };
  const int prev_parity = h->pending_parity;
  rc = run_scan(h, w, q, k_sel, nullptr, stream, P, plan, true, chain ? &collect_prev : nullptr);
  if (rc != BB_OK) return rc;
  // chained searches must alternate between the two inbox halves: another collective search on this handle between them
  // (an NCCL-exchange search that shares thresholds also advances the collective epoch) would make them collide
  if (chain && static_cast<int>(P.epoch & 1u) == prev_parity)
    return fail(h, BB_ERR_STATE, "bb_sharded_scatter_after: another collective search ran between the chained searches");
  CommParams C;
  comm_params(h, slice, static_cast<int>(P.epoch & 1u), C);
  BB_CUDA(h, launch_merge_emit(P, k, rescore ? k_rescore(k, k_sel) : 0, static_cast<const float*>(w->q_f32.p), h->cat_f32, false, nullptr, nullptr, &C, stream));
  h->launches++;
  int n_mine = B - h->comm_rank * slice;
  if (n_mine > slice) n_mine = slice;
  if (n_mine < 0) n_mine = 0;
  h->comm_expected[C.parity] += static_cast<unsigned long long>(n_mine) * world;
  h->pending_scatter = true;
  h->pending_slice = slice;
  h->pending_parity = C.parity;
  h->pending_mine = n_mine;
  h->pending_k = k;
  h->pending_l2 = (q->flags & BB_Q_SCORES_AS_L2) != 0;
  h->pending_expected = h->comm_expected[C.parity];
  return BB_OK;
}
}  // namespace

int bb_sharded_scatter(bb_handle* h, const bb_query_batch* q, int32_t k, void* stream_) {
  return sharded_scatter_impl(h, q, k, nullptr, nullptr, stream_);
}
int bb_sharded_scatter_after(bb_handle* h, const bb_query_batch* q, int32_t k, float* prev_out_scores, int32_t* prev_out_ids,
                             void* stream_) {
  if (prev_out_scores == nullptr || prev_out_ids == nullptr) return h != nullptr ? fail(h, BB_ERR_INVALID, "NULL argument") : BB_ERR_INVALID;
  return sharded_scatter_impl(h, q, k, prev_out_scores, prev_out_ids, stream_);
}

// Second half: wait (on the device, bounded) until every rank's keys of this rank's query slice have arrived, merge
// the `world` lists per query and emit [n_mine x k].  out_* host or device memory.
int bb_sharded_collect(bb_handle* h, float* out_scores, int32_t* out_ids, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (!h->pending_scatter) return fail(h, BB_ERR_STATE, "bb_sharded_collect without bb_sharded_scatter");
  if (out_scores == nullptr || out_ids == nullptr) return fail(h, BB_ERR_INVALID, "NULL argument");
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  return launch_pending_collect(h, w, out_scores, out_ids, stream, false);
}

int bb_search_sharded(bb_handle* h, const bb_query_batch* q, int32_t k, float* out_scores, int32_t* out_ids, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  if (h->comm_world <= 1) return bb_search(h, q, k, out_scores, out_ids, stream_);   // one rank owns every row and every query
  if (out_scores == nullptr || out_ids == nullptr) return fail(h, BB_ERR_INVALID, "NULL argument");
  int rc = bb_sharded_scatter(h, q, k, stream_);
  if (rc != BB_OK) return rc;
  return bb_sharded_collect(h, out_scores, out_ids, stream_);
}

int bb_debug_scores(bb_handle* h, const bb_query_batch* q, float* out_scores, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (q == nullptr || out_scores == nullptr) return fail(h, BB_ERR_INVALID, "NULL argument");
  if (q->struct_size != sizeof(bb_query_batch)) return fail(h, BB_ERR_INVALID, "bb_query_batch.struct_size mismatch");
  if (!h->loaded) return fail(h, BB_ERR_STATE, "catalog not loaded");
  if (q->n_queries <= 0) return fail(h, BB_ERR_INVALID, "n_queries must be > 0");
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  const size_t bytes = sizeof(float) * static_cast<size_t>(q->n_queries) * h->cfg.n_items;
  const bool dev = is_device_ptr(out_scores);
  float* d = out_scores;
  if (!dev) {
    BB_CUDA(h, w->dump.ensure(bytes));
    d = static_cast<float*>(w->dump.p);
  }
  ScanParams P;
  Plan plan;
  rc = run_scan(h, w, q, 32, d, stream, P, plan, false);
  if (rc != BB_OK) return rc;
  if (!dev) BB_CUDA(h, cudaMemcpyAsync(out_scores, d, bytes, cudaMemcpyDeviceToHost, stream));
  return check_device_flag(h, w, stream);
}

int bb_apply_constraints(bb_handle* h, const bb_constraint* constraints, int32_t n, uint32_t* out_bitmaps,
                         int64_t* out_counts, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (!h->loaded) return fail(h, BB_ERR_STATE, "catalog not loaded");
  if (n <= 0 || constraints == nullptr || out_bitmaps == nullptr) return fail(h, BB_ERR_INVALID, "bad arguments");
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  const int64_t words = (h->cfg.n_items + 31) / 32;
  Stager st{h, w, stream};
  st.account(constraints, sizeof(bb_constraint) * n);
  BB_CUDA(h, st.prepare());
  const void* d_cons;
  BB_CUDA(h, st.get(constraints, sizeof(bb_constraint) * n, &d_cons));
  BB_CUDA(h, st.flush());
  const bool b_dev = is_device_ptr(out_bitmaps);
  const size_t bm_bytes = sizeof(uint32_t) * static_cast<size_t>(n) * words;
  BB_CUDA(h, w->misc.ensure((b_dev ? 0 : align256(bm_bytes)) + sizeof(unsigned long long) * n + 256));
  uint32_t* d_bm = b_dev ? out_bitmaps : static_cast<uint32_t*>(w->misc.p);
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(w->misc.p) + (b_dev ? 0 : align256(bm_bytes)));
  BB_CUDA(h, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * n, stream));
  ScanParams P;
  std::memset(&P, 0, sizeof(P));
  P.n_items = h->cfg.n_items;
  P.id_offset = h->cfg.id_offset;
  P.meta = h->meta;
  P.cons = static_cast<const bb_constraint*>(d_cons);
  P.theme_sets = h->theme_sets;
  P.theme_words = h->theme_words;
  P.allow_bitmaps = h->allow;
  P.allow_words = h->allow_words;
  P.excl_off = h->excl_off;
  P.excl_ids = h->excl_ids;
  BB_CUDA(h, launch_mask(P, n, d_bm, d_cnt, stream));
  h->launches++;
  if (!b_dev) BB_CUDA(h, cudaMemcpyAsync(out_bitmaps, d_bm, bm_bytes, cudaMemcpyDeviceToHost, stream));
  if (out_counts != nullptr)
    BB_CUDA(h, cudaMemcpyAsync(out_counts, d_cnt, sizeof(int64_t) * n, is_device_ptr(out_counts) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
  if (!b_dev || (out_counts != nullptr && !is_device_ptr(out_counts))) BB_CUDA(h, cudaStreamSynchronize(stream));
  return BB_OK;
}

int bb_hybrid_combine(bb_handle* h, int32_t n_rows, int32_t len_c, const int32_t* c_ids, const float* c_scores,
                      int32_t len_f, const int32_t* f_ids, const float* f_scores, const double* weights,
                      int32_t weights_per_row, int32_t k, double* out_scores, int32_t* out_ids, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (n_rows <= 0 || k <= 0 || len_c < 0 || len_f < 0 || len_c + len_f > 512 || len_c + len_f == 0)
    return fail(h, BB_ERR_INVALID, "hybrid_combine: bad shape (len_c + len_f must be in [1, 512])");
  if (weights == nullptr || out_scores == nullptr || out_ids == nullptr) return fail(h, BB_ERR_INVALID, "NULL argument");
  if ((len_c > 0 && (c_ids == nullptr || c_scores == nullptr)) || (len_f > 0 && (f_ids == nullptr || f_scores == nullptr)))
    return fail(h, BB_ERR_INVALID, "NULL list");
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  Stager st{h, w, stream};
  const size_t nc = static_cast<size_t>(n_rows) * len_c, nf = static_cast<size_t>(n_rows) * len_f;
  const size_t nw = weights_per_row ? static_cast<size_t>(n_rows) * 2 : 2;
  st.account(c_ids, 4 * nc); st.account(c_scores, 4 * nc); st.account(f_ids, 4 * nf); st.account(f_scores, 4 * nf);
  st.account(weights, 8 * nw);
  BB_CUDA(h, st.prepare());
  const void *dci, *dcs, *dfi, *dfs, *dw;
  BB_CUDA(h, st.get(c_ids, 4 * nc, &dci)); BB_CUDA(h, st.get(c_scores, 4 * nc, &dcs));
  BB_CUDA(h, st.get(f_ids, 4 * nf, &dfi)); BB_CUDA(h, st.get(f_scores, 4 * nf, &dfs));
  BB_CUDA(h, st.get(weights, 8 * nw, &dw));
  BB_CUDA(h, st.flush());
  const bool s_dev = is_device_ptr(out_scores), i_dev = is_device_ptr(out_ids);
  const size_t so = align256(sizeof(double) * static_cast<size_t>(n_rows) * k);
  BB_CUDA(h, w->misc.ensure(so + sizeof(int32_t) * static_cast<size_t>(n_rows) * k + 256));
  double* d_s = s_dev ? out_scores : static_cast<double*>(w->misc.p);
  int32_t* d_i = i_dev ? out_ids : reinterpret_cast<int32_t*>(static_cast<unsigned char*>(w->misc.p) + so);
  BB_CUDA(h, launch_hybrid_combine(n_rows, len_c, static_cast<const int32_t*>(dci), static_cast<const float*>(dcs), len_f,
                                   static_cast<const int32_t*>(dfi), static_cast<const float*>(dfs),
                                   static_cast<const double*>(dw), weights_per_row ? 1 : 0, k, d_s, d_i, stream));
  h->launches++;
  if (!s_dev) BB_CUDA(h, cudaMemcpyAsync(out_scores, d_s, sizeof(double) * static_cast<size_t>(n_rows) * k, cudaMemcpyDeviceToHost, stream));
  if (!i_dev) BB_CUDA(h, cudaMemcpyAsync(out_ids, d_i, sizeof(int32_t) * static_cast<size_t>(n_rows) * k, cudaMemcpyDeviceToHost, stream));
  if (!s_dev || !i_dev) BB_CUDA(h, cudaStreamSynchronize(stream));
  return BB_OK;
}

int bb_merge_topk(bb_handle* h, int32_t n_queries, int32_t n_lists, int32_t len, const float* scores,
                  const int32_t* ids, int32_t k, float* out_scores, int32_t* out_ids, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (n_queries <= 0 || n_lists <= 0 || len <= 0 || k <= 0 || k > kMaxSel) return fail(h, BB_ERR_INVALID, "merge_topk: bad shape");
  if (scores == nullptr || ids == nullptr || out_scores == nullptr || out_ids == nullptr) return fail(h, BB_ERR_INVALID, "NULL argument");
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  const size_t n_in = static_cast<size_t>(n_lists) * n_queries * len;
  Stager st{h, w, stream};
  st.account(scores, 4 * n_in); st.account(ids, 4 * n_in);
  BB_CUDA(h, st.prepare());
  const void *ds, *di;
  BB_CUDA(h, st.get(scores, 4 * n_in, &ds)); BB_CUDA(h, st.get(ids, 4 * n_in, &di));
  BB_CUDA(h, st.flush());
  const bool s_dev = is_device_ptr(out_scores), i_dev = is_device_ptr(out_ids);
  const size_t n_out = static_cast<size_t>(n_queries) * k;
  if (!s_dev) BB_CUDA(h, w->out_scores.ensure(4 * n_out));
  if (!i_dev) BB_CUDA(h, w->out_ids.ensure(4 * n_out));
  float* d_s = s_dev ? out_scores : static_cast<float*>(w->out_scores.p);
  int32_t* d_i = i_dev ? out_ids : static_cast<int32_t*>(w->out_ids.p);
  BB_CUDA(h, launch_merge_lists(n_queries, n_lists, len, static_cast<const float*>(ds), static_cast<const int32_t*>(di), k, d_s, d_i, stream));
  h->launches++;
  if (!s_dev) BB_CUDA(h, cudaMemcpyAsync(out_scores, d_s, 4 * n_out, cudaMemcpyDeviceToHost, stream));
  if (!i_dev) BB_CUDA(h, cudaMemcpyAsync(out_ids, d_i, 4 * n_out, cudaMemcpyDeviceToHost, stream));
  if (!s_dev || !i_dev) BB_CUDA(h, cudaStreamSynchronize(stream));
  return BB_OK;
}

int bb_csr_spmm_f64(bb_handle* h, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                    const float* values, const double* x, int32_t l, double* y, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (n_rows <= 0 || n_cols <= 0 || l <= 0 || l > 128) return fail(h, BB_ERR_INVALID, "csr_spmm: bad shape (L must be in [1, 128])");
  if (indptr == nullptr || indices == nullptr || values == nullptr || x == nullptr || y == nullptr) return fail(h, BB_ERR_INVALID, "NULL argument");
  if (!is_device_ptr(indptr) || !is_device_ptr(indices) || !is_device_ptr(values) || !is_device_ptr(x) || !is_device_ptr(y))
    return fail(h, BB_ERR_INVALID, "csr_spmm: all arrays must be device pointers");
  if (nnz <= 0) return fail(h, BB_ERR_INVALID, "csr_spmm: nnz must be positive");
  BB_CUDA(h, launch_csr_spmm_f64(n_rows, nnz, indptr, indices, values, x, l, y, h->sm_count, stream));
  h->launches += 2;
  return BB_OK;
}

int bb_item_item_topk(bb_handle* h, int64_t n_users, int32_t n_items, const int64_t* indptr, const int32_t* indices,
                      const float* values, int32_t k, float* out_scores, int32_t* out_ids, void* stream_) {
  if (h == nullptr) return BB_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = device_guard(h);
  if (rc != BB_OK) return rc;
  if (n_users <= 0 || n_items <= 1 || k <= 0 || k > kMaxSel || k >= n_items) return fail(h, BB_ERR_INVALID, "item_item: bad shape");
  if (indptr == nullptr || indices == nullptr || values == nullptr || out_scores == nullptr || out_ids == nullptr)
    return fail(h, BB_ERR_INVALID, "NULL argument");
  Workspace* w = get_ws(h, stream);
  if (w == nullptr) return fail(h, BB_ERR_NOMEM, "workspace allocation failed");
  std::lock_guard<std::mutex> g(w->mu);
  int64_t nnz = 0;
  if (is_device_ptr(indptr)) BB_CUDA(h, cudaMemcpy(&nnz, indptr + n_users, sizeof(int64_t), cudaMemcpyDeviceToHost));
  else nnz = indptr[n_users];
  if (nnz < 0) return fail(h, BB_ERR_INVALID, "negative nnz");
  Stager st{h, w, stream};
  st.account(indptr, 8 * static_cast<size_t>(n_users + 1)); st.account(indices, 4 * static_cast<size_t>(nnz)); st.account(values, 4 * static_cast<size_t>(nnz));
  BB_CUDA(h, st.prepare());
  const void *dp, *di, *dv;
  BB_CUDA(h, st.get(indptr, 8 * static_cast<size_t>(n_users + 1), &dp));
  BB_CUDA(h, st.get(indices, 4 * static_cast<size_t>(nnz), &di));
  BB_CUDA(h, st.get(values, 4 * static_cast<size_t>(nnz), &dv));
  BB_CUDA(h, st.flush());
  const size_t ws_bytes = item_item_workspace_bytes(n_users, n_items, nnz);
  BB_CUDA(h, w->misc.ensure(ws_bytes));
  const bool s_dev = is_device_ptr(out_scores), i_dev = is_device_ptr(out_ids);
  const size_t n_out = static_cast<size_t>(n_items) * k;
  if (!s_dev) BB_CUDA(h, w->out_scores.ensure(4 * n_out));
  if (!i_dev) BB_CUDA(h, w->out_ids.ensure(4 * n_out));
  float* d_s = s_dev ? out_scores : static_cast<float*>(w->out_scores.p);
  int32_t* d_i = i_dev ? out_ids : static_cast<int32_t*>(w->out_ids.p);
  int n_launch = 0;
  BB_CUDA(h, launch_item_item(n_users, n_items, static_cast<const int64_t*>(dp), static_cast<const int32_t*>(di),
                              static_cast<const float*>(dv), k, d_s, d_i, w->misc.p, w->misc.cap, stream, &n_launch));
  h->launches += n_launch;
  if (!s_dev) BB_CUDA(h, cudaMemcpyAsync(out_scores, d_s, 4 * n_out, cudaMemcpyDeviceToHost, stream));
  if (!i_dev) BB_CUDA(h, cudaMemcpyAsync(out_ids, d_i, 4 * n_out, cudaMemcpyDeviceToHost, stream));
  if (!s_dev || !i_dev) BB_CUDA(h, cudaStreamSynchronize(stream));
  return BB_OK;
}

}  // extern "C"
