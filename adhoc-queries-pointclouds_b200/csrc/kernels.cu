// kernels.cu — hand-written sm_100a kernels of the full-scan query path.
//
// What each kernel replaces in the reference (all CPU, one sequential loop per file):
//   k_scan_staged / k_scan_direct   the per-point loops of
//        search_las_file_by_bounds_optimized            query/src/search/las.rs:101-146
//        search_las_file_by_classification_optimized    query/src/search/las.rs:221-259
//        search_last_file_by_bounds_optimized           query/src/search/last.rs:117-164
//        search_last_file_by_classification_optimized   query/src/search/last.rs:253-291
//     fused with the collector they feed (query/src/collect_points.rs):
//        MODE_COUNT  -> CountCollector::collect_one   (:84-86)   per-CTA popc, one atomic per lane
//        MODE_SELECT -> BufferCollector::collect_one  (:29-31)   stable stream compaction with a
//                                                                decoupled look-back prefix
//        MODE_GRID   -> GridSampledCollector::collect_one (:112-114) -> SparseGrid::insert_point
//                       (query/src/grid_sampling.rs:49-105) as an atomic-min cell table
//   k_class_count_soa               LAST class count fast path (1 byte per point)
//   k_grid_*                        finalisation of the density table (HashMap::values, :111-113)
//
// The work is HBM-bound integer/byte work: no tensor cores.  The staged variant moves whole record
// tiles global->shared with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) behind an mbarrier
// ring, so the LSU only sees conflict-free shared loads; the direct variant uses plain coalesced
// global loads and serves every layout/alignment.
//
// Floating point: Rust never contracts a*b+c.  Everything that feeds a stored double uses explicit
// round-to-nearest intrinsics (__dmul_rn/__dadd_rn/...) and the file is compiled with --fmad=false.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cstdint>

#include "pcq_device.h"
#include "grid_math.cuh"

// Measurement hooks of the select kernels (skip the look-back / the emit: results are WRONG).  Compiled out unless the
// library is built with -DPCQ_DEBUG_HOOKS (make EXTRA=-DPCQ_DEBUG_HOOKS); tools/dbg_select.sh then drives them through
// PCQ_SELECT_DEBUG.
#ifdef PCQ_DEBUG_HOOKS
#define PCQ_HOOK(P, bit) (((P).debug & (bit)) != 0u)
#else
#define PCQ_HOOK(P, bit) false
#endif

namespace pcq {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ bool in_range(int32_t v, int32_t lo, int32_t hi) {
  // lo <= v <= hi for lo <= hi (host guarantees), as one subtract + one unsigned compare
  return (uint32_t)(v - lo) <= (uint32_t)(hi - lo);
}

// global loads of possibly unaligned little-endian fields
__device__ __forceinline__ int32_t ldg_i32(const uint8_t* p, int align) {
  if (align == 4) return __ldg(reinterpret_cast<const int32_t*>(p));
  if (align == 2) {
    uint32_t a = __ldg(reinterpret_cast<const uint16_t*>(p));
    uint32_t b = __ldg(reinterpret_cast<const uint16_t*>(p + 2));
    return (int32_t)(a | (b << 16));
  }
  uint32_t b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2), b3 = __ldg(p + 3);
  return (int32_t)(b0 | (b1 << 8) | (b2 << 16) | (b3 << 24));
}
__device__ __forceinline__ uint32_t ldg_u16(const uint8_t* p) {
  if ((reinterpret_cast<uintptr_t>(p) & 1u) == 0) return __ldg(reinterpret_cast<const uint16_t*>(p));
  return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8);
}

struct Hit {
  int32_t x, y, z;
  uint32_t cls;
};

// ------------------------------------------------------------------------------------------------
// record sources
// ------------------------------------------------------------------------------------------------

// Records read straight from global memory (any layout, any alignment).
struct DirectSrc {
  template <bool kNeedAll>
  __device__ __forceinline__ bool eval(const Segment& S, uint32_t qkind, uint32_t qcls, uint64_t idx,
                                       uint32_t /*i_in_tile*/, Hit& h) const {
    const int al = S.align;
    if (S.layout == PCQ_LAYOUT_LAS) {
      const uint8_t* p = S.rec + idx * (uint64_t)S.record_len;
      if (qkind == PCQ_QUERY_BOUNDS) {
        h.x = ldg_i32(p, al);
        h.y = ldg_i32(p + 4, al);
        h.z = ldg_i32(p + 8, al);
        bool m = in_range(h.x, S.lo[0], S.hi[0]) & in_range(h.y, S.lo[1], S.hi[1]) &
                 in_range(h.z, S.lo[2], S.hi[2]);
        if (kNeedAll && m) h.cls = __ldg(p + S.cls_off);
        return m;
      }
      h.cls = __ldg(p + S.cls_off);
      bool m = h.cls == qcls;
      if (kNeedAll && m) {
        h.x = ldg_i32(p, al);
        h.y = ldg_i32(p + 4, al);
        h.z = ldg_i32(p + 8, al);
      }
      return m;
    }
    const uint8_t* p = S.rec + idx * 12ull;
    if (qkind == PCQ_QUERY_BOUNDS) {
      h.x = ldg_i32(p, al);
      h.y = ldg_i32(p + 4, al);
      h.z = ldg_i32(p + 8, al);
      bool m = in_range(h.x, S.lo[0], S.hi[0]) & in_range(h.y, S.lo[1], S.hi[1]) &
               in_range(h.z, S.lo[2], S.hi[2]);
      if (kNeedAll && m) h.cls = __ldg(S.cls + idx);
      return m;
    }
    h.cls = __ldg(S.cls + idx);
    bool m = h.cls == qcls;
    if (kNeedAll && m) {
      h.x = ldg_i32(p, al);
      h.y = ldg_i32(p + 4, al);
      h.z = ldg_i32(p + 8, al);
    }
    return m;
  }
  __device__ __forceinline__ void colour(const Segment& S, uint64_t idx, uint32_t /*i_in_tile*/,
                                         uint32_t rgb[3]) const {
    const uint8_t* p = nullptr;
    if (S.layout == PCQ_LAYOUT_LAS) {
      if (S.rgb_off >= 0) p = S.rec + idx * (uint64_t)S.record_len + (uint32_t)S.rgb_off;
    } else if (S.rgb != nullptr) {
      p = S.rgb + idx * 6ull;
    }
    if (p) {
      rgb[0] = ldg_u16(p);
      rgb[1] = ldg_u16(p + 2);
      rgb[2] = ldg_u16(p + 4);
    } else {
      rgb[0] = rgb[1] = rgb[2] = 0;  // Vector3::new(0, 0, 0), las.rs:134
    }
  }
};

// Records of one tile staged in shared memory (LAS records of length R, or LAST positions, R = 12).
template <int R>
struct SmemSrc {
  const uint8_t* tile;  // shared memory, 128-byte aligned

  __device__ __forceinline__ static int32_t lds_i32(const uint8_t* p) {
    if constexpr (R % 4 == 0) {
      return *reinterpret_cast<const int32_t*>(p);
    } else if constexpr (R % 2 == 0) {
      uint32_t a = *reinterpret_cast<const uint16_t*>(p);
      uint32_t b = *reinterpret_cast<const uint16_t*>(p + 2);
      return (int32_t)(a | (b << 16));
    } else {
      return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    }
  }
  __device__ __forceinline__ static uint32_t lds_u16(const uint8_t* p) {
    if constexpr (R % 2 == 0) {
      return *reinterpret_cast<const uint16_t*>(p);  // colour offsets 20 / 28 are even
    } else {
      return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
    }
  }

  template <bool kNeedAll>
  __device__ __forceinline__ bool eval(const Segment& S, uint32_t qkind, uint32_t qcls, uint64_t idx,
                                       uint32_t i, Hit& h) const {
    const uint8_t* p = tile + i * R;
    if (qkind == PCQ_QUERY_BOUNDS) {
      h.x = lds_i32(p);
      h.y = lds_i32(p + 4);
      h.z = lds_i32(p + 8);
      bool m = in_range(h.x, S.lo[0], S.hi[0]) & in_range(h.y, S.lo[1], S.hi[1]) &
               in_range(h.z, S.lo[2], S.hi[2]);
      if (kNeedAll && m) h.cls = (S.layout == PCQ_LAYOUT_LAS) ? (uint32_t)p[S.cls_off] : (uint32_t)__ldg(S.cls + idx);
      return m;
    }
    // class query: only LAS records are staged (LAST class queries use the direct kernels)
    h.cls = p[S.cls_off];
    bool m = h.cls == qcls;
    if (kNeedAll && m) {
      h.x = lds_i32(p);
      h.y = lds_i32(p + 4);
      h.z = lds_i32(p + 8);
    }
    return m;
  }
  __device__ __forceinline__ void colour(const Segment& S, uint64_t idx, uint32_t i, uint32_t rgb[3]) const {
    if (S.layout == PCQ_LAYOUT_LAS) {
      if (S.rgb_off >= 0) {
        const uint8_t* p = tile + i * R + (uint32_t)S.rgb_off;
        rgb[0] = lds_u16(p);
        rgb[1] = lds_u16(p + 2);
        rgb[2] = lds_u16(p + 4);
        return;
      }
    } else if (S.rgb != nullptr) {
      const uint8_t* p = S.rgb + idx * 6ull;
      rgb[0] = ldg_u16(p);
      rgb[1] = ldg_u16(p + 2);
      rgb[2] = ldg_u16(p + 4);
      return;
    }
    rgb[0] = rgb[1] = rgb[2] = 0;
  }
};

// ------------------------------------------------------------------------------------------------
// readers::Point (31 bytes) as 8 little-endian words (top byte of w[7] unused)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void point_words(const Segment& S, const Hit& h, const uint32_t rgb[3], uint32_t w[8]) {
  double px = reconstruct(h.x, S.scale[0], S.offset[0]);
  double py = reconstruct(h.y, S.scale[1], S.offset[1]);
  double pz = reconstruct(h.z, S.scale[2], S.offset[2]);
  unsigned long long ux = (unsigned long long)__double_as_longlong(px);
  unsigned long long uy = (unsigned long long)__double_as_longlong(py);
  unsigned long long uz = (unsigned long long)__double_as_longlong(pz);
  w[0] = (uint32_t)ux;
  w[1] = (uint32_t)(ux >> 32);
  w[2] = (uint32_t)uy;
  w[3] = (uint32_t)(uy >> 32);
  w[4] = (uint32_t)uz;
  w[5] = (uint32_t)(uz >> 32);
  w[6] = (rgb[0] & 0xFFFFu) | (rgb[1] << 16);
  w[7] = (rgb[2] & 0xFFFFu) | ((h.cls & 0xFFu) << 16);
}

// position of the n-th (0-based) set bit of `mask` (which has more than n bits set): five popc steps
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t mask, uint32_t n) {
  uint32_t pos = 0;
  uint32_t c = (uint32_t)__popc(mask & 0xFFFFu);
  if (n >= c) { n -= c; pos += 16u; mask >>= 16; }
  c = (uint32_t)__popc(mask & 0xFFu);
  if (n >= c) { n -= c; pos += 8u; mask >>= 8; }
  c = (uint32_t)__popc(mask & 0xFu);
  if (n >= c) { n -= c; pos += 4u; mask >>= 4; }
  c = (uint32_t)__popc(mask & 0x3u);
  if (n >= c) { n -= c; pos += 2u; mask >>= 2; }
  if (n >= (mask & 1u)) pos += 1u;
  return pos;
}

// Store a 31-byte record at an arbitrarily aligned shared-memory address as 7 word stores, one 16-bit
// store and one byte store (w[j] = record bytes 4j .. 4j+3, top byte of w[7] zero).
__device__ __forceinline__ void sts_point31(uint8_t* dst, const uint32_t w[8]) {
  // Every lane of a warp runs the SAME 14 stores (a 4-way `switch` on the alignment made the warp run 36, nine per
  // group of eight lanes, and the shared-memory pipe is what bounds a dense select): seven aligned words produced by
  // a funnel shift whose amount depends on the lane, plus predicated 1- and 2-byte stores for the ragged ends.
  const uint32_t o = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u);
  uint8_t* a = dst - o;  // word aligned; record byte i lives at a[o + i]
  uint32_t* d = reinterpret_cast<uint32_t*>(a) + (o != 0u ? 1u : 0u);
  const uint32_t sh = (32u - 8u * o) & 31u;  // o == 0: the words as they are
#pragma unroll
  for (int j = 0; j < 7; ++j) d[j] = __funnelshift_r(w[j], w[j + 1], sh);
  // head: the 4 - o bytes before the first aligned word
  if (o == 1u) a[1] = (uint8_t)w[0];
  if (o == 1u || o == 2u) *reinterpret_cast<uint16_t*>(a + 2) = (uint16_t)(o == 1u ? w[0] >> 8 : w[0]);
  if (o == 3u) a[3] = (uint8_t)w[0];
  // tail: what is left of w[7] (bytes 28..30 of the record) after the last aligned word
  if (o == 0u) *reinterpret_cast<uint16_t*>(a + 28) = (uint16_t)w[7];
  if (o == 0u) a[30] = (uint8_t)(w[7] >> 16);
  if (o == 2u) a[32] = (uint8_t)(w[7] >> 16);
  if (o == 3u) *reinterpret_cast<uint16_t*>(a + 32) = (uint16_t)(w[7] >> 8);
}

// the same for a record that goes straight to global memory (density winners)
__device__ __forceinline__ void stg_point31(uint8_t* dst, const uint32_t w[8]) {
  const uint32_t o = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u);
  uint8_t* a = dst - o;
  uint32_t* d = reinterpret_cast<uint32_t*>(a) + (o != 0u ? 1u : 0u);
  const uint32_t sh = (32u - 8u * o) & 31u;
#pragma unroll
  for (int j = 0; j < 7; ++j) d[j] = __funnelshift_r(w[j], w[j + 1], sh);
  if (o == 1u) a[1] = (uint8_t)w[0];
  if (o == 1u || o == 2u) *reinterpret_cast<uint16_t*>(a + 2) = (uint16_t)(o == 1u ? w[0] >> 8 : w[0]);
  if (o == 3u) a[3] = (uint8_t)w[0];
  if (o == 0u) *reinterpret_cast<uint16_t*>(a + 28) = (uint16_t)w[7];
  if (o == 0u) a[30] = (uint8_t)(w[7] >> 16);
  if (o == 2u) a[32] = (uint8_t)(w[7] >> 16);
  if (o == 3u) *reinterpret_cast<uint16_t*>(a + 32) = (uint16_t)(w[7] >> 8);
}

// ------------------------------------------------------------------------------------------------
// decoupled look-back (single pass prefix over tiles of one lane)
//   descriptor = status << 62 | value;  status 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kStatusShift = 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1ull;
constexpr unsigned long long kStAgg = 1ull;
constexpr unsigned long long kStPrefix = 2ull;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Executed by all 32 lanes of one warp.  Returns the number of matches in units
// [lane_first_tile, tile) — the exclusive prefix of `tile` within its lane.  Every lane inspects
// kLookback descriptors per round trip (window of 32 * kLookback units); each load instruction of the
// warp covers 32 consecutive descriptors (256 bytes).
constexpr int kLookback = 8;

template <bool WAIT = true>
__device__ __forceinline__ bool lookback_window(const unsigned long long* state, uint64_t tile, uint64_t lane_first_tile,
                                                unsigned long long& excl_out) {
  unsigned long long excl = 0;
  long long hi = (long long)tile - 1;
  const long long lo = (long long)lane_first_tile;
  const uint32_t ln = lane_id();
  while (hi >= lo) {
    unsigned long long sv[kLookback];  // sv[k] = descriptor at distance 32 k + lane below `hi` (coalesced)
    bool pending;
    do {
      pending = false;
#pragma unroll
      for (int k = 0; k < kLookback; ++k) {
        const long long t = hi - (long long)(32u * (uint32_t)k + ln);
        sv[k] = t >= lo ? ld_state(state + t * (long long)kDescStride) : (kStPrefix << kStatusShift);  // below the lane start: prefix 0
        pending |= (sv[k] >> kStatusShift) == 0ull;
      }
      pending = __any_sync(0xffffffffu, pending);
      // (a descriptor that is not ready only matters when it lies above the closest prefix; waiting for all of the
      // window is simpler and the window is published within the same microsecond)
      if (!WAIT && pending) return false;
    } while (pending);
    // closest descriptor that already carries an inclusive prefix: row kf, lane lf
    int kf = kLookback;
    uint32_t lf = 32u;
#pragma unroll
    for (int k = kLookback - 1; k >= 0; --k) {
      const uint32_t pmk = __ballot_sync(0xffffffffu, (sv[k] >> kStatusShift) == kStPrefix);
      if (pmk) {
        kf = k;
        lf = (uint32_t)__ffs((int)pmk) - 1u;
      }
    }
    const uint32_t pm = kf < kLookback ? 1u : 0u;
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < kLookback; ++k) {
      const bool take = k < kf || (k == kf && ln <= lf);
      if (take) v += sv[k] & kValueMask;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (pm != 0u) break;
    if (!WAIT) return false;  // one window, one round trip: no prefix in reach
    hi -= 32 * kLookback;
  }
  excl_out = excl;
  return true;
}

__device__ __forceinline__ unsigned long long lookback_exclusive(const unsigned long long* state, uint64_t tile,
                                                                 uint64_t lane_first_tile) {
  unsigned long long excl = 0;
  lookback_window<true>(state, tile, lane_first_tile, excl);
  return excl;
}

// ------------------------------------------------------------------------------------------------
// SparseGrid::insert_point as an atomic-min table (grid_sampling.rs:49-105)
// ------------------------------------------------------------------------------------------------

// find-or-insert the slot of `key`.  dense: slot == key.  hashed: linear probing on hkeys.
__device__ __forceinline__ uint64_t grid_slot(const GridDev& g, uint64_t key, bool insert) {
  if (g.hkeys == nullptr) {
    if (!g.sub_on) return key;  // dense over the whole grid: slot == key
    // Dense over a sub-box: a collector that is fed by one file (run_search_parallel makes one grid per file,
    // main.rs:253-273) only ever sees the cells under that file's header box — a 64th of the doc grid — so its table
    // covers those cells only.  A cell outside (a header that lies about its bounds) reports "no slot": the host then
    // moves the collector to a table over the whole grid and runs the launch again (rehash_grid).
    const uint64_t dx = (key & g.mask[0]) - g.sub_lo[0];
    const uint64_t dy = ((key >> g.shift_y) & g.mask[1]) - g.sub_lo[1];
    const uint64_t dz = (key >> g.shift_z) - g.sub_lo[2];
    if (dx >= g.sub_n[0] || dy >= g.sub_n[1] || dz >= g.sub_n[2]) return ~0ull;
    return dx + g.sub_n[0] * (dy + g.sub_n[1] * dz);
  }
  const uint64_t mask = g.table_slots - 1ull;
  uint64_t s = mix64(key) & mask;
  for (uint64_t probe = 0; probe < g.table_slots; ++probe) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(g.hkeys + s);
    if (cur == key) return s;
    if (cur == ~0ull) {
      if (!insert) return ~0ull;
      unsigned long long prev = atomicCAS(g.hkeys + s, ~0ull, (unsigned long long)key);
      if (prev == ~0ull || prev == key) return s;
    }
    s = (s + 1ull) & mask;
  }
  return ~0ull;
}

// Candidate arena allocation.  One global counter bumped once per warp and append serialises on its L2 atomic
// unit (it was a quarter of the insert kernel's time); instead every warp owns a private chunk of kCandChunk
// (pcq_device.h) slots, fills it without any atomic and takes a new chunk with ONE atomicAdd when it is full.  Slots a warp
// reserved but did not use are marked empty (scan_idx == ~0) and skipped by every consumer of the arena.

struct CandChunk {
  unsigned long long base = 0;
  uint32_t used = kCandChunk;  // == kCandChunk: no chunk yet
};

__device__ __forceinline__ void cand_pad(const GridDev& g, const CandChunk& ch) {
  for (uint32_t i = ch.used + lane_id(); i < kCandChunk; i += 32u) {
    const unsigned long long ci = ch.base + i;
    if (ci < g.cand_cap) g.cands[ci].scan_idx = kCandEmpty;
  }
}

struct LaneChunk {  // a warp's chunk belongs to the arena of one lane (collector)
  CandChunk c;
  uint32_t lane = 0xFFFFFFFFu;
};

// all 32 lanes; n (1..32) is warp-uniform; returns the arena index of the group's first slot
__device__ __forceinline__ unsigned long long cand_reserve(const GridDev& g, CandChunk& ch, uint32_t n) {
  if (ch.used + n > kCandChunk) {
    cand_pad(g, ch);
    unsigned long long nb = 0;
    if (lane_id() == 0) nb = atomicAdd(g.cand_count, (unsigned long long)kCandChunk);
    ch.base = __shfl_sync(0xffffffffu, nb, 0);
    ch.used = 0;
  }
  const unsigned long long r = ch.base + ch.used;
  ch.used += n;
  return r;
}

// One queued match of the sparse density insert (see process_tile): everything the insert needs, so that the tile the
// point came from can be released.
struct alignas(16) GridQEntry {
  int32_t x, y, z;
  uint32_t rg;   // r | g << 16
  uint32_t bc;   // b | cls << 16
  uint32_t pad_;
  unsigned long long gidx;  // collector-wide scan index
};

// Where the kPPT points of a lane come from: straight from the tile ...
template <class Src>
struct TileFetch {
  Src src;
  Hit h[kPPT];  // (by value: a reference would force the array into local memory)
  uint64_t p0;
  __device__ __forceinline__ void xyz(int j, int32_t& x, int32_t& y, int32_t& z) const {
    x = h[j].x;
    y = h[j].y;
    z = h[j].z;
  }
  __device__ __forceinline__ void words(const Segment& S, int j, uint32_t w[8]) const {
    const uint32_t i = (uint32_t)j * kBlock + threadIdx.x;
    uint32_t rgb[3];
    src.colour(S, p0 + i, i, rgb);
    point_words(S, h[j], rgb, w);
  }
  __device__ __forceinline__ unsigned long long gidx(const Segment& S, int j) const {
    return S.scan_base + p0 + (uint32_t)j * kBlock + threadIdx.x;
  }
};
// ... or from the CTA's match queue (entries stay in shared memory until the flush ends)
struct QueueFetch {
  const GridQEntry* qe[kPPT];
  __device__ __forceinline__ void xyz(int j, int32_t& x, int32_t& y, int32_t& z) const {
    x = qe[j]->x;
    y = qe[j]->y;
    z = qe[j]->z;
  }
  __device__ __forceinline__ void words(const Segment& S, int j, uint32_t w[8]) const {
    const GridQEntry& q = *qe[j];
    Hit h;
    h.x = q.x;
    h.y = q.y;
    h.z = q.z;
    h.cls = q.bc >> 16;
    const uint32_t rgb[3] = {q.rg & 0xFFFFu, q.rg >> 16, q.bc & 0xFFFFu};
    point_words(S, h, rgb, w);
  }
  __device__ __forceinline__ unsigned long long gidx(const Segment&, int j) const { return qe[j]->gidx; }
};

// rare path, kept out of line so that it costs the insert kernel neither registers nor instruction-cache space.
// Everything travels BY VALUE: a reference to the caller's point registers would force them into local memory, for
// every point of every tile (that was 1.8 GB of local stores per navvis-XL launch).
__device__ __noinline__ void grid_log_entry(Candidate* log, unsigned long long* log_count, uint64_t log_cap, uint32_t* flags,
                                            uint64_t key, unsigned long long gidx, uint4 wa, uint4 wb) {
  const unsigned long long li = atomicAdd(log_count, 1ull);
  if (li < log_cap) {
    uint4* c4 = reinterpret_cast<uint4*>(log + li);
    c4[0] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), 0u, 0u);
    c4[1] = make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), wa.x, wa.y);
    c4[2] = make_uint4(wa.z, wa.w, wb.x, wb.y);
    c4[3] = make_uint4(wb.z, wb.w & 0x00FFFFFFu, 0u, 0u);
  } else {
    atomicOr(flags, kFlagLogOverflow);
  }
}
template <class Fetch>
__device__ __forceinline__ void grid_log_point(const GridDev& g, const Segment& S, const Fetch& f, int j, uint64_t key) {
  uint32_t w[8];
  f.words(S, j, w);
  grid_log_entry(g.log, g.log_count, g.log_cap, g.flags, key, f.gidx(S, j), make_uint4(w[0], w[1], w[2], w[3]),
                 make_uint4(w[4], w[5], w[6], w[7]));
}

// L2 residency of the density insert: the cell table is the one structure that is re-read (every matching point reads
// its cell), while records stream through once and candidates are written once — table accesses ask L2 to keep their
// lines (evict_last), candidate stores not to (evict_first), so that a table whose touched sectors fit L2 stays there.
__device__ __forceinline__ uint64_t l2_keep_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_stream_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long ld_table(const unsigned long long* p, uint64_t pol) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void red_min_table(unsigned long long* p, unsigned long long v, uint64_t pol) {
  asm volatile("red.relaxed.gpu.global.min.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol));
}
// One whole 32-byte sector per store (sm_100: 256-bit st.global, SASS STG.256): half the store instructions of four
// 16-byte stores per 64-byte candidate, and no half-written sectors in L2.
__device__ __forceinline__ void stg_sector_stream(void* p, const uint4& a, const uint4& b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "l"(pol));
}

// Warp-convergent: every lane calls it; m[j] says whether this lane's j-th point matches.
// A point survives as a *candidate* iff its distance is <= the cell minimum it reads (a possibly stale value that is
// never below the true minimum, so the true winner — smallest distance, then smallest scan index == the strict `<`
// fold of :97-102 — always survives); the minimum itself is maintained with a fire-and-forget red.min, so nothing
// waits for an atomic's round trip.  The kPPT points of a lane are taken through each step together so that their
// table reads are in flight at the same time.
template <class Fetch, int kPPT>
__device__ __forceinline__ void grid_insert_tile(const GridDev& g, const Segment& S, const bool (&m)[kPPT], const Fetch& f,
                                                 CandChunk& ch) {
  bool any_m = false;
#pragma unroll
  for (int j = 0; j < kPPT; ++j) any_m |= m[j];
  if (!__any_sync(0xffffffffu, any_m)) return;  // most tiles of a small query box hold no match at all
  const uint64_t keep = l2_keep_policy(), stream = l2_stream_policy();
  uint64_t key[kPPT], slot[kPPT];
  unsigned long long dist[kPPT], seen[kPPT];
  bool live[kPPT];
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    live[j] = false;
    slot[j] = 0;
    key[j] = 0;
    dist[j] = 0ull;
    if (m[j]) {
      int32_t vx, vy, vz;
      f.xyz(j, vx, vy, vz);
      const double px = reconstruct(vx, S.scale[0], S.offset[0]);
      const double py = reconstruct(vy, S.scale[1], S.offset[1]);
      const double pz = reconstruct(vz, S.scale[2], S.offset[2]);
      const CellEval e = grid_eval(g, px, py, pz);
      key[j] = e.key;
      dist[j] = e.dist_bits;
      if (e.aliased || alias_find(g, e.key) != ~0u) {
        // a point of an affected key: logged for the ordered replay, never enters the table
        grid_log_point(g, S, f, j, e.key);
      } else if (!g.log_only) {
        slot[j] = grid_slot(g, e.key, true);
        if (slot[j] == ~0ull)
          atomicOr(g.flags, kFlagHashFull);
        else
          live[j] = true;
      }
    }
  }
  // the table reads of a lane are in flight together (random accesses into a table far larger than L1)
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    seen[j] = 0ull;
    if (live[j]) seen[j] = ld_table(g.table + slot[j], keep);
  }
  bool want[kPPT];
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    // The cell minimum only ever decreases, so a (possibly stale) read that is already smaller than this point's
    // distance proves the point can never win: no atomic, no candidate.  In dense data that is most points.
    want[j] = live[j] && dist[j] <= seen[j];
    if (want[j] && dist[j] < seen[j]) red_min_table(g.table + slot[j], dist[j], keep);
  }
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    const uint32_t bal = __ballot_sync(0xffffffffu, want[j]);
    if (bal == 0u) continue;
    const unsigned long long base = cand_reserve(g, ch, (uint32_t)__popc(bal));
    if (want[j]) {
      const unsigned long long ci = base + (unsigned long long)__popc(bal & ((1u << lane_id()) - 1u));
      if (ci < g.cand_cap) {
        uint32_t w[8];
        f.words(S, j, w);
        uint4* c4 = reinterpret_cast<uint4*>(g.cands + ci);
        const unsigned long long gidx = f.gidx(S, j);
        stg_sector_stream(c4, make_uint4((uint32_t)key[j], (uint32_t)(key[j] >> 32), (uint32_t)dist[j], (uint32_t)(dist[j] >> 32)),
                          make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), w[0], w[1]), stream);
        stg_sector_stream(c4 + 2, make_uint4(w[2], w[3], w[4], w[5]), make_uint4(w[6], w[7] & 0x00FFFFFFu, 0u, 0u), stream);
      } else {
        atomicOr(g.flags, kFlagCandOverflow);
      }
    }
  }
}

// Index of the segment that holds `tile`, searching forward from `cur` (segments are ordered by first_tile).  A launch
// over the surviving chunk runs of an index (api.cu) can hold thousands of short segments, and a CTA's next tile lies
// gridDim tiles further on: gallop, then bisect, instead of walking.
__device__ __forceinline__ uint32_t seg_forward(const ScanParams& P, uint32_t cur, uint64_t tile) {
  if (cur + 1 < P.n_segs && tile >= P.segs[cur + 1].first_tile) {
    uint32_t lo = cur + 1, step = 1;
    while (lo + step < P.n_segs && tile >= P.segs[lo + step].first_tile) {
      lo += step;
      step <<= 1;
    }
    uint32_t hi = lo + step < P.n_segs ? lo + step : P.n_segs;  // segs[lo].first_tile <= tile < segs[hi].first_tile
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (tile >= P.segs[mid].first_tile) lo = mid; else hi = mid;
    }
    cur = lo;
  }
  return cur;
}

// ------------------------------------------------------------------------------------------------
// per-tile work shared by the direct and the staged scan kernels (MODE_COUNT / MODE_GRID)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long* scratch /* kBlock/32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane_id() == 0) scratch[warp_id()] = v;
  __syncthreads();
  unsigned long long t = 0;
#pragma unroll
  for (int w = 0; w < kBlock / 32; ++w) t += scratch[w];
  return t;
}

// Match queue of the density insert (MODE_GRIDQ).  With a small query box only a few lanes of a warp carry a match,
// yet the insert (three IEEE divisions, table read, atomic, candidate write) costs a warp the same whether 1 or 32
// lanes take part — a 3 % box paid two thirds of the all-match price.  Matching points are queued in shared memory
// as they are found and inserted kGridQFlush at a time with every lane busy; the queue is emptied before the CTA
// moves to another segment and at the end.  Dense queries keep the direct path (MODE_GRID): the host picks per launch.
constexpr uint32_t kGridQFlush = kBlock;                     // insert when at least this many matches are queued
constexpr uint32_t kGridQCap = kGridQFlush - 1u + kTilePts;  // one more tile always fits

struct GridQ {
  GridQEntry* q;
  uint32_t* n;
};

template <int MODE, class Src>
__device__ __forceinline__ void process_tile(const ScanParams& P, const Segment& S, uint64_t tile, const Src& src,
                                             unsigned long long& acc, LaneChunk& ch, const GridQ& gq) {
  static_assert(MODE == MODE_COUNT || MODE == MODE_GRID || MODE == MODE_GRIDQ, "select has its own kernels");
  const uint64_t p0 = (tile - S.first_tile) * (uint64_t)kTilePts;
  const uint64_t rem = S.n_points - p0;
  const uint32_t npts = rem < (uint64_t)kTilePts ? (uint32_t)rem : (uint32_t)kTilePts;
  const uint32_t tid = threadIdx.x;

  Hit h[kPPT];
  bool m[kPPT];
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    const uint32_t i = (uint32_t)j * kBlock + tid;
    m[j] = false;
    if (i < npts) m[j] = src.template eval<MODE != MODE_COUNT>(S, P.query_kind, P.cls, p0 + i, i, h[j]);
  }

  if constexpr (MODE == MODE_COUNT) {
#pragma unroll
    for (int j = 0; j < kPPT; ++j) acc += m[j] ? 1ull : 0ull;
  } else if constexpr (MODE == MODE_GRID) {
    const GridDev& g = P.lanes[S.lane].grid;
    if (ch.lane != S.lane) {  // a warp's chunk belongs to one collector's arena
      if (ch.lane != 0xFFFFFFFFu) cand_pad(P.lanes[ch.lane].grid, ch.c);
      ch.c = CandChunk();
      ch.lane = S.lane;
    }
    TileFetch<Src> f{src, {}, p0};
#pragma unroll
    for (int j = 0; j < kPPT; ++j) f.h[j] = h[j];
    grid_insert_tile(g, S, m, f, ch.c);
  } else {
#pragma unroll
    for (int j = 0; j < kPPT; ++j) {
      const uint32_t bal = __ballot_sync(0xffffffffu, m[j]);
      if (bal == 0u) continue;
      uint32_t base = 0;
      if (lane_id() == 0) base = atomicAdd(gq.n, (uint32_t)__popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (m[j]) {
        const uint32_t i = (uint32_t)j * kBlock + tid;
        uint32_t rgb[3];
        src.colour(S, p0 + i, i, rgb);
        GridQEntry& e = gq.q[base + (uint32_t)__popc(bal & ((1u << lane_id()) - 1u))];
        e.x = h[j].x;
        e.y = h[j].y;
        e.z = h[j].z;
        e.rg = (rgb[0] & 0xFFFFu) | (rgb[1] << 16);
        e.bc = (rgb[2] & 0xFFFFu) | ((h[j].cls & 0xFFu) << 16);
        e.gidx = S.scan_base + p0 + i;
      }
    }
  }
}

// MODE_GRIDQ: insert everything that is queued (all threads of the CTA; S is the segment the queued points belong to)
__device__ __forceinline__ void grid_flush(const ScanParams& P, const Segment& S, const GridQ& gq, LaneChunk& ch) {
  __syncthreads();  // every queued entry is written
  const uint32_t n = *gq.n;
  if (n != 0u) {
    const GridDev& g = P.lanes[S.lane].grid;
    if (ch.lane != S.lane) {
      if (ch.lane != 0xFFFFFFFFu) cand_pad(P.lanes[ch.lane].grid, ch.c);
      ch.c = CandChunk();
      ch.lane = S.lane;
    }
    for (uint32_t base = 0; base < n; base += kBlock * kPPT) {
      bool m[kPPT];
      QueueFetch f;
#pragma unroll
      for (int j = 0; j < kPPT; ++j) {
        const uint32_t e = base + (uint32_t)j * kBlock + threadIdx.x;
        m[j] = e < n;
        f.qe[j] = gq.q + (m[j] ? e : 0u);
      }
      grid_insert_tile(g, S, m, f, ch.c);
    }
  }
  __syncthreads();  // every entry is consumed
  if (threadIdx.x == 0) *gq.n = 0u;
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// direct scan kernel: persistent CTAs, plain global loads
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_scan_direct(ScanParams P) {
  __shared__ Segment sseg;
  __shared__ unsigned long long s_red[kBlock / 32];
  __shared__ GridQEntry s_q[MODE == MODE_GRIDQ ? kGridQCap : 1];
  __shared__ uint32_t s_qn;
  const GridQ gq{s_q, &s_qn};
  if (threadIdx.x == 0) s_qn = 0u;
  __syncthreads();

  uint32_t seg_i = 0xFFFFFFFFu;
  uint32_t seg_cursor = 0;
  unsigned long long acc = 0;
  LaneChunk lch;
  DirectSrc src;

  for (uint64_t iter = 0;; ++iter) {
    const uint64_t tile = (uint64_t)blockIdx.x + iter * (uint64_t)gridDim.x;
    if (tile >= P.n_tiles) break;
    seg_cursor = seg_forward(P, seg_cursor, tile);
    if (seg_cursor != seg_i) {
      // segment change: flush the per-lane match count, cache the new segment in shared memory
      if constexpr (MODE == MODE_COUNT) {
        if (seg_i != 0xFFFFFFFFu) {
          unsigned long long t = block_sum(acc, s_red);
          if (threadIdx.x == 0 && t) atomicAdd(P.lanes[sseg.lane].count, t);
          acc = 0;
        }
      }
      if constexpr (MODE == MODE_GRIDQ) {
        if (seg_i != 0xFFFFFFFFu) grid_flush(P, sseg, gq, lch);
      }
      __syncthreads();
      const uint32_t* srcw = reinterpret_cast<const uint32_t*>(P.segs + seg_cursor);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&sseg);
      for (uint32_t k = threadIdx.x; k < sizeof(Segment) / 4; k += kBlock) dstw[k] = srcw[k];
      seg_i = seg_cursor;
      __syncthreads();
    }
    process_tile<MODE>(P, sseg, tile, src, acc, lch, gq);
    if constexpr (MODE == MODE_GRIDQ) {
      __syncthreads();
      if (s_qn >= kGridQFlush) grid_flush(P, sseg, gq, lch);
    }
  }
  if constexpr (MODE == MODE_COUNT) {
    if (seg_i != 0xFFFFFFFFu) {
      unsigned long long t = block_sum(acc, s_red);
      if (threadIdx.x == 0 && t) atomicAdd(P.lanes[sseg.lane].count, t);
    }
  }
  if constexpr (MODE == MODE_GRIDQ) {
    if (seg_i != 0xFFFFFFFFu) grid_flush(P, sseg, gq, lch);
  }
  if constexpr (MODE == MODE_GRID || MODE == MODE_GRIDQ) {
    if (lch.lane != 0xFFFFFFFFu) cand_pad(P.lanes[lch.lane].grid, lch.c);
  }
}

// ------------------------------------------------------------------------------------------------
// staged scan kernel: record tiles travel global -> shared as 1-D bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx) into a STAGES-deep ring; thread 0 is the producer.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// same, with an L2 cache policy (record tiles are read once: evict_first keeps them from displacing the density table)
__device__ __forceinline__ void bulk_copy_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// the same with a suspend-time hint (ns): the thread may sleep in hardware until the phase completes or the time is up
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra SWAIT_DONE;\n"
      "bra SWAIT_LOOP;\n"
      "SWAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(2000u)
      : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// LAST positions (12-byte records) in count mode: 1024-record tiles, each thread takes 4 consecutive records with
// three conflict-free 16-byte shared loads — a quarter of the load instructions of the generic path, which is what
// a 12-byte-per-point scan needs to stay memory- rather than issue-bound.
constexpr int kTilePtsPos = 1024;

template <int R, int MODE, int STAGES, int TP = kTilePts>
__global__ void __launch_bounds__(kBlock) k_scan_staged(ScanParams P) {
  constexpr uint32_t kTileBytes = (uint32_t)TP * (uint32_t)R;
  extern __shared__ __align__(128) uint8_t dsm[];  // STAGES * kTileBytes
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ unsigned long long stage_tile[STAGES];
  __shared__ uint32_t stage_seg[STAGES];
  __shared__ Segment sseg;
  __shared__ unsigned long long s_red[kBlock / 32];

  const uint32_t tid = threadIdx.x;
  // producer state (meaningful in thread 0 only)
  uint64_t stream_policy = 0;
  if constexpr (MODE == MODE_GRID || MODE == MODE_GRIDQ) {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream_policy));
  }
  uint32_t prod_seg = 0;
  uint64_t prod_iter = 0;

  auto produce = [&](int s) {
    const uint64_t tile = (uint64_t)blockIdx.x + prod_iter * (uint64_t)gridDim.x;
    ++prod_iter;
    if (tile >= P.n_tiles) {
      stage_tile[s] = ~0ull;
      return;
    }
    prod_seg = seg_forward(P, prod_seg, tile);
    const Segment* sg = P.segs + prod_seg;
    const uint64_t p0 = (tile - sg->first_tile) * (uint64_t)TP;
    const uint64_t rem = sg->n_points - p0;
    const uint32_t npts = rem < (uint64_t)TP ? (uint32_t)rem : (uint32_t)TP;
    const uint32_t bytes = (npts * (uint32_t)R + 15u) & ~15u;  // bulk copies move multiples of 16 bytes
    stage_tile[s] = tile;
    stage_seg[s] = prod_seg;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    if constexpr (MODE == MODE_GRID || MODE == MODE_GRIDQ)
      bulk_copy_g2s_hint(dsm + (size_t)s * kTileBytes, sg->rec + p0 * (uint64_t)R, bytes, &full_bar[s], stream_policy);
    else
      bulk_copy_g2s(dsm + (size_t)s * kTileBytes, sg->rec + p0 * (uint64_t)R, bytes, &full_bar[s]);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1u);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll 1
    for (int s = 0; s < STAGES; ++s) produce(s);
  }
  __syncthreads();

  __shared__ GridQEntry s_q[MODE == MODE_GRIDQ ? kGridQCap : 1];
  __shared__ uint32_t s_qn;
  const GridQ gq{s_q, &s_qn};
  if (tid == 0) s_qn = 0u;
  __syncthreads();

  uint32_t seg_i = 0xFFFFFFFFu;
  unsigned long long acc = 0;
  LaneChunk lch;
  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it % STAGES;
    const uint32_t parity = (it / STAGES) & 1u;
    const unsigned long long tile = stage_tile[s];
    if (tile == ~0ull) break;
    const uint32_t seg_now = stage_seg[s];
    if (seg_now != seg_i) {
      if constexpr (MODE == MODE_COUNT) {
        if (seg_i != 0xFFFFFFFFu) {
          unsigned long long t = block_sum(acc, s_red);
          if (tid == 0 && t) atomicAdd(P.lanes[sseg.lane].count, t);
          acc = 0;
        }
      }
      if constexpr (MODE == MODE_GRIDQ) {
        if (seg_i != 0xFFFFFFFFu) grid_flush(P, sseg, gq, lch);
      }
      __syncthreads();
      const uint32_t* srcw = reinterpret_cast<const uint32_t*>(P.segs + seg_now);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&sseg);
      for (uint32_t k = tid; k < sizeof(Segment) / 4; k += kBlock) dstw[k] = srcw[k];
      seg_i = seg_now;
      __syncthreads();
    }
    mbar_wait(&full_bar[s], parity);
    if constexpr (R == 12 && MODE == MODE_COUNT && TP == kTilePtsPos) {
      const Segment& S = sseg;
      const uint64_t rem = S.n_points - (tile - S.first_tile) * (uint64_t)TP;
      const uint32_t npts = rem < (uint64_t)TP ? (uint32_t)rem : (uint32_t)TP;
      const uint4* q = reinterpret_cast<const uint4*>(dsm + (size_t)s * kTileBytes + 48u * tid);  // records 4 tid .. 4 tid + 3
      const uint4 a = q[0], b = q[1], c = q[2];
      const uint32_t i0 = 4u * tid;
      const int32_t x0 = (int32_t)a.x, y0 = (int32_t)a.y, z0 = (int32_t)a.z;
      const int32_t x1 = (int32_t)a.w, y1 = (int32_t)b.x, z1 = (int32_t)b.y;
      const int32_t x2 = (int32_t)b.z, y2 = (int32_t)b.w, z2 = (int32_t)c.x;
      const int32_t x3 = (int32_t)c.y, y3 = (int32_t)c.z, z3 = (int32_t)c.w;
      const int32_t lx = S.lo[0], hx = S.hi[0], ly = S.lo[1], hy = S.hi[1], lz = S.lo[2], hz = S.hi[2];
      uint32_t n = 0;
      n += (i0 + 0u < npts) & in_range(x0, lx, hx) & in_range(y0, ly, hy) & in_range(z0, lz, hz);
      n += (i0 + 1u < npts) & in_range(x1, lx, hx) & in_range(y1, ly, hy) & in_range(z1, lz, hz);
      n += (i0 + 2u < npts) & in_range(x2, lx, hx) & in_range(y2, ly, hy) & in_range(z2, lz, hz);
      n += (i0 + 3u < npts) & in_range(x3, lx, hx) & in_range(y3, ly, hy) & in_range(z3, lz, hz);
      acc += n;
    } else {
      SmemSrc<R> src{dsm + (size_t)s * kTileBytes};
      process_tile<MODE>(P, sseg, tile, src, acc, lch, gq);
    }
    __syncthreads();  // every thread is done reading stage s
    if (tid == 0) produce((int)s);
    if constexpr (MODE == MODE_GRIDQ) {
      if (s_qn >= kGridQFlush) grid_flush(P, sseg, gq, lch);  // (the barrier above ordered the queue counter)
    }
  }
  if constexpr (MODE == MODE_COUNT) {
    if (seg_i != 0xFFFFFFFFu) {
      unsigned long long t = block_sum(acc, s_red);
      if (tid == 0 && t) atomicAdd(P.lanes[sseg.lane].count, t);
    }
  }
  if constexpr (MODE == MODE_GRIDQ) {
    if (seg_i != 0xFFFFFFFFu) grid_flush(P, sseg, gq, lch);
  }
  if constexpr (MODE == MODE_GRID || MODE == MODE_GRIDQ) {
    if (lch.lane != 0xFFFFFFFFu) cand_pad(P.lanes[lch.lane].grid, lch.c);
  }
}

// ------------------------------------------------------------------------------------------------
// k_grid_scan — the density scan (MODE_GRID) as a warp-specialised kernel.
//
// k_scan_staged ends every tile with a CTA-wide barrier (its producer is also a consumer), which is free for a count
// but not for the density insert: a tile's insert is a chain of long-latency operations (table read, atomic,
// candidate stores), and the barrier made every warp of the CTA wait for the slowest one, tile after tile (10 % of
// the samples on the barrier, 25 % on the scoreboards behind it).  Here a PRODUCER warp streams record tiles into a
// ring with bulk async copies and eight CONSUMER warps meet it on mbarriers only: a consumer copies the fields of
// its 64 records of a tile into registers, releases the stage at once and then runs its insert on its own, so the
// warps of a CTA drift apart and one warp's table misses are covered by the others' arithmetic.
// QUEUE = true (sparse matches, chosen by the host like MODE_GRIDQ): matching points are appended to a per-warp queue
// in shared memory (no atomics: the warp is the only writer) and inserted 64 at a time with every lane busy.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kGsQFlush = 64;                // insert when a warp has queued this many matches
template <int PPT>
struct GsShape {
  static constexpr int kWarpPts = 32 * PPT;               // records of a tile per consumer warp
  static constexpr int kWarps = kTilePts / kWarpPts;      // consumer warps
  static constexpr int kThreads = (kWarps + 1) * 32;      // + the producer warp
  static constexpr uint32_t kQCap = kGsQFlush - 1u + (uint32_t)kWarpPts + 1u;  // one more tile always fits
};

// Records of a tile (or of any 16-byte aligned run of records) in shared memory, taken out with aligned 32-bit loads
// and one funnel shift per field: 26- and 34-byte records are only 2-byte aligned, and 16-bit loads + merges were a
// sixth of the density insert's instructions.
template <int R>
struct UnitSrc {
  static_assert(R % 2 == 0, "record fields must be 2-byte aligned in the unit buffer");
  const uint8_t* unit;  // shared memory, 16-byte aligned
  __device__ __forceinline__ const uint32_t* words(uint32_t byte_off) const {
    return reinterpret_cast<const uint32_t*>(unit + (byte_off & ~3u));
  }
  __device__ __forceinline__ void xyz(uint32_t i, int32_t& x, int32_t& y, int32_t& z) const {
    const uint32_t b = i * (uint32_t)R;
    const uint32_t* w = words(b);
    if constexpr (R % 4 == 0) {
      x = (int32_t)w[0];
      y = (int32_t)w[1];
      z = (int32_t)w[2];
    } else {
      const uint32_t sh = (b & 2u) * 8u;  // 0 or 16
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
      x = (int32_t)__funnelshift_r(w0, w1, sh);
      y = (int32_t)__funnelshift_r(w1, w2, sh);
      z = (int32_t)__funnelshift_r(w2, w3, sh);
    }
  }
  // r | g << 16 and b of the colour at (even) byte offset o; reads the aligned word pair that holds bytes o .. o+5
  __device__ __forceinline__ void rgb(uint32_t o, uint32_t& rg, uint32_t& bl) const {
    const uint32_t* w = words(o);
    const uint32_t sh = (o & 2u) * 8u;
    const uint32_t a = w[0], c = w[1];
    rg = __funnelshift_r(a, c, sh);
    bl = (c >> sh) & 0xFFFFu;
  }
};


// the PPT points of a lane, held in registers (the stage they came from is long gone when they are inserted)
template <int PPT>
struct RegFetch {
  static constexpr int N = PPT;
  int32_t x[PPT], y[PPT], z[PPT];
  uint32_t rg[PPT], bc[PPT];  // r | g << 16,  b | cls << 16
  unsigned long long gi0;      // collector-wide scan index of point 0 of this lane; point j is 32 j further on
  __device__ __forceinline__ void xyz(int j, int32_t& vx, int32_t& vy, int32_t& vz) const {
    vx = x[j];
    vy = y[j];
    vz = z[j];
  }
  __device__ __forceinline__ void words(const Segment& S, int j, uint32_t w[8]) const {
    Hit h;
    h.x = x[j];
    h.y = y[j];
    h.z = z[j];
    h.cls = bc[j] >> 16;
    const uint32_t rgb[3] = {rg[j] & 0xFFFFu, rg[j] >> 16, bc[j] & 0xFFFFu};
    point_words(S, h, rgb, w);
  }
  __device__ __forceinline__ unsigned long long gidx(const Segment&, int j) const { return gi0 + 32ull * (unsigned)j; }
};

// PPT: records per lane and tile (2: eight consumer warps per CTA).  Measured at navvis-XL: with 4 records per lane
// (four consumer warps, 96 registers, four CTAs per SM) the insert takes 1.95 instead of 1.35 ms, and so does every
// other way of putting more table reads in flight (56 registers and four CTAs: 1.80 ms; one ring per warp, 24 warps:
// 1.82 ms) — the random reads of the cell table are bound by DRAM, not by the SMs' ability to issue them.
// ONE: the launch has one collector — its grid is read from the kernel parameters (constant-bank operands).
template <int R, int STAGES, bool QUEUE, int PPT, bool ONE>
__global__ void __launch_bounds__(GsShape<PPT>::kThreads, QUEUE ? 2 : 3) k_grid_scan(const __grid_constant__ ScanParams P) {
  using Sh = GsShape<PPT>;
  constexpr int kWarps = Sh::kWarps;
  constexpr uint32_t kTileBytes = (uint32_t)kTilePts * (uint32_t)R;
  extern __shared__ __align__(128) uint8_t dsm[];  // STAGES * kTileBytes [+ kWarps * kQCap queue entries]
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ unsigned long long st_tile[STAGES];
  __shared__ uint32_t st_segi[STAGES];
  __shared__ Segment st_seg[STAGES];
  __shared__ Segment wseg[kWarps];

  const uint32_t wid = warp_id(), ln = lane_id();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1u);
      mbar_init(&empty_bar[s], (uint32_t)kWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (wid == (uint32_t)kWarps) {
    // ---------------- producer warp ----------------
    uint64_t stream_policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream_policy));
    uint32_t seg = 0;
    for (uint32_t it = 0;; ++it) {
      const uint32_t s = it % STAGES;
      // (every consumer has left the stage; the producer has nothing else to do, so it sleeps in the wait instead of
      // spinning: its polling was 9 % of the kernel's issued instructions)
      if (it >= (uint32_t)STAGES) mbar_wait_sleepy(&empty_bar[s], ((it / STAGES) - 1u) & 1u);
      const uint64_t tile = (uint64_t)blockIdx.x + (uint64_t)it * (uint64_t)gridDim.x;
      if (tile >= P.n_tiles) {
        if (ln == 0) {
          st_tile[s] = ~0ull;
          mbar_arrive(&full_bar[s]);
        }
        break;
      }
      seg = seg_forward(P, seg, tile);
      const Segment* sg = P.segs + seg;
      {  // the tile's segment travels with it: consumers are not in step with each other
        const uint32_t* srcw = reinterpret_cast<const uint32_t*>(sg);
        uint32_t* dstw = reinterpret_cast<uint32_t*>(&st_seg[s]);
        for (uint32_t k = ln; k < sizeof(Segment) / 4; k += 32u) dstw[k] = srcw[k];
      }
      __syncwarp();
      if (ln == 0) {
        const uint64_t p0 = (tile - sg->first_tile) * (uint64_t)kTilePts;
        const uint64_t rem = sg->n_points - p0;
        const uint32_t npts = rem < (uint64_t)kTilePts ? (uint32_t)rem : (uint32_t)kTilePts;
        const uint32_t bytes = (npts * (uint32_t)R + 15u) & ~15u;  // bulk copies move multiples of 16 bytes
        st_tile[s] = tile;
        st_segi[s] = seg;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_copy_g2s_hint(dsm + (size_t)s * kTileBytes, sg->rec + p0 * (uint64_t)R, bytes, &full_bar[s], stream_policy);
      }
    }
    return;
  }

  // ---------------- consumer warps ----------------
  GridQEntry* const wq = reinterpret_cast<GridQEntry*>(dsm + (size_t)STAGES * kTileBytes) + (QUEUE ? wid * Sh::kQCap : 0u);
  uint32_t qn = 0;  // queued matches of this warp (warp-uniform)
  uint32_t my_seg = 0xFFFFFFFFu;
  LaneChunk lch;
  Segment& S = wseg[wid];
  auto grid_of = [&]() -> const GridDev& { return ONE ? P.grid0 : P.lanes[S.lane].grid; };

  auto bind_lane = [&]() {
    if (lch.lane != S.lane) {  // a warp's chunk belongs to one collector's arena
      if (lch.lane != 0xFFFFFFFFu) cand_pad(ONE ? P.grid0 : P.lanes[lch.lane].grid, lch.c);
      lch.c = CandChunk();
      lch.lane = S.lane;
    }
  };
  // insert queued matches, newest first, kGsQFlush at a time (`all`: until the queue is empty)
  auto flush = [&](bool all) {
    if constexpr (QUEUE) {
      __syncwarp();
      while (qn >= (all ? 1u : kGsQFlush)) {
        const uint32_t take = qn < kGsQFlush ? qn : kGsQFlush;
        const uint32_t base = qn - take;
        bool m[kPPT];
        QueueFetch f;
#pragma unroll
        for (int j = 0; j < kPPT; ++j) {
          const uint32_t e = (uint32_t)j * 32u + ln;
          m[j] = e < take;
          f.qe[j] = wq + base + (m[j] ? e : 0u);
        }
        bind_lane();
        grid_insert_tile(grid_of(), S, m, f, lch.c);
        __syncwarp();
        qn = base;
      }
    }
  };

  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it % STAGES;
    mbar_wait(&full_bar[s], (it / STAGES) & 1u);
    const unsigned long long tile = st_tile[s];
    if (tile == ~0ull) break;
    const uint32_t seg_now = st_segi[s];
    if (seg_now != my_seg) {
      if (my_seg != 0xFFFFFFFFu) flush(true);  // queued points belong to the old segment
      const uint32_t* srcw = reinterpret_cast<const uint32_t*>(&st_seg[s]);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&S);
      for (uint32_t k = ln; k < sizeof(Segment) / 4; k += 32u) dstw[k] = srcw[k];
      my_seg = seg_now;
      __syncwarp();
    }
    const uint64_t p0 = (tile - S.first_tile) * (uint64_t)kTilePts;
    const uint64_t rem = S.n_points - p0;
    const uint32_t npts = rem < (uint64_t)kTilePts ? (uint32_t)rem : (uint32_t)kTilePts;
    const UnitSrc<R> src{dsm + (size_t)s * kTileBytes};
    const bool las = S.layout == PCQ_LAYOUT_LAS;
    RegFetch<PPT> f;
    bool m[PPT];
    f.gi0 = S.scan_base + p0 + wid * (uint32_t)Sh::kWarpPts + ln;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const uint32_t i = wid * (uint32_t)Sh::kWarpPts + (uint32_t)j * 32u + ln;
      int32_t x = 0, y = 0, z = 0;
      uint32_t cls = 0, rg = 0, bl = 0;
      m[j] = false;
      if (i < npts) {
        src.xyz(i, x, y, z);
        // (LAST: only bounds queries come here; their class and colour columns are read from global memory)
        cls = las ? (uint32_t)src.unit[i * (uint32_t)R + S.cls_off] : 0u;
        if (P.query_kind == PCQ_QUERY_BOUNDS)
          m[j] = in_range(x, S.lo[0], S.hi[0]) & in_range(y, S.lo[1], S.hi[1]) & in_range(z, S.lo[2], S.hi[2]);
        else
          m[j] = cls == P.cls;
        if (m[j]) {
          if (las) {
            if (S.rgb_off >= 0) src.rgb(i * (uint32_t)R + (uint32_t)S.rgb_off, rg, bl);
          } else {
            cls = (uint32_t)__ldg(S.cls + p0 + i);
            if (S.rgb != nullptr) {
              const uint8_t* p = S.rgb + (p0 + i) * 6ull;
              rg = ldg_u16(p) | (ldg_u16(p + 2) << 16);
              bl = ldg_u16(p + 4);
            }
          }
        }
      }
      f.x[j] = x;
      f.y[j] = y;
      f.z[j] = z;
      f.rg[j] = rg;
      f.bc[j] = (bl & 0xFFFFu) | ((cls & 0xFFu) << 16);
    }
    __syncwarp();
    if (ln == 0) mbar_arrive(&empty_bar[s]);  // the stage may be refilled: everything this warp needs is in registers

    if constexpr (QUEUE) {
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const uint32_t bal = __ballot_sync(0xffffffffu, m[j]);
        if (m[j]) {
          GridQEntry& e = wq[qn + (uint32_t)__popc(bal & ((1u << ln) - 1u))];
          e.x = f.x[j];
          e.y = f.y[j];
          e.z = f.z[j];
          e.rg = f.rg[j];
          e.bc = f.bc[j];
          e.gidx = f.gidx(S, j);
        }
        qn += (uint32_t)__popc(bal);
      }
      flush(false);
    } else {
      bind_lane();
      grid_insert_tile(grid_of(), S, m, f, lch.c);
    }
  }
  if (my_seg != 0xFFFFFFFFu) flush(true);
  if (lch.lane != 0xFFFFFFFFu) cand_pad(ONE ? P.grid0 : P.lanes[lch.lane].grid, lch.c);
}

// ------------------------------------------------------------------------------------------------
// MODE_SELECT — BufferCollector::collect_one in scan order (collect_points.rs:29-31): single-pass
// stable stream compaction with a decoupled look-back prefix.
//
// A *unit* is kSelUnitPts consecutive records of one segment.  A CTA is kSelWarps consumer warps plus
// one look-back warp; there is no CTA-wide barrier, the warps meet on mbarriers only:
//   look-back warp   takes the unit tickets (two units ahead), copies the unit's Segment into shared
//                    memory, and — once the consumers have counted a unit — publishes its aggregate,
//                    resolves the exclusive prefix over earlier units (256 descriptors per round trip)
//                    and hands the unit's first output slot to the consumers.
//   consumer warp w  owns records [256 w, 256 w + 256) of every unit, as 8 rows of 32 consecutive
//                    records (8 independent loads per lane in flight).  Per unit: evaluate the
//                    predicate, keep the 8 ballot masks, post the warp count; then emit the PREVIOUS
//                    unit (whose look-back ran while this unit's loads were in flight): matching lanes
//                    re-read their record (L2), compose the 31-byte Point in a warp-private staging
//                    buffer that has the 16-byte phase of its destination, and the warp flushes it
//                    with aligned 16-byte stores (loose bytes at both ends as byte stores).
// Tickets (not blockIdx) order the units, so a unit only ever waits for units held by running CTAs.
// At 6.5 TB/s a 2048-record unit lasts 4-9 ns while one L2 round trip is ~500 ns: the 256-wide
// look-back window is what keeps the chain of unresolved predecessors from limiting throughput
// (bound = 256 units * unit bytes / round trip, > 12 TB/s for every record length).
// ------------------------------------------------------------------------------------------------
constexpr int kSelWarps = 8;
constexpr int kSelRows = 8;
constexpr int kSelWarpPts = 32 * kSelRows;
constexpr int kSelUnitPts = kSelWarps * kSelWarpPts;  // 2048
constexpr int kSelLbWarps = 1;      // look-back warps per CTA (unit n is resolved by warp n % kSelLbWarps)
constexpr int kSelThreads = (kSelWarps + kSelLbWarps + 1) * 32;  // + look-back warps + dispatcher warp
constexpr int kSelLag = 3;         // a unit is emitted this many units behind the counted front
constexpr int kSelBufs = 8;        // unit descriptors per CTA: n-2 (emit) .. n+1 (loads) + two ticketed ahead
constexpr int kSelStageRecs = 64;  // records composed per flush round of a warp
constexpr int kSelStageBytes = 2048;  // 64 * 31 + 15 phase bytes, rounded up

struct SelUnit {
  Segment seg;
  unsigned long long tile;      // ~0 = no more units
  unsigned long long u0;        // first record of the unit within its segment
  unsigned long long out_rec;   // lane.out_base + exclusive prefix (set by the look-back warp)
  unsigned long long out_base;
  unsigned long long out_cap;
  uint8_t* out;
  unsigned long long* count;
  uint32_t npts;
  uint32_t acc;  // consumers: sum of posted warp counts | number of posted warps << 24
  uint32_t warp_cnt[16];  // one per consumer warp (8 in k_select / k_select_bytes, 16 in k_select_ring)
};

// L2 cache policies.  The select kernel reads every record once from HBM (`keep`: evict_last) and re-reads the
// matching ones from L2 a few microseconds later when it emits them (`drop`: evict_first, like the output stores),
// so that neither the re-read nor the 31-byte output stream pushes not-yet-emitted records out of L2.
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_drop() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t ldg_u32_h(const uint8_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u16_h(const uint8_t* p, uint64_t pol) {
  unsigned short v;
  asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
  return (uint32_t)v;
}
__device__ __forceinline__ uint32_t ldg_u8_h(const uint8_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 ldg_v4_h(const uint8_t* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_v4_h(uint4* p, const uint4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ void stg_u8_h(uint8_t* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// loads of little-endian fields whose alignment AL (4, 2 or 1) is a compile-time constant
template <int AL>
__device__ __forceinline__ int32_t ldg_i32_a(const uint8_t* p, uint64_t pol) {
  if constexpr (AL == 4) {
    return (int32_t)ldg_u32_h(p, pol);
  } else if constexpr (AL == 2) {
    return (int32_t)(ldg_u16_h(p, pol) | (ldg_u16_h(p + 2, pol) << 16));
  } else {
    return (int32_t)(ldg_u8_h(p, pol) | (ldg_u8_h(p + 1, pol) << 8) | (ldg_u8_h(p + 2, pol) << 16) | (ldg_u8_h(p + 3, pol) << 24));
  }
}
template <int AL>
__device__ __forceinline__ uint32_t ldg_u16_a(const uint8_t* p, uint64_t pol) {
  if constexpr (AL >= 2) {
    return ldg_u16_h(p, pol);
  } else {
    return ldg_u8_h(p, pol) | (ldg_u8_h(p + 1, pol) << 8);
  }
}

struct RawPoint {
  int32_t x, y, z;
  uint32_t cls, r, g, b;
};

// all fields of two records at once (14 independent loads per lane in flight)
template <int AL>
__device__ __forceinline__ void select_fetch2(const Segment& S, uint64_t i0, uint64_t i1, RawPoint& a, RawPoint& b,
                                              uint64_t pol) {
  const uint8_t* p0 = S.rec + i0 * (uint64_t)S.record_len;  // LAST: record_len == 12 (positions column)
  const uint8_t* p1 = S.rec + i1 * (uint64_t)S.record_len;
  a.x = ldg_i32_a<AL>(p0, pol);
  a.y = ldg_i32_a<AL>(p0 + 4, pol);
  a.z = ldg_i32_a<AL>(p0 + 8, pol);
  b.x = ldg_i32_a<AL>(p1, pol);
  b.y = ldg_i32_a<AL>(p1 + 4, pol);
  b.z = ldg_i32_a<AL>(p1 + 8, pol);
  a.r = a.g = a.b = b.r = b.g = b.b = 0u;  // Vector3::new(0, 0, 0), las.rs:134
  if (S.layout == PCQ_LAYOUT_LAS) {
    a.cls = ldg_u8_h(p0 + S.cls_off, pol);
    b.cls = ldg_u8_h(p1 + S.cls_off, pol);
    if (S.rgb_off >= 0) {
      const uint8_t* c0 = p0 + (uint32_t)S.rgb_off;  // 20 / 28: as aligned as the record itself (up to 2)
      const uint8_t* c1 = p1 + (uint32_t)S.rgb_off;
      a.r = ldg_u16_a<AL>(c0, pol);
      a.g = ldg_u16_a<AL>(c0 + 2, pol);
      a.b = ldg_u16_a<AL>(c0 + 4, pol);
      b.r = ldg_u16_a<AL>(c1, pol);
      b.g = ldg_u16_a<AL>(c1 + 2, pol);
      b.b = ldg_u16_a<AL>(c1 + 4, pol);
    }
  } else {
    a.cls = ldg_u8_h(S.cls + i0, pol);
    b.cls = ldg_u8_h(S.cls + i1, pol);
    if (S.rgb != nullptr) {
      const uint8_t* c0 = S.rgb + i0 * 6ull;
      const uint8_t* c1 = S.rgb + i1 * 6ull;
      if (S.rgb_align2) {
        a.r = ldg_u16_a<2>(c0, pol);
        a.g = ldg_u16_a<2>(c0 + 2, pol);
        a.b = ldg_u16_a<2>(c0 + 4, pol);
        b.r = ldg_u16_a<2>(c1, pol);
        b.g = ldg_u16_a<2>(c1 + 2, pol);
        b.b = ldg_u16_a<2>(c1 + 4, pol);
      } else {
        a.r = ldg_u16_a<1>(c0, pol);
        a.g = ldg_u16_a<1>(c0 + 2, pol);
        a.b = ldg_u16_a<1>(c0 + 4, pol);
        b.r = ldg_u16_a<1>(c1, pol);
        b.g = ldg_u16_a<1>(c1 + 2, pol);
        b.b = ldg_u16_a<1>(c1 + 4, pol);
      }
    }
  }
}

__device__ __forceinline__ void select_compose(const Segment& S, const RawPoint& q, uint8_t* dst) {
  Hit h;
  h.x = q.x;
  h.y = q.y;
  h.z = q.z;
  h.cls = q.cls;
  const uint32_t rgb[3] = {q.r, q.g, q.b};
  uint32_t wd[8];
  point_words(S, h, rgb, wd);
  sts_point31(dst, wd);
}

// Copy `nb` staged bytes (whole records, so at least 31) to their place in the output: `stage` is the image of the
// output stream with the same 16-byte phase as global memory, so the body moves as aligned 16-byte chunks, lane l taking
// chunks l, l + 32, ... (ITERS * 32 of them at most), and the ragged ends as single bytes.
template <int ITERS>
__device__ __forceinline__ void select_flush(uint8_t* out, const uint8_t* stage, unsigned long long g0, uint32_t phase,
                                             uint32_t nb, uint32_t ln, uint64_t pol) {
  uint8_t* gbase = out + (g0 - phase);  // 16-byte aligned; byte k of `stage` belongs at gbase[k]
  const uint32_t end = phase + nb;
  const uint32_t c_first = (phase + 15u) >> 4, c_end = end >> 4;  // full 16-byte chunks [c_first, c_end)
  if (phase + ln < (c_first << 4)) stg_u8_h(gbase + phase + ln, stage[phase + ln], pol);
  const uint32_t c0 = c_first + ln;
  const uint4* sv = reinterpret_cast<const uint4*>(stage) + c0;
  uint4* gv = reinterpret_cast<uint4*>(gbase) + c0;
#pragma unroll
  for (int i = 0; i < ITERS; ++i)
    if (c0 + 32u * (uint32_t)i < c_end) stg_v4_h(gv + 32 * i, sv[32 * i], pol);
  const uint32_t tb = (c_end << 4) + ln;
  if (tb < end) stg_u8_h(gbase + tb, stage[tb], pol);
}

// Emit the matches of consumer warp `w` in unit U (all 32 lanes call it).  index_of(r) is the index, within the
// warp's `warp_pts` consecutive records of the unit, of the warp's r-th match.  Dense: lane l composes matches l and
// l + 32 of each round of 64.
// [r_begin, r_begin + r_count) restricts the call to a range of the warp's matches (index_of takes the full rank).
template <int AL, class IndexOf>
__device__ __forceinline__ void select_emit_warp(const SelUnit& U, const IndexOf& index_of_full, uint32_t warp_pts, uint8_t* stage,
                                                 uint32_t r_begin = 0u, uint32_t r_count = 0xFFFFFFFFu) {
  const uint32_t w = warp_id(), ln = lane_id();
  const uint32_t mine = r_count == 0xFFFFFFFFu ? U.warp_cnt[w] : r_count;
  if (mine == 0) return;
  uint32_t before = r_begin;
  for (uint32_t k = 0; k < w; ++k) before += U.warp_cnt[k];
  const unsigned long long out0 = U.out_rec + before;
  auto index_of = [&](uint32_t r) -> uint32_t { return index_of_full(r_begin + r); };
  const Segment& S = U.seg;
  const uint64_t wbase = U.u0 + (uint64_t)w * warp_pts;
  const uint64_t pol = l2_policy_drop();
  // the fields of round r + 1 are fetched (L2) while round r is composed and flushed
  RawPoint q0, q1;
  {
    const uint32_t n = min((uint32_t)kSelStageRecs, mine);
    // lanes beyond the round's records fetch the round's first record again (always a valid address)
    select_fetch2<AL>(S, wbase + index_of(ln < n ? ln : 0u), wbase + index_of(ln + 32u < n ? ln + 32u : 0u), q0, q1, pol);
  }
#pragma unroll 1
  for (uint32_t base = 0; base < mine; base += kSelStageRecs) {
    const unsigned long long orec = out0 + base;
    const unsigned long long g0 = orec * 31ull;      // first output byte of this round
    const uint32_t phase = (uint32_t)(g0 & 15ull);   // keep global and shared 16-byte phases equal
    const uint32_t n = min((uint32_t)kSelStageRecs, mine - base);
    const uint32_t r0 = ln, r1 = ln + 32u;
    RawPoint p0 = q0, p1 = q1;
    if (base + kSelStageRecs < mine) {
      const uint32_t nbase = base + kSelStageRecs;
      const uint32_t nn = min((uint32_t)kSelStageRecs, mine - nbase);
      select_fetch2<AL>(S, wbase + index_of(nbase + (r0 < nn ? r0 : 0u)), wbase + index_of(nbase + (r1 < nn ? r1 : 0u)), q0, q1, pol);
    }
    if (r0 < n) select_compose(S, p0, stage + phase + r0 * 31u);
    if (r1 < n) select_compose(S, p1, stage + phase + r1 * 31u);
    __syncwarp();
    // records beyond the lane's capacity are counted but not written (host grows the buffer and re-runs)
    const unsigned long long room = orec < U.out_cap ? U.out_cap - orec : 0ull;
    const uint32_t nb = (room < (unsigned long long)n ? (uint32_t)room : n) * 31u;
    if (nb) select_flush<(kSelStageRecs * 31 + 15 + 511) / 512>(U.out, stage, g0, phase, nb, ln, pol);
    __syncwarp();  // the staging buffer is reused by the next round / unit
  }
}

// ---------------- dispatcher state shared by the select kernels ----------------
// The dispatcher warp hands a CTA its units.  Every unit used to cost it a chain of dependent round trips to L2 — the
// ticket atomic, the first tile of the next segment, the segment itself, the collector's buffer.  The atomic of the
// NEXT ticket is now in flight while the current unit is set up, and the segment and its collector are cached (shared
// memory / lane 0's registers) until a unit crosses into another segment.
// Tickets are taken ONE at a time: with batches of two (measured) a CTA counts unit t + 1 a whole unit time after
// unit t, every look-back that crosses such a pair waits for it, and all selects lost 10-25 %.
constexpr unsigned long long kSelTicketBatch = 1ull;

struct DispatchCache {
  Segment seg;                    // the segment of the last unit
  unsigned long long next_first;  // first tile of the following segment (~0: none)
};

struct Dispatcher {
  unsigned long long cur = 0, end = 0;  // tickets in hand: [cur, end)
  unsigned long long ahead = 0;         // lane 0: base of the batch whose atomic is in flight
  uint32_t seg_cur = 0xFFFFFFFFu;
  // lane 0: the collector of the cached segment
  uint8_t* out = nullptr;
  unsigned long long out_cap = 0, out_base = 0;
  unsigned long long* count = nullptr;

  __device__ __forceinline__ void start(const ScanParams& P) {
    if (lane_id() == 0) ahead = atomicAdd(P.ticket, kSelTicketBatch);
  }
  __device__ __forceinline__ unsigned long long next_tile(const ScanParams& P) {
    if (cur == end) {
      cur = __shfl_sync(0xffffffffu, ahead, 0);
      end = cur + kSelTicketBatch;
      if (lane_id() == 0 && cur < P.n_tiles) ahead = atomicAdd(P.ticket, kSelTicketBatch);
    }
    return cur++;
  }
  // makes `dc` describe the segment of `tile` (the tiles of a CTA only move forward)
  __device__ __forceinline__ void locate(const ScanParams& P, DispatchCache& dc, unsigned long long tile) {
    if (seg_cur != 0xFFFFFFFFu && tile < dc.next_first) return;
    seg_cur = seg_forward(P, seg_cur == 0xFFFFFFFFu ? 0u : seg_cur, tile);
    const Segment* sg = P.segs + seg_cur;
    const uint32_t* srcw = reinterpret_cast<const uint32_t*>(sg);
    uint32_t* dstw = reinterpret_cast<uint32_t*>(&dc.seg);
    __syncwarp();
    for (uint32_t k = lane_id(); k < sizeof(Segment) / 4; k += 32u) dstw[k] = srcw[k];
    if (lane_id() == 0) {
      dc.next_first = seg_cur + 1 < P.n_segs ? P.segs[seg_cur + 1].first_tile : ~0ull;
      const LaneDev* L = P.lanes + sg->lane;
      out = L->out;
      out_cap = L->out_cap;
      out_base = L->out_base;
      count = L->count;
    }
    __syncwarp();
  }
  // fills the unit descriptor (all lanes copy the segment, lane 0 the scalars); returns the unit's point count
  template <int UNIT_PTS>
  __device__ __forceinline__ uint32_t describe(const DispatchCache& dc, SelUnit& U, unsigned long long tile) {
    const uint32_t* srcw = reinterpret_cast<const uint32_t*>(&dc.seg);
    uint32_t* dstw = reinterpret_cast<uint32_t*>(&U.seg);
    for (uint32_t k = lane_id(); k < sizeof(Segment) / 4; k += 32u) dstw[k] = srcw[k];
    const uint64_t u0 = (tile - dc.seg.first_tile) * (uint64_t)UNIT_PTS;
    const uint64_t rem = dc.seg.n_points - u0;
    const uint32_t npts = rem < (uint64_t)UNIT_PTS ? (uint32_t)rem : (uint32_t)UNIT_PTS;
    if (lane_id() == 0) {
      U.tile = tile;
      U.u0 = u0;
      U.npts = npts;
      U.out = out;
      U.out_cap = out_cap;
      U.out_base = out_base;
      U.count = count;
    }
    return npts;
  }
};

// ---------------- dispatcher warp: tickets and unit descriptors, never blocked by a look-back ----------------
template <int UNIT_PTS>
__device__ __forceinline__ void select_dispatcher_warp(const ScanParams& P, SelUnit* unit, uint64_t* bar_tk, uint64_t* bar_free,
                                                       DispatchCache& dc) {
  const uint32_t ln = lane_id();
  Dispatcher D;
  D.start(P);
  uint32_t end_marks = 0;  // every look-back warp needs to meet an end marker
  for (uint32_t n = 0;; ++n) {
    const uint32_t b = n % kSelBufs;
    if (n >= (uint32_t)kSelBufs) mbar_wait(&bar_free[b], (n / kSelBufs - 1u) & 1u);
    unsigned long long tile = ~0ull;
    if (end_marks == 0u) tile = D.next_tile(P);
    SelUnit& U = unit[b];
    if (ln == 0) U.acc = 0u;
    if (tile >= P.n_tiles) {
      if (ln == 0) {
        U.tile = ~0ull;
        mbar_arrive(&bar_tk[b]);
      }
      if (++end_marks == (uint32_t)kSelLbWarps) break;
      continue;
    }
    D.locate(P, dc, tile);
    D.template describe<UNIT_PTS>(dc, U, tile);
    __syncwarp();
    if (ln == 0) mbar_arrive(&bar_tk[b]);
  }
}

// ---------------- look-back warp(s): unit n is resolved by look-back warp n % kSelLbWarps ----------------
template <int NW = kSelWarps>
__device__ __forceinline__ void select_lookback_warp(const ScanParams& P, SelUnit* unit, uint64_t* bar_tk, uint64_t* bar_cnt,
                                                     uint64_t* bar_pre) {
  const uint32_t ln = lane_id();
  for (uint32_t n = warp_id() - NW;; n += kSelLbWarps) {
    const uint32_t b = n % kSelBufs;
    const uint32_t par = (n / kSelBufs) & 1u;
    mbar_wait(&bar_tk[b], par);
    SelUnit& U = unit[b];
    const unsigned long long tile = U.tile;
    if (tile == ~0ull) break;
    const uint64_t lane_first = U.seg.lane_first_tile;
    mbar_wait(&bar_cnt[b], par);
    uint32_t total = ln < (uint32_t)NW ? U.warp_cnt[ln] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    // (the unit's aggregate was published by the consumer warp that posted the last count)
    if (total == 0u && tile != lane_first && !PCQ_HOOK(P, 1u)) {
      // Nothing to emit, so the unit needs no output slot and nobody has to wait for its prefix: successors look
      // straight through its (zero) aggregate.  A query without matches would otherwise run at the rate at which the
      // GPU resolves look-backs instead of at memory speed.  So that later units do not have to walk back over long
      // runs of such units, the prefix is still published when the predecessor already carries one (one read).
      if (ln == 0) {
        const unsigned long long pv = ld_state(P.tile_state + (tile - 1) * kDescStride);
        if ((pv >> kStatusShift) == kStPrefix) st_state(P.tile_state + tile * kDescStride, pv);  // same prefix: nothing added
        U.out_rec = U.out_base;
        mbar_arrive(&bar_pre[b]);
      }
      continue;
    }
    unsigned long long excl = 0;
    if (tile != lane_first && !PCQ_HOOK(P, 1u)) excl = lookback_exclusive(P.tile_state, tile, lane_first);
    if (ln == 0) {
      if (tile != lane_first) st_state(P.tile_state + tile * kDescStride, (kStPrefix << kStatusShift) | (excl + (unsigned long long)total));
      U.out_rec = U.out_base + excl;
      if (total) atomicAdd(U.count, (unsigned long long)total);
      mbar_arrive(&bar_pre[b]);
    }
  }
}

// a consumer warp posts its match count of unit U; the warp that posts the last count publishes the unit's aggregate
// right away, so that no look-back of another CTA ever waits behind one of this CTA's look-backs (lane 0 only)
template <int NW = kSelWarps>
__device__ __forceinline__ void select_post_count(const ScanParams& P, SelUnit& U, uint32_t cnt, uint64_t* bar_cnt_b) {
  U.warp_cnt[warp_id()] = cnt;
  const uint32_t old = atomicAdd(&U.acc, cnt + (1u << 24));
  if ((old >> 24) == (uint32_t)NW - 1u) {
    const unsigned long long total = (unsigned long long)((old & 0xFFFFFFu) + cnt);
    const unsigned long long st = U.tile == U.seg.lane_first_tile ? kStPrefix : kStAgg;  // first unit: prefix == aggregate
    st_state(P.tile_state + U.tile * kDescStride, (st << kStatusShift) | total);
  }
  mbar_arrive(bar_cnt_b);
}

// the barriers of a select CTA
template <int NW = kSelWarps>
__device__ __forceinline__ void select_init_barriers(uint64_t* bar_tk, uint64_t* bar_cnt, uint64_t* bar_pre, uint64_t* bar_free) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < kSelBufs; ++b) {
      mbar_init(&bar_tk[b], 1u);
      mbar_init(&bar_cnt[b], (uint32_t)NW);
      mbar_init(&bar_pre[b], 1u);
      mbar_init(&bar_free[b], (uint32_t)NW);
    }
    mbar_fence_init();
  }
  __syncthreads();
}

template <int AL>
__global__ void __launch_bounds__(kSelThreads, 2) k_select(ScanParams P) {
  __shared__ SelUnit unit[kSelBufs];
  __shared__ __align__(8) uint64_t bar_tk[kSelBufs];    // unit descriptor filled    (dispatcher -> consumers, look-back warp)
  __shared__ __align__(8) uint64_t bar_cnt[kSelBufs];   // all warp counts posted    (consumers -> look-back warp)
  __shared__ __align__(8) uint64_t bar_pre[kSelBufs];   // exclusive prefix resolved (look-back warp -> consumers)
  __shared__ __align__(8) uint64_t bar_free[kSelBufs];  // unit emitted              (consumers -> dispatcher)
  __shared__ DispatchCache dcache;
  __shared__ __align__(16) uint8_t stage[kSelWarps][kSelStageBytes];
  __shared__ uint16_t match_list[kSelWarps][kSelLag + 1][kSelWarpPts];  // unit-local index of each warp's r-th match

  const uint32_t ln = lane_id();
  select_init_barriers(bar_tk, bar_cnt, bar_pre, bar_free);

  if (warp_id() == kSelWarps + kSelLbWarps) {
    select_dispatcher_warp<kSelUnitPts>(P, unit, bar_tk, bar_free, dcache);
    return;
  }
  if (warp_id() >= kSelWarps) {
    select_lookback_warp(P, unit, bar_tk, bar_cnt, bar_pre);
    return;
  }

  // ---------------- consumer warps ----------------
  // iteration n: ballot unit n (its loads were issued one iteration ago), record the matches, post the count
  // (the last warp to post publishes the aggregate); issue the predicate loads of unit n + 1; emit unit n - 2
  // while they are in flight.  Counting never waits for a look-back; only the emit does, two units later.
  const uint32_t w = warp_id();
  const uint32_t lt = (1u << ln) - 1u;
  int32_t vx[kSelRows], vy[kSelRows], vz[kSelRows];
  uint32_t npts = 0;

  const uint64_t pol_keep = l2_policy_keep();
  auto issue_loads = [&](const SelUnit& U) {
    const Segment& S = U.seg;
    npts = U.npts;
    const uint64_t u0 = U.u0;
    if (P.query_kind == PCQ_QUERY_BOUNDS) {
      const uint64_t stride = S.record_len;  // LAST: 12
#pragma unroll
      for (int k = 0; k < kSelRows; ++k) {
        const uint32_t i = min(w * kSelWarpPts + (uint32_t)k * 32u + ln, npts - 1u);  // clamped: always loadable
        const uint8_t* p = S.rec + (u0 + i) * stride;
        vx[k] = ldg_i32_a<AL>(p, pol_keep);
        vy[k] = ldg_i32_a<AL>(p + 4, pol_keep);
        vz[k] = ldg_i32_a<AL>(p + 8, pol_keep);
      }
    } else {
      const bool las = S.layout == PCQ_LAYOUT_LAS;
      const uint8_t* cbase = las ? S.rec + S.cls_off : S.cls;
      const uint64_t stride = las ? (uint64_t)S.record_len : 1ull;
#pragma unroll
      for (int k = 0; k < kSelRows; ++k) {
        const uint32_t i = min(w * kSelWarpPts + (uint32_t)k * 32u + ln, npts - 1u);
        vx[k] = (int32_t)ldg_u8_h(cbase + (u0 + i) * stride, pol_keep);
        vy[k] = vz[k] = 0;
      }
    }
  };

  mbar_wait(&bar_tk[0], 0u);
  bool cur = unit[0].tile != ~0ull;
  if (cur) issue_loads(unit[0]);
  uint32_t emit_next = 0;  // oldest unit of this CTA that this warp has not emitted yet
  for (uint32_t n = 0;; ++n) {
    const uint32_t b = n % kSelBufs;
    if (cur) {
      SelUnit& U = unit[b];
      const Segment& S = U.seg;
      uint16_t* list = match_list[w][n % (uint32_t)(kSelLag + 1)];
      uint32_t cnt = 0;
#pragma unroll
      for (int k = 0; k < kSelRows; ++k) {
        const uint32_t il = (uint32_t)k * 32u + ln;  // index within the warp's 256 records
        bool m = w * kSelWarpPts + il < npts;
        if (P.query_kind == PCQ_QUERY_BOUNDS)
          m = m & in_range(vx[k], S.lo[0], S.hi[0]) & in_range(vy[k], S.lo[1], S.hi[1]) & in_range(vz[k], S.lo[2], S.hi[2]);
        else
          m = m & ((uint32_t)vx[k] == P.cls);
        const uint32_t bal = __ballot_sync(0xffffffffu, m);
        if (m) list[cnt + (uint32_t)__popc(bal & lt)] = (uint16_t)il;
        cnt += (uint32_t)__popc(bal);
      }
      __syncwarp();  // the list is read by other lanes when the unit is emitted
      if (ln == 0) select_post_count(P, U, cnt, &bar_cnt[b]);
    }
    const uint32_t counted = cur ? n + 1u : n;  // units [0, counted) of this CTA are counted
    bool nxt = false;
    if (cur) {
      const uint32_t nb = (n + 1u) % kSelBufs;
      mbar_wait(&bar_tk[nb], ((n + 1u) / kSelBufs) & 1u);
      nxt = unit[nb].tile != ~0ull;
      if (nxt) issue_loads(unit[nb]);
    }
    // Emit while the loads of unit n + 1 are in flight: units two or more behind the counted front must go now
    // (their match list and descriptor are about to be reused); a younger unit goes early when its prefix is
    // already there, which keeps the records it re-reads resident in L2.  After the last unit everything goes.
    while (emit_next < counted) {
      const uint32_t pb = emit_next % kSelBufs;
      const uint32_t par = (emit_next / kSelBufs) & 1u;
      const bool must = !nxt || emit_next + (uint32_t)kSelLag < counted;
      if (must) {
        mbar_wait(&bar_pre[pb], par);
      } else if (!PCQ_HOOK(P, 4u) || !mbar_test(&bar_pre[pb], par)) {
        break;  // (debug 4: emit a younger unit early when its prefix is already there)
      }
      if (!PCQ_HOOK(P, 2u)) {
        const uint16_t* list = match_list[w][emit_next % (uint32_t)(kSelLag + 1)];
        select_emit_warp<AL>(unit[pb], [list](uint32_t r) { return (uint32_t)list[r]; }, (uint32_t)kSelWarpPts, stage[w]);
      }
      __syncwarp();
      if (ln == 0) mbar_arrive(&bar_free[pb]);
      ++emit_next;
    }
    if (!nxt) break;
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------
// k_select_ring<R>: k_select for launches whose segments all have record length R and 16-byte aligned ranges
// (LAS records, or LAST positions with R = 12, bounds queries and LAS class queries).
// In k_select the predicate fields of the next unit wait in REGISTERS (24 per thread), which bounds the bytes in
// flight per SM at ~114 KB — enough for a pure scan, not enough once consumers also wait for look-backs and emit.
// Here the dispatcher streams every unit global -> shared with one bulk async copy (cp.async.bulk + mbarrier
// complete_tx, the scan kernels' mechanism) into a ring of unit-sized slots, so 140-170 KB stay in flight whatever
// the consumers are doing; 16 consumer warps read the predicate from shared memory without bank conflicts and
// release the slot right after the ballots.  Units are 80-93 KB of records whatever the record length (the
// look-back machinery resolves a fixed number of units per microsecond GPU-wide, so bytes per unit are what turns
// that rate into bandwidth).  Emit, look-back and barriers are k_select's.
// ------------------------------------------------------------------------------------------------
constexpr int kSelRWarps = 16;
constexpr int kSelRThreads = (kSelRWarps + kSelLbWarps + 1) * 32;
constexpr int kSelRSlots = 2;  // unit-sized ring slots (the next unit lands while the current one is balloted)

template <int R>
struct SelRing {
  static constexpr int rows = (int)sel_ring_rows(R);           // rows of 32 records per consumer warp and unit
  static constexpr int warp_pts = 32 * rows;
  static constexpr int unit_pts = kSelRWarps * warp_pts;       // == sel_ring_unit_points(R): 80-93 KB of records
  static constexpr uint32_t slot_bytes = (uint32_t)unit_pts * (uint32_t)R;
  static constexpr size_t smem = (size_t)kSelRSlots * slot_bytes + (size_t)kSelRWarps * kSelStageBytes;
};

template <int R>
__global__ void __launch_bounds__(kSelRThreads, 1) k_select_ring(ScanParams P) {
  constexpr int AL = (R % 4 == 0) ? 4 : 2;  // a 16-byte aligned range of R-byte records
  constexpr int RS = kSelRSlots;
  constexpr int ROWS = SelRing<R>::rows;
  constexpr int WPTS = SelRing<R>::warp_pts;
  constexpr int UNIT = SelRing<R>::unit_pts;
  constexpr uint32_t kSlotBytes = SelRing<R>::slot_bytes;
  __shared__ SelUnit unit[kSelBufs];
  __shared__ __align__(8) uint64_t bar_tk[kSelBufs];
  __shared__ __align__(8) uint64_t bar_cnt[kSelBufs];
  __shared__ __align__(8) uint64_t bar_pre[kSelBufs];
  __shared__ __align__(8) uint64_t bar_free[kSelBufs];
  __shared__ DispatchCache dcache;
  __shared__ __align__(8) uint64_t bar_full[RS];   // the unit's records have landed   (bulk copy -> consumers)
  __shared__ __align__(8) uint64_t bar_sfree[RS];  // the slot has been balloted        (consumers -> dispatcher)
  // ballot mask of every (warp, row) of the units that are counted but not emitted yet: 60 bytes per warp and unit
  // replace an index list; the emit finds its r-th match with a scan over the row counts and a find-nth-set-bit
  __shared__ uint32_t m_bal[kSelRWarps][kSelLag + 1][ROWS];
  __shared__ uint16_t m_run[kSelRWarps][kSelLag + 1][ROWS];  // matches of the warp before row k
  extern __shared__ __align__(128) uint8_t sel_ring_dsm[];
  uint8_t* ring = sel_ring_dsm;
  uint8_t(*stage)[kSelStageBytes] = reinterpret_cast<uint8_t(*)[kSelStageBytes]>(sel_ring_dsm + (size_t)RS * kSlotBytes);

  const uint32_t ln = lane_id();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < RS; ++s) {
      mbar_init(&bar_full[s], 1u);
      mbar_init(&bar_sfree[s], (uint32_t)kSelRWarps);
    }
  }
  select_init_barriers<kSelRWarps>(bar_tk, bar_cnt, bar_pre, bar_free);  // (fences and synchronises the CTA)

  if (warp_id() == kSelRWarps + kSelLbWarps) {
    // ---------------- dispatcher warp: tickets, unit descriptors and the bulk copies ----------------
    Dispatcher D;
    D.start(P);
    for (uint32_t n = 0;; ++n) {
      const uint32_t b = n % kSelBufs;
      if (n >= (uint32_t)kSelBufs) mbar_wait(&bar_free[b], (n / kSelBufs - 1u) & 1u);
      const unsigned long long tile = D.next_tile(P);
      SelUnit& U = unit[b];
      if (ln == 0) U.acc = 0u;
      if (tile >= P.n_tiles) {
        if (ln == 0) {
          U.tile = ~0ull;
          mbar_arrive(&bar_tk[b]);
        }
        break;
      }
      D.locate(P, dcache, tile);
      const uint32_t npts = D.template describe<UNIT>(dcache, U, tile);
      const uint32_t slot = n % (uint32_t)RS;
      if (n >= (uint32_t)RS) mbar_wait(&bar_sfree[slot], (n / (uint32_t)RS - 1u) & 1u);
      if (ln == 0) {
        const uint64_t u0 = (tile - dcache.seg.first_tile) * (uint64_t)UNIT;
        const uint32_t bytes = (npts * (uint32_t)R + 15u) & ~15u;  // bulk copies move multiples of 16 bytes
        mbar_arrive_expect_tx(&bar_full[slot], bytes);
        bulk_copy_g2s(ring + (size_t)slot * kSlotBytes, dcache.seg.rec + u0 * (uint64_t)R, bytes, &bar_full[slot]);
      }
      __syncwarp();
      if (ln == 0) mbar_arrive(&bar_tk[b]);
    }
    return;
  }
  if (warp_id() >= kSelRWarps) {
    select_lookback_warp<kSelRWarps>(P, unit, bar_tk, bar_cnt, bar_pre);
    return;
  }

  // ---------------- consumer warps ----------------
  const uint32_t w = warp_id();
  uint32_t emit_next = 0, counted = 0;
  bool dense = false;  // more than three quarters of the warp's records of the latest unit matched
  auto emit_one = [&]() {
    const uint32_t pb = emit_next % kSelBufs;
    mbar_wait(&bar_pre[pb], (emit_next / kSelBufs) & 1u);
    if (!PCQ_HOOK(P, 2u)) {
      const uint32_t* bal = m_bal[w][emit_next % (uint32_t)(kSelLag + 1)];
      const uint16_t* run = m_run[w][emit_next % (uint32_t)(kSelLag + 1)];
      auto index_of = [bal, run](uint32_t r) -> uint32_t {
        uint32_t k = 0;  // last row whose first match is not beyond r
#pragma unroll
        for (int j = 1; j < ROWS; ++j) k += (uint32_t)run[j] <= r ? 1u : 0u;  // run[] is non-decreasing
        const uint32_t b = bal[k], n = r - (uint32_t)run[k];
        return k * 32u + (b == 0xFFFFFFFFu ? n : nth_set_bit(b, n));  // a full row needs no search
      };
      select_emit_warp<AL>(unit[pb], index_of, (uint32_t)WPTS, stage[w]);
    }
    __syncwarp();
    if (ln == 0) mbar_arrive(&bar_free[pb]);
    ++emit_next;
  };
  for (uint32_t n = 0;; ++n) {
    const uint32_t b = n % kSelBufs;
    mbar_wait(&bar_tk[b], (n / kSelBufs) & 1u);
    SelUnit& U = unit[b];
    if (U.tile == ~0ull) break;
    const uint32_t slot = n % (uint32_t)RS;
    mbar_wait(&bar_full[slot], (n / (uint32_t)RS) & 1u);
    {
      const Segment& S = U.seg;
      const uint32_t npts = U.npts;
      const uint8_t* rec = ring + (size_t)slot * kSlotBytes + (size_t)(w * WPTS) * R;
      uint32_t* bal_out = m_bal[w][n % (uint32_t)(kSelLag + 1)];
      uint16_t* run_out = m_run[w][n % (uint32_t)(kSelLag + 1)];
      uint32_t cnt = 0;
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        const uint32_t il = (uint32_t)k * 32u + ln;  // index within the warp's records
        const uint8_t* p = rec + il * R;
        bool m = w * WPTS + il < npts;  // (beyond the unit's end the slot holds stale bytes)
        if (P.query_kind == PCQ_QUERY_BOUNDS) {
          const int32_t x = SmemSrc<R>::lds_i32(p), y = SmemSrc<R>::lds_i32(p + 4), z = SmemSrc<R>::lds_i32(p + 8);
          m = m & in_range(x, S.lo[0], S.hi[0]) & in_range(y, S.lo[1], S.hi[1]) & in_range(z, S.lo[2], S.hi[2]);
        } else {
          m = m & ((uint32_t)p[S.cls_off] == P.cls);  // LAS records only (LAST class queries never come here)
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, m);
        if (ln == 0) {
          bal_out[k] = bal;
          run_out[k] = (uint16_t)cnt;
        }
        cnt += (uint32_t)__popc(bal);
      }
      __syncwarp();  // the masks are read by every lane when the unit is emitted; the slot is no longer needed
      if (ln == 0) {
        mbar_arrive(&bar_sfree[slot]);
        select_post_count<kSelRWarps>(P, U, cnt, &bar_cnt[b]);
      }
      dense = cnt * 4u > 3u * (uint32_t)WPTS;
    }
    counted = n + 1u;
    // Dense matches: the emit is what the warp spends its time on, so look-backs resolve in its shadow anyway —
    // emit close behind the front, while the matching records are still in L2.  Sparse matches: stay three units
    // behind, so that the warp never waits for a look-back.
    const uint32_t lag = dense ? 1u : (uint32_t)kSelLag;
    while (emit_next + lag < counted) emit_one();
  }
  while (emit_next < counted) emit_one();
}

// ------------------------------------------------------------------------------------------------
// k_select_bytes, dense units of colourless LAST files: the emit is a gather of 12-byte positions, and with the
// register path a warp has 64 + 64 records in flight, which at DRAM latency is ~55 G records/s for the GPU.  Here the
// positions of 128 matches per round are gathered with cp.async into shared memory, one round ahead: up to 256 records
// in flight per warp and none of them held in registers.  (The class byte of a match is the query's class.)
// ------------------------------------------------------------------------------------------------
constexpr int kSelBRows = 8;
constexpr int kSelBWarpPts = kSelBRows * 32 * 16;       // 4096
constexpr int kSelBUnitPts = kSelWarps * kSelBWarpPts;  // 32768
constexpr int kSelBLag = 2;
constexpr int kSelBDenseRecs = 128;
constexpr int kSelBStageBytes = 4096;                                    // 128 * 31 + 15 phase bytes, rounded up
constexpr int kSelBGatherBytes = kSelBDenseRecs * 12;                    // positions of one round
constexpr int kSelBChunkRecs = 2 * kSelBGatherBytes / 32;                // sparse warp-units: records per round of 32-byte slots (96)
#ifndef PCQ_SELB_CHUNKS
#define PCQ_SELB_CHUNKS 1
#endif
#ifndef PCQ_SELB_CLASS_POLICY
#define PCQ_SELB_CLASS_POLICY l2_policy_drop()
#endif
#ifndef PCQ_SELB_SEARCH_MAX
#define PCQ_SELB_SEARCH_MAX 512
#endif
constexpr int kSelBSearchMax = PCQ_SELB_SEARCH_MAX;                                      // up to this many matches per warp-unit: binary search, no list
constexpr int kSelBListPts = 512;                                        // matches per part of the emit (one full row always fits)
// (1 KB of list per warp keeps the CTA at 92 KB: two CTAs per SM leave the L1 the 60 KB it had — a 2 KB list pushed the
// carve-out to 228 KB and the sparse gathers, which live on L1 hits for the second and third word of a position,
// went from 0.45 to 0.61 ms on 326 M points)
constexpr int kSelBWarpSmem = kSelBStageBytes + 2 * kSelBGatherBytes + kSelBListPts * 2;  // per consumer warp (dynamic)
static_assert(kSelBStageBytes >= kSelStageBytes, "the register path stages through the same buffer");

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <class IndexOf>
__device__ __forceinline__ void select_emit_warp_gather(const SelUnit& U, const IndexOf& index_of_full, uint32_t warp_pts,
                                                        uint8_t* stage, uint8_t* gbuf, uint32_t cls, uint32_t r_begin = 0u,
                                                        uint32_t r_count = 0xFFFFFFFFu) {
  const uint32_t w = warp_id(), ln = lane_id();
  const uint32_t mine = r_count == 0xFFFFFFFFu ? U.warp_cnt[w] : r_count;
  uint32_t before = r_begin;
  for (uint32_t k = 0; k < w; ++k) before += U.warp_cnt[k];
  const unsigned long long out0 = U.out_rec + before;
  auto index_of = [&](uint32_t r) -> uint32_t { return index_of_full(r_begin + r); };
  const Segment& S = U.seg;
  const uint64_t wbase = U.u0 + (uint64_t)w * warp_pts;
  const uint64_t pol = l2_policy_drop();
  // lane l gathers, and later composes, matches l, l + 32, l + 64, l + 96 of a round: no other lane reads its slots.
  // (Measured alternatives, both slower and with the same DRAM bytes: the words of a round gathered as one coalesced
  // stream — lane t takes word t % 3 of match t / 3 — and, for sparse warp-units, the owner of a (row, lane) entry
  // gathering its own matches without any search: profiles/r02_notes.md.)
  auto gather = [&](uint32_t base, uint32_t buf) {
    const uint32_t n = min((uint32_t)kSelBDenseRecs, mine - base);
    const uint32_t dst0 = smem_u32(gbuf + buf * (uint32_t)kSelBGatherBytes);
#pragma unroll
    for (uint32_t j = 0; j < (uint32_t)kSelBDenseRecs / 32u; ++j) {
      const uint32_t r = ln + 32u * j;
      if (r < n) {
        const uint8_t* p = S.rec + (wbase + index_of(base + r)) * 12ull;
        cp_async4(dst0 + r * 12u, p);
        cp_async4(dst0 + r * 12u + 4u, p + 4);
        cp_async4(dst0 + r * 12u + 8u, p + 8);
      }
    }
    cp_async_commit();
  };
  gather(0u, 0u);
  uint32_t buf = 0;
#pragma unroll 1
  for (uint32_t base = 0; base < mine; base += (uint32_t)kSelBDenseRecs, buf ^= 1u) {
    const unsigned long long orec = out0 + base;
    const unsigned long long g0 = orec * 31ull;
    const uint32_t phase = (uint32_t)(g0 & 15ull);  // keep global and shared 16-byte phases equal
    const uint32_t n = min((uint32_t)kSelBDenseRecs, mine - base);
    if (base + (uint32_t)kSelBDenseRecs < mine) {
      gather(base + (uint32_t)kSelBDenseRecs, buf ^ 1u);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const int32_t* gp = reinterpret_cast<const int32_t*>(gbuf + buf * (uint32_t)kSelBGatherBytes);
#pragma unroll
    for (uint32_t j = 0; j < (uint32_t)kSelBDenseRecs / 32u; ++j) {
      const uint32_t r = ln + 32u * j;
      if (r < n) {
        RawPoint q;
        q.x = gp[r * 3u];
        q.y = gp[r * 3u + 1u];
        q.z = gp[r * 3u + 2u];
        q.cls = cls;
        q.r = q.g = q.b = 0u;  // Vector3::new(0, 0, 0), last.rs:152
        select_compose(S, q, stage + phase + r * 31u);
      }
    }
    __syncwarp();
    const unsigned long long room = orec < U.out_cap ? U.out_cap - orec : 0ull;
    const uint32_t nb = (room < (unsigned long long)n ? (uint32_t)room : n) * 31u;
    if (nb) select_flush<(kSelBDenseRecs * 31 + 15 + 511) / 512>(U.out, stage, g0, phase, nb, ln, pol);
    __syncwarp();  // the staging buffer is reused by the next round / unit
  }
}

// k_select_bytes, SPARSE warp-units of colourless LAST files (a rare class: a few hundred matches among a warp's 4096
// class bytes).  ncu on a 4 % class: 114 bytes of DRAM reads per gathered 12-byte position, and the kernel, moving
// 2.2 GB for 0.9 GB of algorithmic bytes, runs at 5.6 TB/s of DRAM traffic — it is bound by what a random 12-byte read
// costs in HBM, whatever the request looks like (three 4-byte requests, one coalesced word stream, aligned 16-byte
// chunks, cudaLimitMaxL2FetchGranularity = 32: the same DRAM bytes every time, profiles/r02_notes.md).  What the request
// shape does change is the work around it: a position fetched as the one or two aligned 16-byte chunks that hold it
// (cp.async.cg: past L1, 1.5 requests per position instead of 3) makes the sparse select 4 % faster.  Each lane gathers
// and later composes matches l, l + 32, l + 64 of a round of 96; a position's phase inside its first chunk stays in a
// register.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <class IndexOf>
__device__ __forceinline__ void select_emit_warp_chunks(const SelUnit& U, const IndexOf& index_of, uint32_t warp_pts, uint8_t* stage,
                                                        uint8_t* gbuf /* kSelBChunkRecs * 32 bytes */, uint32_t cls) {
  const uint32_t w = warp_id(), ln = lane_id();
  const uint32_t mine = U.warp_cnt[w];
  uint32_t before = 0;
  for (uint32_t k = 0; k < w; ++k) before += U.warp_cnt[k];
  const unsigned long long out0 = U.out_rec + before;
  const Segment& S = U.seg;
  const uint8_t* pos0 = S.rec + (U.u0 + (uint64_t)w * warp_pts) * 12ull;
  const uint64_t pol = l2_policy_drop();
  const uint32_t dst0 = smem_u32(gbuf);
#pragma unroll 1
  for (uint32_t base = 0; base < mine; base += (uint32_t)kSelBChunkRecs) {
    const unsigned long long orec = out0 + base;
    const unsigned long long g0 = orec * 31ull;
    const uint32_t phase = (uint32_t)(g0 & 15ull);  // keep global and shared 16-byte phases equal
    const uint32_t n = min((uint32_t)kSelBChunkRecs, mine - base);
    uint32_t off[kSelBChunkRecs / 32];  // byte offset of the position inside its first chunk: 0, 4, 8 or 12
#pragma unroll
    for (uint32_t j = 0; j < (uint32_t)kSelBChunkRecs / 32u; ++j) {
      const uint32_t r = ln + 32u * j;
      off[j] = 0u;
      if (r < n) {
        const uint8_t* p = pos0 + index_of(base + r) * 12u;
        off[j] = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u);
        const uint8_t* a = p - off[j];
        cp_async16(dst0 + r * 32u, a);
        if (off[j] > 4u) cp_async16(dst0 + r * 32u + 16u, a + 16);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
#pragma unroll
    for (uint32_t j = 0; j < (uint32_t)kSelBChunkRecs / 32u; ++j) {
      const uint32_t r = ln + 32u * j;
      if (r < n) {
        const int32_t* gp = reinterpret_cast<const int32_t*>(gbuf + r * 32u + off[j]);
        RawPoint q;
        q.x = gp[0];
        q.y = gp[1];
        q.z = gp[2];
        q.cls = cls;
        q.r = q.g = q.b = 0u;  // Vector3::new(0, 0, 0), last.rs:152
        select_compose(S, q, stage + phase + r * 31u);
      }
    }
    __syncwarp();
    const unsigned long long room = orec < U.out_cap ? U.out_cap - orec : 0ull;
    const uint32_t nb = (room < (unsigned long long)n ? (uint32_t)room : n) * 31u;
    if (nb) select_flush<(kSelBChunkRecs * 31 + 15 + 511) / 512>(U.out, stage, g0, phase, nb, ln, pol);
    __syncwarp();  // the staging and gather buffers are reused by the next round / unit
  }
}

// ------------------------------------------------------------------------------------------------
// MODE_SELECT for LAST class queries (last.rs:253-291): the predicate stream is ONE byte per point, so a 2048-point
// unit would be 2 KB and the kernel would run at the unit rate of the look-back machinery, not at memory speed.
// Same roles and barriers as k_select, but a unit is 32768 points: every consumer warp owns 4096 consecutive class
// bytes as 8 rows of 32 lanes x 16 bytes (one 16-byte load per lane and row, SIMD byte compare).  Instead of an index
// list the warp keeps, per lane and row, the 16-bit match mask and the number of matches before it; the emit expands
// them into an index list, half a warp-unit at a time (see below).
// ------------------------------------------------------------------------------------------------

template <int AL>
__global__ void __launch_bounds__(kSelThreads, 2) k_select_bytes(ScanParams P) {
  __shared__ SelUnit unit[kSelBufs];
  __shared__ __align__(8) uint64_t bar_tk[kSelBufs];
  __shared__ __align__(8) uint64_t bar_cnt[kSelBufs];
  __shared__ __align__(8) uint64_t bar_pre[kSelBufs];
  __shared__ __align__(8) uint64_t bar_free[kSelBufs];
  __shared__ DispatchCache dcache;
  extern __shared__ __align__(16) uint8_t selb_dsm[];  // per consumer warp: staging buffer, two gather buffers
  __shared__ uint16_t m_mask[kSelWarps][kSelBLag + 1][kSelBRows * 32];  // match mask of (row, lane)
  __shared__ uint16_t m_pre[kSelWarps][kSelBLag + 1][kSelBRows * 32];   // matches of the warp before (row, lane)

  const uint32_t ln = lane_id();
  select_init_barriers(bar_tk, bar_cnt, bar_pre, bar_free);
  if (warp_id() == kSelWarps + kSelLbWarps) {
    select_dispatcher_warp<kSelBUnitPts>(P, unit, bar_tk, bar_free, dcache);
    return;
  }
  if (warp_id() >= kSelWarps) {
    select_lookback_warp(P, unit, bar_tk, bar_cnt, bar_pre);
    return;
  }

  const uint32_t w = warp_id();
  const uint32_t pat = (P.cls & 0xFFu) * 0x01010101u;
  // The class bytes are read once and never again (the emit works from the match masks): they stream through L2 with
  // evict_first.  (They used to carry the `keep` policy of the record-reading select kernels, whose emit re-reads the
  // matching records; here that pinned 1 byte per point of dead data in L2 and pushed out what IS re-read — the second
  // and third word of every gathered position and the look-back descriptors.)
  const uint64_t pol_stream = PCQ_SELB_CLASS_POLICY;
  uint4 v[kSelBRows];
  uint32_t npts = 0;  // points of the unit whose class bytes are in (or on their way into) v
  auto load_row = [&](const uint8_t* col, uint32_t n_unit, int k) {
    const uint32_t i = (uint32_t)k * 512u + ln * 16u;
    v[k] = make_uint4(0u, 0u, 0u, 0u);
    // a partly valid 16-byte group is still loadable: columns are followed by other columns or by padding
    if (w * kSelBWarpPts + i < n_unit) v[k] = ldg_v4_h(col + i, pol_stream);
  };
  auto nibble = [&](uint32_t wd) -> uint32_t { return ((__vcmpeq4(wd, pat) & 0x08040201u) * 0x01010101u) >> 24; };

  mbar_wait(&bar_tk[0], 0u);
  bool cur = unit[0].tile != ~0ull;
  if (cur) {
    npts = unit[0].npts;
    const uint8_t* col = unit[0].seg.cls + unit[0].u0 + (uint64_t)w * kSelBWarpPts;
#pragma unroll
    for (int k = 0; k < kSelBRows; ++k) load_row(col, npts, k);
  }
  uint32_t emit_next = 0;
  for (uint32_t n = 0;; ++n) {
    const uint32_t b = n % kSelBufs;
    bool nxt = false;
    if (cur) {
      SelUnit& U = unit[b];
      // The next unit's class bytes are requested row by row, each as soon as the row it replaces has been turned
      // into its match mask: seven of a lane's eight 16-byte loads stay in flight at all times.  (Requesting the whole
      // next unit only after this one had been counted left a warp without a byte in flight for half of its time:
      // a select of a class that does not occur ran at 3.5 TB/s where the count of the same column reaches 6.)
      const uint32_t nb = (n + 1u) % kSelBufs;
      mbar_wait(&bar_tk[nb], ((n + 1u) / kSelBufs) & 1u);
      nxt = unit[nb].tile != ~0ull;
      const uint32_t npts_next = nxt ? unit[nb].npts : 0u;
      const uint8_t* ncol = unit[nb].seg.cls + unit[nb].u0 + (uint64_t)w * kSelBWarpPts;  // (only used when nxt)
      uint16_t* mk = m_mask[w][n % (uint32_t)(kSelBLag + 1)];
      uint16_t* pr = m_pre[w][n % (uint32_t)(kSelBLag + 1)];
      uint32_t cnt = 0;
      uint32_t m16s[kSelBRows];
      uint32_t any_bits = 0u;
#pragma unroll
      for (int k = 0; k < kSelBRows; ++k) {
        const uint32_t gi = w * kSelBWarpPts + (uint32_t)k * 512u + ln * 16u;
        uint32_t m16 = nibble(v[k].x) | (nibble(v[k].y) << 4) | (nibble(v[k].z) << 8) | (nibble(v[k].w) << 12);
        const uint32_t valid = gi < npts ? min(16u, npts - gi) : 0u;
        m16 &= (1u << valid) - 1u;
        m16s[k] = m16;
        any_bits |= m16;
        if (nxt) load_row(ncol, npts_next, k);
      }
      npts = npts_next;
      // ... and the class bytes of the unit after that are asked into L2 (one 128-byte line per lane), if the
      // dispatcher has described it already: a warp's eight loads are issued back to back and return together, so
      // without this a warp has 4 KB in flight for one DRAM latency per unit and nothing while it counts.
      if (nxt) {  // (two units ahead; three or four measured the same)
        const uint32_t nb2 = (n + 2u) % kSelBufs;
        if (mbar_test(&bar_tk[nb2], ((n + 2u) / kSelBufs) & 1u) && unit[nb2].tile != ~0ull) {
          const uint32_t off = w * (uint32_t)kSelBWarpPts + ln * 128u;
          if (off < unit[nb2].npts) prefetch_l2(unit[nb2].seg.cls + unit[nb2].u0 + off);
        }
      }
      // a warp without a single match in its 4096 bytes (every warp of a class that does not occur) posts zero and is
      // done: no prefix scans, no mask / prefix stores — its emit is skipped on warp_cnt == 0
      if (__any_sync(0xffffffffu, any_bits != 0u)) {
#pragma unroll
        for (int k = 0; k < kSelBRows; ++k) {
          const uint32_t m16 = m16s[k];
          const uint32_t c = (uint32_t)__popc(m16);
          uint32_t incl = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (ln >= (uint32_t)o) incl += t;
          }
          mk[k * 32 + ln] = (uint16_t)m16;
          pr[k * 32 + ln] = (uint16_t)(cnt + incl - c);
          cnt += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      __syncwarp();
      if (ln == 0) select_post_count(P, U, cnt, &bar_cnt[b]);
    }
    const uint32_t counted = cur ? n + 1u : n;
    while (emit_next < counted) {
      const uint32_t pb = emit_next % kSelBufs;
      const uint32_t par = (emit_next / kSelBufs) & 1u;
      if (nxt && emit_next + (uint32_t)kSelBLag >= counted) break;
      mbar_wait(&bar_pre[pb], par);
      if (!PCQ_HOOK(P, 2u)) {
        const uint16_t* mk = m_mask[w][emit_next % (uint32_t)(kSelBLag + 1)];
        const uint16_t* pr = m_pre[w][emit_next % (uint32_t)(kSelBLag + 1)];
        uint8_t* stage_w = selb_dsm + (size_t)w * kSelBWarpSmem;
        uint16_t* list = reinterpret_cast<uint16_t*>(stage_w + kSelBStageBytes + 2 * kSelBGatherBytes);
        const SelUnit& EU = unit[pb];
        const uint32_t mine = EU.warp_cnt[w];
        const bool plain = AL == 4 && EU.seg.rgb == nullptr;
        // The warp's matches are emitted in parts of at most kSelBListPts matches (whole rows of its 4096 bytes).  For each part the
        // index of every match is first written to a list — the owner of a (row, lane) entry walks the set bits of
        // its 16-bit mask and drops them at the entry's prefix, 32 matches per step of the warp — so that the emit
        // finds its r-th match with ONE shared-memory load (a binary search over the 256 prefixes per record was a
        // quarter of this kernel).
        // (a part = as many whole rows as fit the list: the whole warp-unit when matches are sparse, so that the
        // rounds of the emit stay as full as they can be)
        auto pre_at_row = [&](uint32_t row) -> uint32_t { return row < (uint32_t)kSelBRows ? (uint32_t)pr[row * 32u] : mine; };
        if (mine != 0u && mine <= (uint32_t)kSelBSearchMax) {
          // few matches (a rare class): the list would cost a pass over all 256 entries for a handful of records;
          // each lane finds its matches with a binary search over the prefixes instead
          auto index_of = [mk, pr](uint32_t r) -> uint32_t {
            uint32_t lo = 0, hi = (uint32_t)kSelBRows * 32u;  // first entry whose prefix exceeds r
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if ((uint32_t)pr[mid] > r) hi = mid; else lo = mid + 1u;
            }
            const uint32_t e = lo - 1u;  // holds the r-th match (an entry without matches never ends the search)
            return e * 16u + nth_set_bit((uint32_t)mk[e], r - (uint32_t)pr[e]);
          };
          if (plain && PCQ_SELB_CHUNKS)
            select_emit_warp_chunks(EU, index_of, (uint32_t)kSelBWarpPts, stage_w, stage_w + kSelBStageBytes, P.cls & 0xFFu);
          else if (plain && mine > (uint32_t)kSelBDenseRecs)
            select_emit_warp_gather(EU, index_of, (uint32_t)kSelBWarpPts, stage_w, stage_w + kSelBStageBytes, P.cls & 0xFFu);
          else
            select_emit_warp<AL>(EU, index_of, (uint32_t)kSelBWarpPts, stage_w);
          __syncwarp();
        }
        for (uint32_t row0 = 0; row0 < (uint32_t)kSelBRows && mine > (uint32_t)kSelBSearchMax;) {
          const uint32_t r_begin = pre_at_row(row0);
          uint32_t row1 = row0 + 1u;
          while (row1 < (uint32_t)kSelBRows && pre_at_row(row1 + 1u) - r_begin <= (uint32_t)kSelBListPts) ++row1;
          const uint32_t r_count = pre_at_row(row1) - r_begin;
          const uint32_t rows_lo = row0, rows_hi = row1;
          row0 = row1;
          if (r_count == 0u) continue;
          for (uint32_t row = rows_lo; row < rows_hi; ++row) {
            const uint32_t e = row * 32u + ln;
            uint32_t m = mk[e];
            uint32_t pos = (uint32_t)pr[e] - r_begin;
            while (__any_sync(0xffffffffu, m != 0u)) {
              if (m != 0u) {
                const uint32_t bit = (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                list[pos++] = (uint16_t)(e * 16u + bit);
              }
            }
          }
          __syncwarp();
          auto index_of = [list, r_begin](uint32_t r) -> uint32_t { return (uint32_t)list[r - r_begin]; };
          if (plain && r_count > (uint32_t)kSelBDenseRecs)
            select_emit_warp_gather(EU, index_of, (uint32_t)kSelBWarpPts, stage_w, stage_w + kSelBStageBytes, P.cls & 0xFFu, r_begin, r_count);
          else
            select_emit_warp<AL>(EU, index_of, (uint32_t)kSelBWarpPts, stage_w, r_begin, r_count);
          __syncwarp();
        }
      }
      __syncwarp();
      if (ln == 0) mbar_arrive(&bar_free[pb]);
      ++emit_next;
    }
    if (!nxt) break;
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------
// LAST class count fast path: the classification column is a plain byte stream (last.rs:253-262).
// 16 bytes per load, SIMD byte compare, one atomic per CTA and segment.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_class_count_soa(ScanParams P) {
  __shared__ unsigned long long s_red[kBlock / 32];
  const uint32_t pat = P.cls * 0x01010101u;
  for (uint32_t si = 0; si < P.n_segs; ++si) {
    const Segment& S = P.segs[si];
    const uint8_t* col = S.cls;
    const uint64_t n = S.n_points;
    unsigned long long acc = 0;
    // bytes before the first 16-byte boundary and after the last one are handled one by one
    uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(col) & 15u)) & 15u;
    if (head > n) head = n;
    const uint64_t nvec = (n - head) >> 4;
    const uint64_t tail0 = head + (nvec << 4);
    if (blockIdx.x == 0) {
      if (threadIdx.x < head) acc += (__ldg(col + threadIdx.x) == P.cls) ? 1ull : 0ull;
      if (threadIdx.x < n - tail0) acc += (__ldg(col + tail0 + threadIdx.x) == P.cls) ? 1ull : 0ull;
    }
    const uint4* v = reinterpret_cast<const uint4*>(col + head);
    const uint64_t stride = (uint64_t)gridDim.x * kBlock;
    uint64_t k = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    uint32_t c = 0;
    for (; k + 3 * stride < nvec; k += 4 * stride) {
      uint4 a0 = __ldg(v + k), a1 = __ldg(v + k + stride), a2 = __ldg(v + k + 2 * stride), a3 = __ldg(v + k + 3 * stride);
      c += __popc(__vcmpeq4(a0.x, pat)) + __popc(__vcmpeq4(a0.y, pat)) + __popc(__vcmpeq4(a0.z, pat)) + __popc(__vcmpeq4(a0.w, pat));
      c += __popc(__vcmpeq4(a1.x, pat)) + __popc(__vcmpeq4(a1.y, pat)) + __popc(__vcmpeq4(a1.z, pat)) + __popc(__vcmpeq4(a1.w, pat));
      c += __popc(__vcmpeq4(a2.x, pat)) + __popc(__vcmpeq4(a2.y, pat)) + __popc(__vcmpeq4(a2.z, pat)) + __popc(__vcmpeq4(a2.w, pat));
      c += __popc(__vcmpeq4(a3.x, pat)) + __popc(__vcmpeq4(a3.y, pat)) + __popc(__vcmpeq4(a3.z, pat)) + __popc(__vcmpeq4(a3.w, pat));
      if (c > 0x7FFF0000u) {
        acc += c >> 3;
        c = 0;
      }
    }
    for (; k < nvec; k += stride) {
      uint4 a0 = __ldg(v + k);
      c += __popc(__vcmpeq4(a0.x, pat)) + __popc(__vcmpeq4(a0.y, pat)) + __popc(__vcmpeq4(a0.z, pat)) + __popc(__vcmpeq4(a0.w, pat));
    }
    acc += c >> 3;  // __vcmpeq4 sets 8 bits per equal byte
    unsigned long long t = block_sum(acc, s_red);
    if (threadIdx.x == 0 && t) atomicAdd(P.lanes[S.lane].count, t);
  }
}

// ------------------------------------------------------------------------------------------------
// density table finalisation
// ------------------------------------------------------------------------------------------------

// keep only candidates that still hold their cell's minimum distance
// Slot of this thread's item in an output that is appended to through ONE global counter: the block's items of this
// iteration are counted in shared memory and placed with one atomic (a counter bumped once per warp serialises in its L2
// atomic unit: 280 k same-address atomics were 0.3 of the 0.4 ms k_grid_finalists took at navvis-XL).  Block-convergent.
__device__ __forceinline__ unsigned long long block_append(bool have, unsigned long long* counter, uint32_t* s_warp /* 8 */,
                                                           unsigned long long* s_base) {
  const uint32_t bal = __ballot_sync(0xffffffffu, have);
  const uint32_t w = warp_id(), ln = lane_id();
  if (ln == 0) s_warp[w] = (uint32_t)__popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += s_warp[k];
    *s_base = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  unsigned long long at = *s_base;
  for (uint32_t k = 0; k < w; ++k) at += s_warp[k];
  at += (unsigned long long)__popc(bal & ((1u << ln) - 1u));
  __syncthreads();  // s_warp / s_base are reused by the next iteration
  return at;
}

__global__ void __launch_bounds__(256) k_grid_prune(GridDev g, uint64_t n_in, Candidate* dst, unsigned long long* dst_count) {
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned long long s_base;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n_in; i0 += stride) {
    const uint64_t i = i0 + threadIdx.x;
    bool keep = false;
    uint4 c0, c1, c2, c3;
    if (i < n_in) {
      const uint4* c4 = reinterpret_cast<const uint4*>(g.cands + i);
      c0 = c4[0];
      c1 = c4[1];
      const uint64_t key = (uint64_t)c0.x | ((uint64_t)c0.y << 32);
      const unsigned long long d = (unsigned long long)c0.z | ((unsigned long long)c0.w << 32);
      const unsigned long long sidx = (unsigned long long)c1.x | ((unsigned long long)c1.y << 32);
      if (sidx != kCandEmpty && alias_find(g, key) == ~0u) {  // candidates of affected keys are dead weight
        const uint64_t slot = grid_slot(g, key, false);
        keep = slot != ~0ull && g.table[slot] == d;
      }
      if (keep) {
        c2 = c4[2];
        c3 = c4[3];
      }
    }
    const unsigned long long at = block_append(keep, dst_count, s_warp, &s_base);
    if (keep) {
      uint4* o4 = reinterpret_cast<uint4*>(dst + at);
      o4[0] = c0;
      o4[1] = c1;
      o4[2] = c2;
      o4[3] = c3;
    }
  }
}

// Finalisation (HashMap::values, grid_sampling.rs:111-113) works IN PLACE: the distance table doubles as the
// per-cell "smallest scan index among the candidates at the minimum distance" table, so a 2^27-cell grid needs no
// second 1 GB array (and no 1 GB memset per finalisation).  All over the candidate arena's FINALISTS:
//   k_grid_finalists      list the candidates that sit at their cell's minimum distance (about a quarter of the arena
//                         in dense data; every later step reads only those)
//   k_grid_final_min      the smallest scan index wins:  atomicMax(table[slot], winner code)
//   k_grid_emit           winner == finalist whose code is in the table
//   k_grid_final_restore  finalists put the distance back: table[slot] = dist bits   (the collector can go on collecting)
// A winner code has bit 63 set — no distance has (they are non-negative doubles) — and grows as the scan index
// shrinks, so ONE atomicMax both replaces the distance and keeps the first point in scan order (strict `<` of :97-102).
// (Round 1 kept a finalist flag in the candidates and walked the whole arena four times, two 32-byte sectors per
// candidate and pass: 0.96 ms at navvis-XL, as long as the insert itself.)
constexpr unsigned long long kWinTag = 1ull << 63, kWinClaim = 1ull << 62, kWinIdxMask = (1ull << 62) - 1ull;
__device__ __forceinline__ unsigned long long win_code(unsigned long long scan_idx) {
  return kWinTag | (kWinIdxMask - (scan_idx & kWinIdxMask));
}

__global__ void __launch_bounds__(256) k_grid_finalists(GridDev g, uint64_t n, uint32_t* list, unsigned long long* list_count) {
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned long long s_base;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
    const uint64_t i = i0 + threadIdx.x;
    bool fin = false;
    if (i < n) {
      const uint4* c4 = reinterpret_cast<const uint4*>(g.cands + i);
      const uint4 c0 = c4[0];
      const uint2 c1 = *reinterpret_cast<const uint2*>(c4 + 1);  // scan index (same 32-byte sector)
      const uint64_t key = (uint64_t)c0.x | ((uint64_t)c0.y << 32);
      const unsigned long long d = (unsigned long long)c0.z | ((unsigned long long)c0.w << 32);
      const unsigned long long sidx = (unsigned long long)c1.x | ((unsigned long long)c1.y << 32);
      const bool mine = g.own_parts <= 1u || (uint32_t)(mix64(key) % g.own_parts) == g.own_me;  // multi-GPU: owner only
      if (sidx != kCandEmpty && mine && alias_find(g, key) == ~0u) {  // affected keys come from the replay
        const uint64_t slot = grid_slot(g, key, false);
        fin = slot != ~0ull && g.table[slot] == d;
      }
    }
    const unsigned long long at = block_append(fin, list_count, s_warp, &s_base);
    if (fin) list[at] = (uint32_t)i;
  }
}

__global__ void k_grid_final_min(GridDev g, const uint32_t* list, const unsigned long long* list_count) {
  const uint64_t m = *list_count, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
    const Candidate& c = g.cands[list[j]];
    atomicMax(g.table + grid_slot(g, c.key, false), win_code((unsigned long long)c.scan_idx));
  }
}

__global__ void k_grid_final_restore(GridDev g, const uint32_t* list, const unsigned long long* list_count) {
  const uint64_t m = *list_count, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
    const Candidate& c = g.cands[list[j]];
    g.table[grid_slot(g, c.key, false)] = c.dist_bits;
  }
}

// winners -> 31-byte records (order arbitrary, like HashMap::values) or -> per-owner candidate parts
// mode 0: count per part, mode 1: write candidates into parts, mode 2: write 31-byte points
// (between k_grid_final_min and k_grid_final_restore: table[slot] holds the winner's code)
__global__ void __launch_bounds__(256) k_grid_emit(GridDev g, const uint32_t* list, const unsigned long long* list_count, int mode, uint32_t n_parts,
                            unsigned long long* part_counts, unsigned long long* part_cursor, Candidate* out_cands,
                            uint8_t* out_points, unsigned long long* out_count) {
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned long long s_base;
  const uint64_t m = *list_count, stride = (uint64_t)gridDim.x * blockDim.x;
  if (mode == 2) {
    // winners -> 31-byte points.  One atomic per block and iteration on the single output counter; a record goes out
    // as the same 14 stores sts_point31 uses, so the winners of a block fill one contiguous stretch of the output.
    for (uint64_t j0 = (uint64_t)blockIdx.x * blockDim.x; j0 < m; j0 += stride) {
      const uint64_t j = j0 + threadIdx.x;
      bool win = false;
      uint64_t i = 0;
      if (j < m) {
        i = list[j];
        const Candidate& c = g.cands[i];
        unsigned long long* cell = g.table + grid_slot(g, c.key, false);
        const unsigned long long mine = win_code((unsigned long long)c.scan_idx);
        // (claimed once even if the arena holds the same point twice)
        win = *cell == mine && atomicCAS(cell, mine, mine | kWinClaim) == mine;
      }
      const unsigned long long at = block_append(win, out_count, s_warp, &s_base);
      if (win) {
        const uint4* s4 = reinterpret_cast<const uint4*>(g.cands + i);  // the point sits at byte 24 of the candidate
        const uint4 a = s4[1], b = s4[2], c3 = s4[3];
        // bytes 24..54 of the candidate: a.z a.w | b.x b.y b.z b.w | c3.x c3.y (low 3 bytes)
        const uint32_t w[8] = {a.z, a.w, b.x, b.y, b.z, b.w, c3.x, c3.y & 0x00FFFFFFu};
        stg_point31(out_points + at * 31ull, w);
      }
    }
    return;
  }
  // Modes 0 and 1 (owner partitioning of the multi-GPU exchange).  The per-part counters are a handful of addresses:
  // one atomic per winner serialised 2.4 M atomics on 2-8 words (2.9 ms of export at navvis-XL on two GPUs), so the
  // winners of a warp that go to the same part are counted / placed with ONE atomic (__match_any_sync on the part).
  for (uint64_t j0 = (uint64_t)blockIdx.x * blockDim.x; j0 < m; j0 += stride) {
    const uint64_t j = j0 + threadIdx.x;
    bool win = false;
    uint32_t part = 0xFFFFFFFFu;
    uint64_t i = 0;
    if (j < m) {
      i = list[j];
      const Candidate& c = g.cands[i];
      unsigned long long* cell = g.table + grid_slot(g, c.key, false);
      const unsigned long long mine = win_code((unsigned long long)c.scan_idx);
      const unsigned long long claimed = mine | kWinClaim;
      if (mode == 0) {
        // count the winner once even if the candidate list holds duplicates of it
        win = *cell == mine && atomicCAS(cell, mine, claimed) == mine;
      } else {
        // second walk after mode 0: claimed entries carry the claim bit; release the claim while emitting
        win = *cell == claimed && atomicCAS(cell, claimed, mine) == claimed;
      }
      if (win) part = (uint32_t)(mix64(c.key) % n_parts);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, part);  // lanes of this warp that go to the same part
    if (!win) continue;
    const uint32_t leader = (uint32_t)__ffs((int)peers) - 1u;
    const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane_id()) - 1u));
    if (mode == 0) {
      if (lane_id() == leader) atomicAdd(part_counts + part, (unsigned long long)__popc(peers));
    } else {
      unsigned long long base = 0;
      if (lane_id() == leader) base = atomicAdd(part_cursor + part, (unsigned long long)__popc(peers));
      base = __shfl_sync(peers, base, (int)leader);
      const uint4* s4 = reinterpret_cast<const uint4*>(g.cands + i);
      uint4* o4 = reinterpret_cast<uint4*>(out_cands + base + rank);
      o4[0] = s4[0];
      o4[1] = s4[1];
      o4[2] = s4[2];
      o4[3] = s4[3];
    }
  }
}

// fold candidates received from peers into the table (multi-GPU density merge)
__global__ void __launch_bounds__(256) k_grid_import(GridDev g, const Candidate* in, uint64_t n) {
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned long long s_base;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
    const uint64_t i = i0 + threadIdx.x;
    bool want = false;
    uint4 c0, c1, c2, c3;
    if (i < n) {
      const uint4* c4 = reinterpret_cast<const uint4*>(in + i);
      c0 = c4[0];
      c1 = c4[1];
      c2 = c4[2];
      c3 = c4[3];
      const uint64_t key = (uint64_t)c0.x | ((uint64_t)c0.y << 32);
      const unsigned long long d = (unsigned long long)c0.z | ((unsigned long long)c0.w << 32);
      const uint64_t slot = grid_slot(g, key, true);
      if (slot == ~0ull) {
        atomicOr(g.flags, kFlagHashFull);
      } else {
        unsigned long long old = atomicMin(g.table + slot, d);
        want = d <= old;
      }
    }
    const unsigned long long ci = block_append(want, g.cand_count, s_warp, &s_base);
    if (want) {
      if (ci < g.cand_cap) {
        uint4* o4 = reinterpret_cast<uint4*>(g.cands + ci);
        o4[0] = c0;
        o4[1] = c1;
        o4[2] = c2;
        o4[3] = c3;
      } else {
        atomicOr(g.flags, kFlagCandOverflow);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------
static int check_launch() { return cudaGetLastError() == cudaSuccess ? 0 : -1; }

// ring depth of the scan kernels: enough stages for >= ~48 KB in flight per CTA
template <int R>
struct ScanStages {
  static constexpr int value = R <= 12 ? 8 : (R <= 20 ? 6 : 4);
};
static int persistent_grid(const void* kfn, size_t smem, int sm_count, uint64_t n_tiles, int max_per_sm, unsigned* grid_out) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kBlock, smem) != cudaSuccess) return -1;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > max_per_sm) per_sm = max_per_sm;
  uint64_t grid = (uint64_t)sm_count * (uint64_t)per_sm;  // persistent: every CTA resident at once
  if (grid > n_tiles) grid = n_tiles;
  *grid_out = (unsigned)grid;
  return 0;
}

template <int R, int MODE>
static int launch_staged_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  constexpr int TP = (R == 12 && MODE == MODE_COUNT) ? kTilePtsPos : kTilePts;
  constexpr int STAGES = MODE == MODE_GRID ? 2 : (TP == kTilePtsPos ? 5 : ScanStages<R>::value);  // (MODE_GRIDQ streams like a count)
  constexpr size_t smem = (size_t)STAGES * TP * R;
  static bool configured = false;
  auto kfn = k_scan_staged<R, MODE, STAGES, TP>;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  unsigned grid = 0;
  if (persistent_grid((const void*)kfn, smem, sm_count, p.n_tiles, (int)kGridCtasPerSm, &grid) != 0) return -1;
  if (grid == 0) return 0;
  kfn<<<grid, kBlock, smem, st>>>(p);
  return check_launch();
}

template <int R, bool QUEUE, int PPT, bool ONE>
static int launch_grid_scan_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  // dense inserts are bound by the table reads in flight (registers, resident warps), and a fourth stage costs a CTA
  // (navvis-XL, two records per lane: 1.43 ms with three stages, 1.96 ms with four); the queue variant runs two CTAs
  // per SM anyway and gains from the deeper ring (navvis-L: 0.44 vs 0.48 ms)
  using Sh = GsShape<PPT>;
  constexpr int STAGES = QUEUE ? (R <= 12 ? 8 : (R <= 20 ? 6 : 4)) : (R <= 12 ? 6 : (R <= 20 ? 4 : 3));
  constexpr size_t smem = (size_t)STAGES * kTilePts * R + (QUEUE ? (size_t)Sh::kWarps * Sh::kQCap * sizeof(GridQEntry) : 0);
  static bool configured = false;
  auto kfn = k_grid_scan<R, STAGES, QUEUE, PPT, ONE>;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Sh::kThreads, smem) != cudaSuccess) return -1;
  if (std::getenv("PCQ_VERBOSE"))
    std::fprintf(stderr, "k_grid_scan<%d,%d,%d,%d,%d>: %d CTAs per SM, %zu bytes dynamic shared memory\n", R, STAGES, (int)QUEUE, PPT, (int)ONE, per_sm, smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > (int)kGridCtasPerSm) per_sm = (int)kGridCtasPerSm;
  uint64_t grid = (uint64_t)sm_count * (uint64_t)per_sm;  // persistent: every CTA resident at once
  if (grid > p.n_tiles) grid = p.n_tiles;
  if (grid == 0) return 0;
  kfn<<<(unsigned)grid, Sh::kThreads, smem, st>>>(p);
  return check_launch();
}
template <bool QUEUE, int PPT, bool ONE>
static int launch_grid_scan_r(const ScanParams& p, uint32_t R, int sm_count, cudaStream_t st) {
  switch (R) {
    case 12: return launch_grid_scan_t<12, QUEUE, PPT, ONE>(p, sm_count, st);
    case 20: return launch_grid_scan_t<20, QUEUE, PPT, ONE>(p, sm_count, st);
    case 26: return launch_grid_scan_t<26, QUEUE, PPT, ONE>(p, sm_count, st);
    case 28: return launch_grid_scan_t<28, QUEUE, PPT, ONE>(p, sm_count, st);
    case 34: return launch_grid_scan_t<34, QUEUE, PPT, ONE>(p, sm_count, st);
    default: return 1;  // not instantiated
  }
}
static int launch_grid_scan(const ScanParams& p, uint32_t R, int sm_count, cudaStream_t st) {
  if (p.grid_sparse)
    return p.one_grid ? launch_grid_scan_r<true, 2, true>(p, R, sm_count, st) : launch_grid_scan_r<true, 2, false>(p, R, sm_count, st);
  return p.one_grid ? launch_grid_scan_r<false, 2, true>(p, R, sm_count, st) : launch_grid_scan_r<false, 2, false>(p, R, sm_count, st);
}

template <int MODE>
static int launch_staged_r(const ScanParams& p, uint32_t R, int sm_count, cudaStream_t st) {
  switch (R) {
    case 12: return launch_staged_t<12, MODE>(p, sm_count, st);
    case 20: return launch_staged_t<20, MODE>(p, sm_count, st);
    case 26: return launch_staged_t<26, MODE>(p, sm_count, st);
    case 28: return launch_staged_t<28, MODE>(p, sm_count, st);
    case 34: return launch_staged_t<34, MODE>(p, sm_count, st);
    default: return 1;  // not instantiated
  }
}

template <int MODE>
static int launch_direct_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  auto kfn = k_scan_direct<MODE>;
  unsigned grid = 0;
  if (persistent_grid((const void*)kfn, 0, sm_count, p.n_tiles, (MODE == MODE_GRID || MODE == MODE_GRIDQ) ? (int)kGridCtasPerSm : 8, &grid) != 0) return -1;
  if (grid == 0) return 0;
  kfn<<<grid, kBlock, 0, st>>>(p);
  return check_launch();
}

template <int AL>
static int launch_select_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  auto kfn = k_select<AL>;
  unsigned grid = 0;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kSelThreads, 0) != cudaSuccess) return -1;
  if (per_sm < 1) per_sm = 1;
  uint64_t g = (uint64_t)sm_count * (uint64_t)per_sm;  // persistent: every CTA resident at once
  if (g > p.n_tiles) g = p.n_tiles;
  grid = (unsigned)g;
  if (grid == 0) return 0;
  kfn<<<grid, kSelThreads, 0, st>>>(p);
  return check_launch();
}
template <int AL>
static int launch_select_bytes_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  auto kfn = k_select_bytes<AL>;
  constexpr size_t smem = (size_t)kSelWarps * kSelBWarpSmem;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kSelThreads, smem) != cudaSuccess) return -1;
  if (std::getenv("PCQ_VERBOSE")) std::fprintf(stderr, "k_select_bytes<%d>: %d CTAs per SM, %zu bytes dynamic shared memory\n", AL, per_sm, smem);
  if (per_sm < 1) per_sm = 1;
  uint64_t g = (uint64_t)sm_count * (uint64_t)per_sm;
  if (g > p.n_tiles) g = p.n_tiles;
  if (g == 0) return 0;
  kfn<<<(unsigned)g, kSelThreads, smem, st>>>(p);
  return check_launch();
}
template <int R>
static int launch_select_ring_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  auto kfn = k_select_ring<R>;
  constexpr size_t smem = SelRing<R>::smem;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  uint64_t g = (uint64_t)sm_count;  // one CTA per SM, all resident
  if (g > p.n_tiles) g = p.n_tiles;
  if (g == 0) return 0;
  kfn<<<(unsigned)g, kSelRThreads, smem, st>>>(p);
  return check_launch();
}
// `align` = alignment every x/y/z field of every segment of the launch is guaranteed to have (4, 2 or 1)
static int launch_select(const ScanParams& p, int align, int sm_count, cudaStream_t st) {
  switch (p.sel_ring) {  // record length of a launch that qualifies for k_select_ring, else 0
    case 12: return launch_select_ring_t<12>(p, sm_count, st);
    case 20: return launch_select_ring_t<20>(p, sm_count, st);
    case 26: return launch_select_ring_t<26>(p, sm_count, st);
    case 28: return launch_select_ring_t<28>(p, sm_count, st);
    case 34: return launch_select_ring_t<34>(p, sm_count, st);
    default: break;
  }
  if (p.sel_bytes) {
    if (align >= 4) return launch_select_bytes_t<4>(p, sm_count, st);
    if (align >= 2) return launch_select_bytes_t<2>(p, sm_count, st);
    return launch_select_bytes_t<1>(p, sm_count, st);
  }
  if (align >= 4) return launch_select_t<4>(p, sm_count, st);
  if (align >= 2) return launch_select_t<2>(p, sm_count, st);
  return launch_select_t<1>(p, sm_count, st);
}

bool staged_supports(uint32_t R) { return R == 12 || R == 20 || R == 26 || R == 28 || R == 34; }

// records per scheduling unit ("tile") for a launch: 512 for count / grid, a look-back unit for select
uint32_t tile_points(int variant, int mode, uint32_t R, bool select_bytes) {
  if (mode == MODE_SELECT) {
    if (select_bytes) return (uint32_t)kSelBUnitPts;
    if (variant == 2 && staged_supports(R)) return sel_ring_unit_points(R);  // k_select_ring
    return (uint32_t)kSelUnitPts;
  }
  if (mode == MODE_COUNT && variant == 2 && R == 12) return (uint32_t)kTilePtsPos;  // LAST positions, staged
  return (uint32_t)kTilePts;
}

// variant: 1 = direct, 2 = staged (needs uniform_record_len supported); returns 0 ok, <0 CUDA error
int launch_scan(int variant, int mode, const ScanParams& p, uint32_t uniform_record_len, int min_align, int sm_count,
                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == MODE_SELECT) return launch_select(p, min_align, sm_count, st);  // one kernel family serves every layout
  if (variant == 2 && staged_supports(uniform_record_len)) {
    int rc = 1;
    if (mode == MODE_COUNT) rc = launch_staged_r<MODE_COUNT>(p, uniform_record_len, sm_count, st);
    if (mode == MODE_GRID) {
      const char* old_k = std::getenv("PCQ_GRID_KERNEL");  // "staged": the barrier-per-tile kernels (measurement only)
      if (old_k && std::strcmp(old_k, "staged") == 0)
        rc = p.grid_sparse ? launch_staged_r<MODE_GRIDQ>(p, uniform_record_len, sm_count, st)
                           : launch_staged_r<MODE_GRID>(p, uniform_record_len, sm_count, st);
      else
        rc = launch_grid_scan(p, uniform_record_len, sm_count, st);
    }
    if (rc <= 0) return rc;
  }
  if (mode == MODE_COUNT) return launch_direct_t<MODE_COUNT>(p, sm_count, st);
  return p.grid_sparse ? launch_direct_t<MODE_GRIDQ>(p, sm_count, st) : launch_direct_t<MODE_GRID>(p, sm_count, st);
}

int launch_class_count_soa(const ScanParams& p, int sm_count, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  k_class_count_soa<<<(unsigned)(sm_count * 8), kBlock, 0, st>>>(p);
  return check_launch();
}

static unsigned grid_for(uint64_t n, int sm_count) {
  uint64_t g = (n + 255) / 256;
  uint64_t cap = (uint64_t)sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

int launch_grid_prune(const GridDev& g, uint64_t n_in, Candidate* dst, unsigned long long* dst_count, int sm_count,
                      void* stream) {
  if (n_in == 0) return 0;
  k_grid_prune<<<grid_for(n_in, sm_count), 256, 0, (cudaStream_t)stream>>>(g, n_in, dst, dst_count);
  return check_launch();
}

int launch_grid_finalists(const GridDev& g, uint64_t n, uint32_t* list, unsigned long long* list_count, int sm_count, void* stream) {
  if (n == 0) return 0;
  k_grid_finalists<<<grid_for(n, sm_count), 256, 0, (cudaStream_t)stream>>>(g, n, list, list_count);
  return check_launch();
}
// (the list's length lives on the device; n_max, the number of candidates it was made from, sizes the grid)
int launch_grid_final_min(const GridDev& g, uint64_t n_max, const uint32_t* list, const unsigned long long* list_count, int sm_count,
                          void* stream) {
  if (n_max == 0) return 0;
  k_grid_final_min<<<grid_for(n_max, sm_count), 256, 0, (cudaStream_t)stream>>>(g, list, list_count);
  return check_launch();
}
int launch_grid_final_restore(const GridDev& g, uint64_t n_max, const uint32_t* list, const unsigned long long* list_count, int sm_count,
                              void* stream) {
  if (n_max == 0) return 0;
  k_grid_final_restore<<<grid_for(n_max, sm_count), 256, 0, (cudaStream_t)stream>>>(g, list, list_count);
  return check_launch();
}

int launch_grid_emit(const GridDev& g, uint64_t n_max, const uint32_t* list, const unsigned long long* list_count, int mode,
                     uint32_t n_parts, unsigned long long* part_counts, unsigned long long* part_cursor, Candidate* out_cands,
                     uint8_t* out_points, unsigned long long* out_count, int sm_count, void* stream) {
  if (n_max == 0) return 0;
  k_grid_emit<<<grid_for(n_max, sm_count), 256, 0, (cudaStream_t)stream>>>(g, list, list_count, mode, n_parts, part_counts, part_cursor,
                                                                             out_cands, out_points, out_count);
  return check_launch();
}

// one D2H copy instead of one per lane: the per-collector scalar blocks (first `block_bytes` bytes at lanes[l].count)
__global__ void k_gather_blocks(const LaneDev* lanes, uint32_t n, uint32_t words, uint32_t* out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * words; i += gridDim.x * blockDim.x)
    out[i] = reinterpret_cast<const uint32_t*>(lanes[i / words].count)[i % words];
}
int launch_gather_blocks(const LaneDev* lanes, uint32_t n, uint32_t block_bytes, void* out, void* stream) {
  if (n == 0) return 0;
  const uint32_t words = block_bytes / 4u;
  k_gather_blocks<<<(n * words + 255u) / 256u, 256, 0, (cudaStream_t)stream>>>(lanes, n, words, static_cast<uint32_t*>(out));
  return check_launch();
}

int launch_grid_import(const GridDev& g, const Candidate* in, uint64_t n, int sm_count, void* stream) {
  if (n == 0) return 0;
  k_grid_import<<<grid_for(n, sm_count), 256, 0, (cudaStream_t)stream>>>(g, in, n);
  return check_launch();
}

}  // namespace pcq
