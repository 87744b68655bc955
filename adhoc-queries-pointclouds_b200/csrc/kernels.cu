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

#include <cstdint>

#include "pcq_device.h"

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

// (v as f64 * scale) + offset, las.rs:139-141 — two roundings, never an FMA
__device__ __forceinline__ double reconstruct(int32_t v, double scale, double offset) {
  return __dadd_rn(__dmul_rn((double)v, scale), offset);
}

// Rust `f64 as u64`: NaN -> 0, negative -> 0, saturating (grid_sampling.rs:58-60)
__device__ __forceinline__ uint64_t f64_as_u64(double v) {
  if (!(v > 0.0)) return 0ull;  // NaN, -x, +-0
  if (v >= 18446744073709551616.0) return ~0ull;
  return (uint64_t)v;  // truncates toward zero
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
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

// Store a 31-byte record at an arbitrarily aligned shared-memory address: 7 word stores + 3 byte
// stores instead of 31 byte stores.
__device__ __forceinline__ void sts_point31(uint8_t* dst, const uint32_t w[8]) {
  const uint32_t o = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u);
  if (o == 0) {
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
    for (int j = 0; j < 7; ++j) d[j] = w[j];
    dst[28] = (uint8_t)w[7];
    dst[29] = (uint8_t)(w[7] >> 8);
    dst[30] = (uint8_t)(w[7] >> 16);
    return;
  }
  const uint32_t head = 4u - o;  // bytes before the first aligned word
  for (uint32_t b = 0; b < head; ++b) dst[b] = (uint8_t)(w[0] >> (8u * b));
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + head);
  const uint32_t sh = 8u * head;
#pragma unroll
  for (int j = 0; j < 7; ++j) d[j] = __funnelshift_r(w[j], w[j + 1], sh);  // stream bytes 4j+head .. +3
  // bytes written so far: head + 28; remaining = 3 - head
  for (uint32_t b = head + 28u; b < 31u; ++b) dst[b] = (uint8_t)(w[7] >> (8u * (b & 3u)));  // b>>2 == 7
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

// Executed by all 32 lanes of warp 0.  Returns the number of matches in units
// [lane_first_tile, tile) — the exclusive prefix of `tile` within its lane.  Every lane inspects
// four descriptors per round trip (window of 128 units), closest units in the lowest lanes.
__device__ __forceinline__ unsigned long long lookback_exclusive(const unsigned long long* state, uint64_t tile,
                                                                 uint64_t lane_first_tile) {
  unsigned long long excl = 0;
  long long hi = (long long)tile - 1;
  const long long lo = (long long)lane_first_tile;
  const uint32_t ln = lane_id();
  while (hi >= lo) {
    unsigned long long sv[4];
    bool pending;
    do {
      pending = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long t = hi - (long long)(4u * ln + (uint32_t)k);
        sv[k] = t >= lo ? ld_state(state + t) : (kStPrefix << kStatusShift);  // below the lane start: prefix 0
        pending |= (sv[k] >> kStatusShift) == 0ull;
      }
    } while (__any_sync(0xffffffffu, pending));
    int kf = 4;  // first (closest) descriptor of this lane that already carries an inclusive prefix
#pragma unroll
    for (int k = 3; k >= 0; --k)
      if ((sv[k] >> kStatusShift) == kStPrefix) kf = k;
    const uint32_t pm = __ballot_sync(0xffffffffu, kf < 4);
    const uint32_t first = pm ? (uint32_t)__ffs((int)pm) - 1u : 32u;
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool take = ln < first || (ln == first && k <= kf);
      if (take) v += sv[k] & kValueMask;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (pm != 0u) break;
    hi -= 128;
  }
  return excl;
}

// ------------------------------------------------------------------------------------------------
// SparseGrid::insert_point as an atomic-min table (grid_sampling.rs:49-105)
// ------------------------------------------------------------------------------------------------

// find-or-insert the slot of `key`.  dense: slot == key.  hashed: linear probing on hkeys.
__device__ __forceinline__ uint64_t grid_slot(const GridDev& g, uint64_t key, bool insert) {
  if (g.hkeys == nullptr) return key;
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

struct CellEval {
  uint64_t key;
  unsigned long long dist_bits;
  bool aliased;
};

__device__ __forceinline__ CellEval grid_eval(const GridDev& g, double px, double py, double pz) {
  // :51-56  r = (p - min) * dims as f64 / (max - min)
  double rx = __ddiv_rn(__dmul_rn(__dsub_rn(px, g.bmin[0]), g.dims_f[0]), __dsub_rn(g.bmax[0], g.bmin[0]));
  double ry = __ddiv_rn(__dmul_rn(__dsub_rn(py, g.bmin[1]), g.dims_f[1]), __dsub_rn(g.bmax[1], g.bmin[1]));
  double rz = __ddiv_rn(__dmul_rn(__dsub_rn(pz, g.bmin[2]), g.dims_f[2]), __dsub_rn(g.bmax[2], g.bmin[2]));
  // :58-60
  uint64_t cx = f64_as_u64(rx), cy = f64_as_u64(ry), cz = f64_as_u64(rz);
  CellEval e;
  // a cell above its mask aliases a low cell while its centre lies elsewhere (:62-70 vs :78-82)
  e.aliased = (cx > g.mask[0]) | (cy > g.mask[1]) | (cz > g.mask[2]);
  e.key = (cx & g.mask[0]) | ((cy & g.mask[1]) << g.shift_y) | ((cz & g.mask[2]) << g.shift_z);
  // :78-82  centre = (cell as f64 + 0.5) * cell_size + min   (unmasked cell)
  double ccx = __dadd_rn(__dmul_rn(__dadd_rn(__ull2double_rn(cx), 0.5), g.cell_size), g.bmin[0]);
  double ccy = __dadd_rn(__dmul_rn(__dadd_rn(__ull2double_rn(cy), 0.5), g.cell_size), g.bmin[1]);
  double ccz = __dadd_rn(__dmul_rn(__dadd_rn(__ull2double_rn(cz), 0.5), g.cell_size), g.bmin[2]);
  // :84-95  distance_squared = (dx*dx + dy*dy) + dz*dz  (nalgebra 0.23, no FMA)
  double dx = __dsub_rn(ccx, px), dy = __dsub_rn(ccy, py), dz = __dsub_rn(ccz, pz);
  double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  e.dist_bits = (unsigned long long)__double_as_longlong(d);  // d >= +0: bit order == value order
  return e;
}

// Warp-convergent: every lane calls it, `m` says whether this lane carries a matching point.
// A point survives as a candidate iff its distance is <= the cell minimum seen so far; the true
// winner (smallest distance, then smallest scan index == the strict `<` fold of :97-102) always is.
template <class Src>
__device__ __forceinline__ void grid_insert(const GridDev& g, const Segment& S, const Src& src, bool m,
                                            const Hit& h, uint64_t idx, uint32_t i_in_tile) {
  bool want = false;
  CellEval e;
  e.key = 0;
  e.dist_bits = 0;
  e.aliased = false;
  if (m) {
    double px = reconstruct(h.x, S.scale[0], S.offset[0]);
    double py = reconstruct(h.y, S.scale[1], S.offset[1]);
    double pz = reconstruct(h.z, S.scale[2], S.offset[2]);
    e = grid_eval(g, px, py, pz);
    if (e.aliased) {
      atomicOr(g.flags, kFlagAliased);
    } else {
      uint64_t slot = grid_slot(g, e.key, true);
      if (slot == ~0ull) {
        atomicOr(g.flags, kFlagHashFull);
      } else {
        // The cell minimum only ever decreases, so a (possibly stale) plain read that is already smaller than
        // this point's distance proves the point can never win: skip the atomic.  In dense data (many points
        // per cell) that removes most of the read-modify-write traffic.
        const unsigned long long seen = __ldcg(g.table + slot);
        if (e.dist_bits <= seen) {
          const unsigned long long old = atomicMin(g.table + slot, e.dist_bits);
          want = e.dist_bits <= old;
        }
      }
    }
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, want);
  if (bal == 0u) return;
  const uint32_t leader = (uint32_t)__ffs((int)bal) - 1u;
  unsigned long long base = 0;
  if (lane_id() == leader) base = atomicAdd(g.cand_count, (unsigned long long)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, (int)leader);
  if (want) {
    const unsigned long long ci = base + (unsigned long long)__popc(bal & ((1u << lane_id()) - 1u));
    if (ci < g.cand_cap) {
      uint32_t rgb[3];
      src.colour(S, idx, i_in_tile, rgb);
      uint32_t w[8];
      point_words(S, h, rgb, w);
      uint4* c4 = reinterpret_cast<uint4*>(g.cands + ci);
      const unsigned long long gidx = S.scan_base + idx;
      c4[0] = make_uint4((uint32_t)e.key, (uint32_t)(e.key >> 32), (uint32_t)e.dist_bits, (uint32_t)(e.dist_bits >> 32));
      c4[1] = make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), w[0], w[1]);
      c4[2] = make_uint4(w[2], w[3], w[4], w[5]);
      c4[3] = make_uint4(w[6], w[7] & 0x00FFFFFFu, 0u, 0u);
    } else {
      atomicOr(g.flags, kFlagCandOverflow);
    }
  }
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

template <int MODE, class Src>
__device__ __forceinline__ void process_tile(const ScanParams& P, const Segment& S, uint64_t tile, const Src& src,
                                             unsigned long long& acc) {
  static_assert(MODE == MODE_COUNT || MODE == MODE_GRID, "select has its own kernels");
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
  } else {
    const GridDev& g = P.lanes[S.lane].grid;
#pragma unroll
    for (int j = 0; j < kPPT; ++j) {
      const uint32_t i = (uint32_t)j * kBlock + tid;
      grid_insert(g, S, src, m[j], h[j], p0 + i, i);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// direct scan kernel: persistent CTAs, plain global loads
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_scan_direct(ScanParams P) {
  __shared__ Segment sseg;
  __shared__ unsigned long long s_red[kBlock / 32];

  uint32_t seg_i = 0xFFFFFFFFu;
  uint32_t seg_cursor = 0;
  unsigned long long acc = 0;
  DirectSrc src;

  for (uint64_t iter = 0;; ++iter) {
    const uint64_t tile = (uint64_t)blockIdx.x + iter * (uint64_t)gridDim.x;
    if (tile >= P.n_tiles) break;
    while (seg_cursor + 1 < P.n_segs && tile >= P.segs[seg_cursor + 1].first_tile) ++seg_cursor;
    if (seg_cursor != seg_i) {
      // segment change: flush the per-lane match count, cache the new segment in shared memory
      if constexpr (MODE == MODE_COUNT) {
        if (seg_i != 0xFFFFFFFFu) {
          unsigned long long t = block_sum(acc, s_red);
          if (threadIdx.x == 0 && t) atomicAdd(P.lanes[sseg.lane].count, t);
          acc = 0;
        }
      }
      __syncthreads();
      const uint32_t* srcw = reinterpret_cast<const uint32_t*>(P.segs + seg_cursor);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&sseg);
      for (uint32_t k = threadIdx.x; k < sizeof(Segment) / 4; k += kBlock) dstw[k] = srcw[k];
      seg_i = seg_cursor;
      __syncthreads();
    }
    process_tile<MODE>(P, sseg, tile, src, acc);
  }
  if constexpr (MODE == MODE_COUNT) {
    if (seg_i != 0xFFFFFFFFu) {
      unsigned long long t = block_sum(acc, s_red);
      if (threadIdx.x == 0 && t) atomicAdd(P.lanes[sseg.lane].count, t);
    }
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

template <int R, int MODE, int STAGES>
__global__ void __launch_bounds__(kBlock) k_scan_staged(ScanParams P) {
  constexpr uint32_t kTileBytes = (uint32_t)kTilePts * (uint32_t)R;
  extern __shared__ __align__(128) uint8_t dsm[];  // STAGES * kTileBytes
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ unsigned long long stage_tile[STAGES];
  __shared__ uint32_t stage_seg[STAGES];
  __shared__ Segment sseg;
  __shared__ unsigned long long s_red[kBlock / 32];

  const uint32_t tid = threadIdx.x;
  // producer state (meaningful in thread 0 only)
  uint32_t prod_seg = 0;
  uint64_t prod_iter = 0;

  auto produce = [&](int s) {
    const uint64_t tile = (uint64_t)blockIdx.x + prod_iter * (uint64_t)gridDim.x;
    ++prod_iter;
    if (tile >= P.n_tiles) {
      stage_tile[s] = ~0ull;
      return;
    }
    while (prod_seg + 1 < P.n_segs && tile >= P.segs[prod_seg + 1].first_tile) ++prod_seg;
    const Segment* sg = P.segs + prod_seg;
    const uint64_t p0 = (tile - sg->first_tile) * (uint64_t)kTilePts;
    const uint64_t rem = sg->n_points - p0;
    const uint32_t npts = rem < (uint64_t)kTilePts ? (uint32_t)rem : (uint32_t)kTilePts;
    const uint32_t bytes = (npts * (uint32_t)R + 15u) & ~15u;  // bulk copies move multiples of 16 bytes
    stage_tile[s] = tile;
    stage_seg[s] = prod_seg;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
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

  uint32_t seg_i = 0xFFFFFFFFu;
  unsigned long long acc = 0;
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
      __syncthreads();
      const uint32_t* srcw = reinterpret_cast<const uint32_t*>(P.segs + seg_now);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&sseg);
      for (uint32_t k = tid; k < sizeof(Segment) / 4; k += kBlock) dstw[k] = srcw[k];
      seg_i = seg_now;
      __syncthreads();
    }
    mbar_wait(&full_bar[s], parity);
    SmemSrc<R> src{dsm + (size_t)s * kTileBytes};
    process_tile<MODE>(P, sseg, tile, src, acc);
    __syncthreads();  // every thread is done reading stage s
    if (tid == 0) produce((int)s);
  }
  if constexpr (MODE == MODE_COUNT) {
    if (seg_i != 0xFFFFFFFFu) {
      unsigned long long t = block_sum(acc, s_red);
      if (tid == 0 && t) atomicAdd(P.lanes[sseg.lane].count, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// MODE_SELECT — BufferCollector::collect_one in scan order (collect_points.rs:29-31): single-pass
// stable stream compaction.
//
// The unit of the decoupled look-back is a GROUP of up to kMaxGroup sub-tiles (512 records each)
// handled by one CTA: (1) count phase — evaluate every sub-tile, keep per-warp match counts in
// shared memory; (2) publish the unit's aggregate, take the ticket of the NEXT unit, look back
// over earlier units for the exclusive prefix; (3) emit phase — re-evaluate each sub-tile, compose
// its matching 31-byte records contiguously in shared memory and flush them with 16-byte stores.
// At 6.5 TB/s a 512-record tile lasts ~2 ns, far shorter than one L2 round trip; 2-4 K-record units
// keep the number of unresolved predecessors within one or two 128-wide look-back windows.
// Tickets (not blockIdx) order the units, so a unit only ever waits for units held by running CTAs.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxGroup = 8;

struct SelectShared {
  uint32_t warp_cnt[kMaxGroup][kPPT][kBlock / 32];
  unsigned long long out_rec;   // lane.out_base + exclusive prefix of the current unit
  unsigned long long cur_tile;  // direct kernel: ticket broadcast
  alignas(16) uint8_t stage[kTilePts * 31 + 32];
};

template <class Src>
__device__ __forceinline__ void select_count_sub(const ScanParams& P, const Segment& S, const Src& src, uint64_t p0,
                                                 uint32_t npts, uint32_t (*warp_cnt)[kBlock / 32]) {
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    const uint32_t i = (uint32_t)j * kBlock + threadIdx.x;
    Hit h;
    bool m = false;
    if (i < npts) m = src.template eval<false>(S, P.query_kind, P.cls, p0 + i, i, h);
    const uint32_t bal = __ballot_sync(0xffffffffu, m);
    if (lane_id() == 0) warp_cnt[j][warp_id()] = (uint32_t)__popc(bal);
  }
}

// all sub-tile counts of the unit are in shared memory (and synchronised) when this runs;
// returns the number of matches of the sub-tile
template <class Src>
__device__ __forceinline__ uint32_t select_emit_sub(const ScanParams& P, const Segment& S, const LaneDev& L, const Src& src,
                                                    uint64_t p0, uint32_t npts, const uint32_t (*warp_cnt)[kBlock / 32],
                                                    unsigned long long out_rec, SelectShared* sel) {
  const uint32_t tid = threadIdx.x;
  Hit h[kPPT];
  bool m[kPPT];
  uint32_t my_off[kPPT];
  uint32_t total = 0;
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    const uint32_t i = (uint32_t)j * kBlock + tid;
    m[j] = false;
    if (i < npts) m[j] = src.template eval<true>(S, P.query_kind, P.cls, p0 + i, i, h[j]);
    const uint32_t bal = __ballot_sync(0xffffffffu, m[j]);
    my_off[j] = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) {
      const uint32_t c = warp_cnt[j][w];
      if (w == (int)warp_id()) my_off[j] = total + (uint32_t)__popc(bal & ((1u << lane_id()) - 1u));
      total += c;
    }
  }
  if (total == 0) {  // uniform: nothing to write for this sub-tile
    __syncthreads();  // ... but every thread must be done reading the stage before the caller refills it
    return 0;
  }
  const unsigned long long gb0 = out_rec * 31ull;     // first output byte of this sub-tile
  const uint32_t so = (uint32_t)(gb0 & 15ull);        // keep global and shared 16-byte phases equal
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    if (m[j]) {
      const uint32_t i = (uint32_t)j * kBlock + tid;
      uint32_t rgb[3];
      src.colour(S, p0 + i, i, rgb);
      uint32_t w[8];
      point_words(S, h[j], rgb, w);
      sts_point31(sel->stage + so + my_off[j] * 31u, w);
    }
  }
  __syncthreads();
  // records beyond the lane's capacity are counted but not written (host grows the buffer and re-runs)
  const unsigned long long room = out_rec < L.out_cap ? L.out_cap - out_rec : 0ull;
  const uint32_t n_ok = room < (unsigned long long)total ? (uint32_t)room : total;
  const uint32_t nb = n_ok * 31u;
  if (nb) {
    uint8_t* gout = L.out + gb0;  // byte address of stage[so]
    uint32_t head = (16u - so) & 15u;
    if (head > nb) head = nb;
    const uint32_t nvec = (nb - head) >> 4;
    const uint32_t tail0 = head + (nvec << 4);
    if (tid < head) gout[tid] = sel->stage[so + tid];
    const uint4* svec = reinterpret_cast<const uint4*>(sel->stage + so + head);
    uint4* gvec = reinterpret_cast<uint4*>(gout + head);
    for (uint32_t k = tid; k < nvec; k += kBlock) gvec[k] = svec[k];
    if (tid < nb - tail0) gout[tail0 + tid] = sel->stage[so + tail0 + tid];
  }
  __syncthreads();  // the staging buffer is reused by the next sub-tile
  return total;
}

// sum of all sub-tile counts of a unit (uniform across the CTA)
__device__ __forceinline__ uint32_t select_unit_total(const SelectShared* sel, uint32_t n_sub) {
  uint32_t total = 0;
  for (uint32_t j = 0; j < n_sub; ++j)
#pragma unroll
    for (int k = 0; k < kPPT; ++k)
#pragma unroll
      for (int w = 0; w < kBlock / 32; ++w) total += sel->warp_cnt[j][k][w];
  return total;
}

// warp 0: publish the unit's aggregate / prefix, resolve its exclusive prefix, bump the lane's count
__device__ __forceinline__ void select_publish(const ScanParams& P, const Segment& S, const LaneDev& L, uint64_t tile,
                                               uint32_t total, SelectShared* sel) {
  unsigned long long excl = 0;
  const bool first = tile == S.lane_first_tile;
  if (!first) {
    if (lane_id() == 0) st_state(P.tile_state + tile, (kStAgg << kStatusShift) | (unsigned long long)total);
    excl = lookback_exclusive(P.tile_state, tile, S.lane_first_tile);
  }
  if (lane_id() == 0) {
    st_state(P.tile_state + tile, (kStPrefix << kStatusShift) | (excl + (unsigned long long)total));
    sel->out_rec = L.out_base + excl;
    if (total) atomicAdd(L.count, (unsigned long long)total);
  }
}

__global__ void __launch_bounds__(kBlock) k_select_direct(ScanParams P) {
  __shared__ Segment sseg;
  __shared__ SelectShared sel;
  uint32_t seg_i = 0xFFFFFFFFu, seg_cursor = 0;
  DirectSrc src;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sel.cur_tile = atomicAdd(P.ticket, 1ull);
    __syncthreads();
    const uint64_t tile = sel.cur_tile;
    if (tile >= P.n_tiles) break;
    while (seg_cursor + 1 < P.n_segs && tile >= P.segs[seg_cursor + 1].first_tile) ++seg_cursor;
    if (seg_cursor != seg_i) {
      const uint32_t* srcw = reinterpret_cast<const uint32_t*>(P.segs + seg_cursor);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&sseg);
      for (uint32_t k = threadIdx.x; k < sizeof(Segment) / 4; k += kBlock) dstw[k] = srcw[k];
      seg_i = seg_cursor;
      __syncthreads();
    }
    const Segment& S = sseg;
    const LaneDev& L = P.lanes[S.lane];
    const uint64_t u0 = (tile - S.first_tile) * (uint64_t)P.tile_pts;
    const uint64_t rem = S.n_points - u0;
    const uint32_t unit_pts = rem < (uint64_t)P.tile_pts ? (uint32_t)rem : P.tile_pts;
    const uint32_t n_sub = (unit_pts + kTilePts - 1) / kTilePts;
    for (uint32_t j = 0; j < n_sub; ++j) {
      const uint32_t np = min((uint32_t)kTilePts, unit_pts - j * kTilePts);
      select_count_sub(P, S, src, u0 + (uint64_t)j * kTilePts, np, sel.warp_cnt[j]);
    }
    __syncthreads();
    const uint32_t total = select_unit_total(&sel, n_sub);
    if (warp_id() == 0) select_publish(P, S, L, tile, total, &sel);
    __syncthreads();
    unsigned long long run = sel.out_rec;
    if (total == 0) continue;
    for (uint32_t j = 0; j < n_sub; ++j) {
      const uint32_t np = min((uint32_t)kTilePts, unit_pts - j * kTilePts);
      run += select_emit_sub(P, S, L, src, u0 + (uint64_t)j * kTilePts, np, sel.warp_cnt[j], run, &sel);
    }
  }
}

// staged variant: the G sub-tiles of a unit are the G stages of the bulk-copy ring
template <int R, int G>
__global__ void __launch_bounds__(kBlock) k_select_staged(ScanParams P) {
  static_assert(G <= kMaxGroup, "group larger than the count table");
  constexpr uint32_t kSubBytes = (uint32_t)kTilePts * (uint32_t)R;
  extern __shared__ __align__(128) uint8_t dsm[];  // G * kSubBytes
  __shared__ __align__(8) uint64_t full_bar[G];
  __shared__ Segment sseg;
  __shared__ SelectShared sel;
  struct Next {
    unsigned long long tile;  // ~0 = no more units
    const uint8_t* src;       // first record of the unit
    uint32_t npts;
    uint32_t seg;
  };
  __shared__ Next nxt;

  const uint32_t tid = threadIdx.x;
  uint32_t prod_seg = 0;  // thread 0 only

  auto take_ticket = [&]() {
    const unsigned long long tile = atomicAdd(P.ticket, 1ull);
    if (tile >= P.n_tiles) {
      nxt.tile = ~0ull;
      return;
    }
    while (prod_seg + 1 < P.n_segs && tile >= P.segs[prod_seg + 1].first_tile) ++prod_seg;
    const Segment* sg = P.segs + prod_seg;
    const uint64_t u0 = (tile - sg->first_tile) * (uint64_t)P.tile_pts;
    const uint64_t rem = sg->n_points - u0;
    nxt.tile = tile;
    nxt.src = sg->rec + u0 * (uint64_t)R;
    nxt.npts = rem < (uint64_t)P.tile_pts ? (uint32_t)rem : P.tile_pts;
    nxt.seg = prod_seg;
  };
  // load sub-tile j of the NEXT unit into stage j (thread 0)
  auto issue = [&](uint32_t j) {
    if (nxt.tile == ~0ull) return;
    const uint32_t off = j * (uint32_t)kTilePts;
    if (off >= nxt.npts) return;
    const uint32_t n = min((uint32_t)kTilePts, nxt.npts - off);
    const uint32_t bytes = (n * (uint32_t)R + 15u) & ~15u;
    mbar_arrive_expect_tx(&full_bar[j], bytes);
    bulk_copy_g2s(dsm + (size_t)j * kSubBytes, nxt.src + (uint64_t)off * R, bytes, &full_bar[j]);
  };

  if (tid == 0) {
#pragma unroll
    for (int j = 0; j < G; ++j) mbar_init(&full_bar[j], 1u);
    mbar_fence_init();
    take_ticket();
#pragma unroll 1
    for (uint32_t j = 0; j < (uint32_t)G; ++j) issue(j);
  }
  __syncthreads();

  uint32_t parity = 0;  // bit j: parity of the next completion of stage j
  uint32_t seg_i = 0xFFFFFFFFu;
  for (;;) {
    const unsigned long long tile = nxt.tile;
    if (tile == ~0ull) break;
    const uint32_t unit_pts = nxt.npts;
    const uint32_t seg_now = nxt.seg;
    if (seg_now != seg_i) {
      const uint32_t* srcw = reinterpret_cast<const uint32_t*>(P.segs + seg_now);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&sseg);
      for (uint32_t k = tid; k < sizeof(Segment) / 4; k += kBlock) dstw[k] = srcw[k];
      seg_i = seg_now;
    }
    __syncthreads();  // everyone has read nxt (thread 0 overwrites it below); sseg is complete
    const Segment& S = sseg;
    const LaneDev& L = P.lanes[S.lane];
    const uint64_t u0 = (tile - S.first_tile) * (uint64_t)P.tile_pts;
    const uint32_t n_sub = (unit_pts + kTilePts - 1) / kTilePts;

    // ---- count phase: sub-tiles are evaluated as their bulk copies land ----
    for (uint32_t j = 0; j < n_sub; ++j) {
      mbar_wait(&full_bar[j], (parity >> j) & 1u);
      parity ^= 1u << j;
      const uint32_t np = min((uint32_t)kTilePts, unit_pts - j * kTilePts);
      SmemSrc<R> src{dsm + (size_t)j * kSubBytes};
      select_count_sub(P, S, src, u0 + (uint64_t)j * kTilePts, np, sel.warp_cnt[j]);
    }
    __syncthreads();
    const uint32_t total = select_unit_total(&sel, n_sub);
    if (warp_id() == 0) {
      if (tid == 0) {
        take_ticket();  // before the look-back: stages this unit does not use can start loading now
        for (uint32_t j = n_sub; j < (uint32_t)G; ++j) issue(j);
      }
      __syncwarp();
      select_publish(P, S, L, tile, total, &sel);
    }
    __syncthreads();

    // ---- emit phase: each stage is refilled with the next unit's sub-tile as soon as it is drained ----
    unsigned long long run = sel.out_rec;
    for (uint32_t j = 0; j < n_sub; ++j) {
      if (total) {
        const uint32_t np = min((uint32_t)kTilePts, unit_pts - j * kTilePts);
        SmemSrc<R> src{dsm + (size_t)j * kSubBytes};
        run += select_emit_sub(P, S, L, src, u0 + (uint64_t)j * kTilePts, np, sel.warp_cnt[j], run, &sel);
      }
      // select_emit_sub ends with a barrier whenever it read the stage; when it returned early (no match in
      // the sub-tile) the count phase barrier is the last reader of stage j
      if (tid == 0) issue(j);
    }
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
__global__ void k_grid_prune(GridDev g, uint64_t n_in, Candidate* dst, unsigned long long* dst_count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n_in; i0 += stride) {
    const uint64_t i = i0 + threadIdx.x;
    bool keep = false;
    uint4 c0, c1, c2, c3;
    if (i < n_in) {
      const uint4* c4 = reinterpret_cast<const uint4*>(g.cands + i);
      c0 = c4[0];
      const uint64_t key = (uint64_t)c0.x | ((uint64_t)c0.y << 32);
      const unsigned long long d = (unsigned long long)c0.z | ((unsigned long long)c0.w << 32);
      const uint64_t slot = grid_slot(g, key, false);
      keep = slot != ~0ull && g.table[slot] == d;
      if (keep) {
        c1 = c4[1];
        c2 = c4[2];
        c3 = c4[3];
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (bal == 0u) continue;
    const uint32_t leader = (uint32_t)__ffs((int)bal) - 1u;
    unsigned long long base = 0;
    if (lane_id() == leader) base = atomicAdd(dst_count, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, (int)leader);
    if (keep) {
      uint4* o4 = reinterpret_cast<uint4*>(dst + base + __popc(bal & ((1u << lane_id()) - 1u)));
      o4[0] = c0;
      o4[1] = c1;
      o4[2] = c2;
      o4[3] = c3;
    }
  }
}

// among the candidates at their cell's minimum distance, the smallest scan index wins
__global__ void k_grid_min_index(GridDev g, uint64_t n, unsigned long long* idx_table) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Candidate& c = g.cands[i];
    const uint64_t slot = grid_slot(g, c.key, false);
    if (slot != ~0ull && g.table[slot] == c.dist_bits) atomicMin(idx_table + slot, (unsigned long long)c.scan_idx);
  }
}

// winners -> 31-byte records (order arbitrary, like HashMap::values) or -> per-owner candidate parts
// mode 0: count per part, mode 1: write candidates into parts, mode 2: write 31-byte points
__global__ void k_grid_emit(GridDev g, uint64_t n, unsigned long long* idx_table, int mode, uint32_t n_parts,
                            unsigned long long* part_counts, unsigned long long* part_cursor, Candidate* out_cands,
                            uint8_t* out_points, unsigned long long* out_count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Candidate& c = g.cands[i];
    const uint64_t slot = grid_slot(g, c.key, false);
    if (slot == ~0ull || g.table[slot] != c.dist_bits) continue;
    if (mode == 0) {
      // count the winner once even if the candidate list holds duplicates of it
      if (idx_table[slot] != c.scan_idx) continue;
      if (atomicCAS(idx_table + slot, (unsigned long long)c.scan_idx, (unsigned long long)c.scan_idx | (1ull << 63)) !=
          (unsigned long long)c.scan_idx)
        continue;
      atomicAdd(part_counts + (uint32_t)(mix64(c.key) % n_parts), 1ull);
    } else if (mode == 1) {
      // second walk after mode 0: claimed entries carry bit 63; release the claim while emitting
      if (idx_table[slot] != ((unsigned long long)c.scan_idx | (1ull << 63))) continue;
      if (atomicCAS(idx_table + slot, (unsigned long long)c.scan_idx | (1ull << 63), (unsigned long long)c.scan_idx) !=
          ((unsigned long long)c.scan_idx | (1ull << 63)))
        continue;
      const uint32_t part = (uint32_t)(mix64(c.key) % n_parts);
      const unsigned long long o = atomicAdd(part_cursor + part, 1ull);
      const uint4* s4 = reinterpret_cast<const uint4*>(&c);
      uint4* o4 = reinterpret_cast<uint4*>(out_cands + o);
      o4[0] = s4[0];
      o4[1] = s4[1];
      o4[2] = s4[2];
      o4[3] = s4[3];
    } else {
      if (idx_table[slot] != c.scan_idx) continue;
      if (atomicCAS(idx_table + slot, (unsigned long long)c.scan_idx, (unsigned long long)c.scan_idx | (1ull << 63)) !=
          (unsigned long long)c.scan_idx)
        continue;
      const unsigned long long o = atomicAdd(out_count, 1ull);
      uint8_t* dst = out_points + o * 31ull;
#pragma unroll
      for (int b = 0; b < 31; ++b) dst[b] = c.point[b];
    }
  }
}

// fold candidates received from peers into the table (multi-GPU density merge)
__global__ void k_grid_import(GridDev g, const Candidate* in, uint64_t n) {
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
    const uint32_t bal = __ballot_sync(0xffffffffu, want);
    if (bal == 0u) continue;
    const uint32_t leader = (uint32_t)__ffs((int)bal) - 1u;
    unsigned long long base = 0;
    if (lane_id() == leader) base = atomicAdd(g.cand_count, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, (int)leader);
    if (want) {
      const unsigned long long ci = base + (unsigned long long)__popc(bal & ((1u << lane_id()) - 1u));
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
// group size (= ring depth) of the select kernels: ~50-70 KB ring, two CTAs per SM
template <int R>
struct SelectGroup {
  static constexpr int value = R <= 12 ? 8 : (R <= 20 ? 6 : (R <= 28 ? 5 : 4));
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
  constexpr int STAGES = ScanStages<R>::value;
  constexpr size_t smem = (size_t)STAGES * kTilePts * R;
  static bool configured = false;
  auto kfn = k_scan_staged<R, MODE, STAGES>;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  unsigned grid = 0;
  if (persistent_grid((const void*)kfn, smem, sm_count, p.n_tiles, 4, &grid) != 0) return -1;
  if (grid == 0) return 0;
  kfn<<<grid, kBlock, smem, st>>>(p);
  return check_launch();
}

template <int R>
static int launch_select_staged_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  constexpr int G = SelectGroup<R>::value;
  constexpr size_t smem = (size_t)G * kTilePts * R;
  static bool configured = false;
  auto kfn = k_select_staged<R, G>;
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  unsigned grid = 0;
  if (persistent_grid((const void*)kfn, smem, sm_count, p.n_tiles, 3, &grid) != 0) return -1;
  if (grid == 0) return 0;
  kfn<<<grid, kBlock, smem, st>>>(p);
  return check_launch();
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

static int launch_select_staged_r(const ScanParams& p, uint32_t R, int sm_count, cudaStream_t st) {
  switch (R) {
    case 12: return launch_select_staged_t<12>(p, sm_count, st);
    case 20: return launch_select_staged_t<20>(p, sm_count, st);
    case 26: return launch_select_staged_t<26>(p, sm_count, st);
    case 28: return launch_select_staged_t<28>(p, sm_count, st);
    case 34: return launch_select_staged_t<34>(p, sm_count, st);
    default: return 1;
  }
}

template <int MODE>
static int launch_direct_t(const ScanParams& p, int sm_count, cudaStream_t st) {
  auto kfn = k_scan_direct<MODE>;
  unsigned grid = 0;
  if (persistent_grid((const void*)kfn, 0, sm_count, p.n_tiles, 8, &grid) != 0) return -1;
  if (grid == 0) return 0;
  kfn<<<grid, kBlock, 0, st>>>(p);
  return check_launch();
}

static int launch_select_direct(const ScanParams& p, int sm_count, cudaStream_t st) {
  unsigned grid = 0;
  if (persistent_grid((const void*)k_select_direct, 0, sm_count, p.n_tiles, 8, &grid) != 0) return -1;
  if (grid == 0) return 0;
  k_select_direct<<<grid, kBlock, 0, st>>>(p);
  return check_launch();
}

bool staged_supports(uint32_t R) { return R == 12 || R == 20 || R == 26 || R == 28 || R == 34; }

// records per scheduling unit ("tile") for a launch: 512 for count / grid, a whole group for select
uint32_t tile_points(int variant, int mode, uint32_t R) {
  if (mode != MODE_SELECT) return kTilePts;
  if (variant == 2 && staged_supports(R)) {
    switch (R) {
      case 12: return kTilePts * SelectGroup<12>::value;
      case 20: return kTilePts * SelectGroup<20>::value;
      case 26: return kTilePts * SelectGroup<26>::value;
      case 28: return kTilePts * SelectGroup<28>::value;
      default: return kTilePts * SelectGroup<34>::value;
    }
  }
  return kTilePts * kMaxGroup;
}

// variant: 1 = direct, 2 = staged (needs uniform_record_len supported); returns 0 ok, <0 CUDA error
int launch_scan(int variant, int mode, const ScanParams& p, uint32_t uniform_record_len, int sm_count, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == 2 && staged_supports(uniform_record_len)) {
    int rc = 1;
    if (mode == MODE_COUNT) rc = launch_staged_r<MODE_COUNT>(p, uniform_record_len, sm_count, st);
    if (mode == MODE_SELECT) rc = launch_select_staged_r(p, uniform_record_len, sm_count, st);
    if (mode == MODE_GRID) rc = launch_staged_r<MODE_GRID>(p, uniform_record_len, sm_count, st);
    if (rc <= 0) return rc;
  }
  if (mode == MODE_COUNT) return launch_direct_t<MODE_COUNT>(p, sm_count, st);
  if (mode == MODE_SELECT) return launch_select_direct(p, sm_count, st);
  return launch_direct_t<MODE_GRID>(p, sm_count, st);
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

int launch_grid_min_index(const GridDev& g, uint64_t n, unsigned long long* idx_table, int sm_count, void* stream) {
  if (n == 0) return 0;
  k_grid_min_index<<<grid_for(n, sm_count), 256, 0, (cudaStream_t)stream>>>(g, n, idx_table);
  return check_launch();
}

int launch_grid_emit(const GridDev& g, uint64_t n, unsigned long long* idx_table, int mode, uint32_t n_parts,
                     unsigned long long* part_counts, unsigned long long* part_cursor, Candidate* out_cands,
                     uint8_t* out_points, unsigned long long* out_count, int sm_count, void* stream) {
  if (n == 0) return 0;
  k_grid_emit<<<grid_for(n, sm_count), 256, 0, (cudaStream_t)stream>>>(g, n, idx_table, mode, n_parts, part_counts,
                                                                         part_cursor, out_cands, out_points, out_count);
  return check_launch();
}

int launch_grid_import(const GridDev& g, const Candidate* in, uint64_t n, int sm_count, void* stream) {
  if (n == 0) return 0;
  k_grid_import<<<grid_for(n, sm_count), 256, 0, (cudaStream_t)stream>>>(g, in, n);
  return check_launch();
}

}  // namespace pcq
