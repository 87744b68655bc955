// grid_math.cuh — device functions of SparseGrid::insert_point (grid_sampling.rs:49-105) shared by the scan kernels
// (kernels.cu) and the alias replay (alias.cu).  Every operation that feeds a stored double is an explicit
// round-to-nearest intrinsic in the reference's order; the library is compiled with --fmad=false.
#pragma once
#include <cstdint>

#include "pcq_device.h"

namespace pcq {

constexpr unsigned long long kCandEmpty = ~0ull;  // scan_idx of an unused arena slot

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

// (v as f64 * scale) + offset, las.rs:139-141 — two roundings, never an FMA
__device__ __forceinline__ double reconstruct(int32_t v, double scale, double offset) {
  return __dadd_rn(__dmul_rn((double)v, scale), offset);
}

struct CellEval {
  uint64_t key;
  unsigned long long dist_bits;
  bool aliased;
};

// n / d for a denominator d whose correctly rounded reciprocal y = RN(1 / d) the host has computed with an IEEE
// division: two Markstein corrections.  q0 = RN(n y) is within 2 ulp of n / d; the residual r0 = n - d q0 is exact in
// an FMA, so q1 = RN(q0 + r0 y) is a faithful quotient (error below 2^-105 before its rounding); a second exact
// residual and correction then give the correctly rounded quotient (Markstein 1990, Theorem 2: y correctly rounded,
// q faithful, r exact => RN(q + r y) = RN(n / d)).  Five FP64 instructions instead of the ~25 of the general
// division sequence.  Valid while nothing over- or underflows: the caller admits |n| in [2^-500, 2^500) or n == 0,
// the host admits d in [2^-500, 2^500] (GridDev::fast_div); everything else takes __ddiv_rn.  (The FMAs are part of
// the division algorithm itself — every result is the IEEE quotient, which is what Rust's `/` returns.)
__device__ __forceinline__ double div_by_reciprocal(double n, double d, double y) {
  const double q0 = __dmul_rn(n, y);
  const double r0 = __fma_rn(-d, q0, n);
  const double q1 = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-d, q1, n);
  return __fma_rn(r1, y, q1);
}
__device__ __forceinline__ bool div_fast_numerator(double n) {
  const uint32_t e = ((uint32_t)__double2hiint(n) >> 20) & 0x7FFu;  // biased exponent
  return (e - 523u) < 1000u || n == 0.0;                            // 2^-500 <= |n| < 2^500, or +-0
}
// `f64 as u64` of a finite value: the conversion instruction itself truncates toward zero and clamps to the range of
// the destination (PTX cvt: "for float-to-integer conversions the result is clamped to the destination range").
__device__ __forceinline__ uint64_t f64_as_u64_finite(double v) {
  unsigned long long r;
  asm("cvt.rzi.u64.f64 %0, %1;" : "=l"(r) : "d"(v));
  return r;
}

// :51-60  unmasked cell indices of a position
__device__ __forceinline__ void grid_cells(const GridDev& g, double px, double py, double pz, uint64_t c[3]) {
  // r = (p - min) * dims as f64 / (max - min);  cell = r as u64
  const double nx = __dmul_rn(__dsub_rn(px, g.bmin[0]), g.dims_f[0]);
  const double ny = __dmul_rn(__dsub_rn(py, g.bmin[1]), g.dims_f[1]);
  const double nz = __dmul_rn(__dsub_rn(pz, g.bmin[2]), g.dims_f[2]);
  if (g.fast_div && div_fast_numerator(nx) && div_fast_numerator(ny) && div_fast_numerator(nz)) {
    // (a zero numerator gives a zero quotient here as well: cell 0)
    c[0] = f64_as_u64_finite(div_by_reciprocal(nx, g.ext[0], g.inv_ext[0]));
    c[1] = f64_as_u64_finite(div_by_reciprocal(ny, g.ext[1], g.inv_ext[1]));
    c[2] = f64_as_u64_finite(div_by_reciprocal(nz, g.ext[2], g.inv_ext[2]));
    return;
  }
  // A zero numerator (a point exactly on a minimum face) would send the whole warp through the slow path of the IEEE
  // division; 0 / d is 0 or NaN and both cast to cell 0, so such lanes divide 1.0 instead and ignore the quotient.
  const double rx = __ddiv_rn(nx == 0.0 ? 1.0 : nx, g.ext[0]);
  const double ry = __ddiv_rn(ny == 0.0 ? 1.0 : ny, g.ext[1]);
  const double rz = __ddiv_rn(nz == 0.0 ? 1.0 : nz, g.ext[2]);
  c[0] = nx == 0.0 ? 0ull : f64_as_u64(rx);
  c[1] = ny == 0.0 ? 0ull : f64_as_u64(ry);
  c[2] = nz == 0.0 ? 0ull : f64_as_u64(rz);
}

// :78-95  squared distance of a position to the centre of the (unmasked) cell c
__device__ __forceinline__ double grid_dist2(const GridDev& g, const uint64_t c[3], double px, double py, double pz) {
  // centre = (cell as f64 + 0.5) * cell_size + min
  const double ccx = __dadd_rn(__dmul_rn(__dadd_rn(__ull2double_rn(c[0]), 0.5), g.cell_size), g.bmin[0]);
  const double ccy = __dadd_rn(__dmul_rn(__dadd_rn(__ull2double_rn(c[1]), 0.5), g.cell_size), g.bmin[1]);
  const double ccz = __dadd_rn(__dmul_rn(__dadd_rn(__ull2double_rn(c[2]), 0.5), g.cell_size), g.bmin[2]);
  // distance_squared = (dx*dx + dy*dy) + dz*dz  (nalgebra 0.23, no FMA)
  const double dx = __dsub_rn(ccx, px), dy = __dsub_rn(ccy, py), dz = __dsub_rn(ccz, pz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ CellEval grid_eval(const GridDev& g, double px, double py, double pz) {
  uint64_t c[3];
  grid_cells(g, px, py, pz, c);
  CellEval e;
  // a cell above its mask aliases a low cell while its centre lies elsewhere (:62-70 vs :78-82)
  e.aliased = (c[0] > g.mask[0]) | (c[1] > g.mask[1]) | (c[2] > g.mask[2]);
  e.key = (c[0] & g.mask[0]) | ((c[1] & g.mask[1]) << g.shift_y) | ((c[2] & g.mask[2]) << g.shift_z);
  const double d = grid_dist2(g, c, px, py, pz);
  e.dist_bits = (unsigned long long)__double_as_longlong(d);  // d >= +0: bit order == value order
  return e;
}

// affected-key set lookup: ordinal of `key`, or ~0u
__device__ __forceinline__ uint32_t alias_find(const GridDev& g, uint64_t key) {
  if (g.alias_slots == 0) return ~0u;
  const uint64_t mask = g.alias_slots - 1ull;
  uint64_t s = mix64(key) & mask;
  for (uint64_t probe = 0; probe < g.alias_slots; ++probe) {
    const unsigned long long cur = g.alias_keys[s];
    if (cur == key) return g.alias_ord[s];
    if (cur == ~0ull) return ~0u;
    s = (s + 1ull) & mask;
  }
  return ~0u;
}

}  // namespace pcq
