// Build-time probe (tests/test_no_fma.py): the shared arithmetic of the path, one function per kernel, compiled with the
// library's own flags.  The reference never contracts a * b + c (las.rs:139-141, grid_sampling.rs:78-95); the SASS of
// probe_reconstruct and probe_dist2 must therefore hold no FMA.  probe_cells holds the division (IEEE quotients,
// computed with FMAs by design — grid_math.cuh) and is listed for contrast only.
#include "grid_math.cuh"

using namespace pcq;

extern "C" __global__ void probe_reconstruct(const int* v, double s, double o, double* out) {
  out[threadIdx.x] = reconstruct(v[threadIdx.x], s, o);
}
extern "C" __global__ void probe_dist2(GridDev g, const unsigned long long* c, const double* p, double* out) {
  const uint64_t cc[3] = {c[0], c[1], c[2]};
  out[threadIdx.x] = grid_dist2(g, cc, p[0], p[1], p[2]);
}
extern "C" __global__ void probe_cells(GridDev g, const double* p, unsigned long long* out) {
  uint64_t c[3];
  grid_cells(g, p[0], p[1], p[2], c);
  out[0] = c[0];
  out[1] = c[1];
  out[2] = c[2];
}
