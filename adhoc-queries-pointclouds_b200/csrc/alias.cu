// alias.cu — exact replay of SparseGrid keys that suffer key aliasing.
//
// SparseGrid::insert_point (query/src/grid_sampling.rs:49-105) masks each cell index to its bit width when it
// builds the HashMap key (:62-70) but computes the cell centre from the UNMASKED index (:78-82).  A point whose
// cell index exceeds its mask on some axis (possible when the axis has a power-of-two number of cells and the point
// lies on the inclusive max face of the grid, e.g. z = 200.00 in the doc-S / doc-L grids at 25 m) therefore shares
// the key of a low cell while it is compared against a different centre, and the reference's result for that key
// becomes the outcome of a SEQUENTIAL fold in scan order — not an argmin.  The scan kernels keep every point of
// such an "affected" key out of the atomic-min table and write it to a replay log; the routines here order the log
// by (key, scan index) and run the reference's fold, one affected key per thread.  Affected keys are rare (a
// handful of columns under the max face), so none of this is on the hot path.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include "grid_math.cuh"
#include "pcq_device.h"

namespace pcq {

namespace {

struct PointView {
  double x, y, z;
};
__device__ __forceinline__ PointView pos_of(const uint8_t* p31) {  // 8-byte aligned in Candidate and AliasState
  const double* d = reinterpret_cast<const double*>(p31);
  return PointView{d[0], d[1], d[2]};
}

// ---- the fold state a newly affected key starts from: HashMap entry after all EARLIER launches ----------------
// All earlier points of such a key were un-aliased (else the key would have been affected before), so the entry is
// the argmin of (distance, scan index) over them; that candidate is still in the arena.
__global__ void k_alias_pre_dist(GridDev g, uint64_t n, uint32_t ord0, unsigned long long before, unsigned long long* best_dist) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Candidate& c = g.cands[i];
    if (c.scan_idx == kCandEmpty || c.scan_idx >= before) continue;
    const uint32_t ord = alias_find(g, c.key);
    if (ord == ~0u || ord < ord0) continue;
    atomicMin(best_dist + (ord - ord0), (unsigned long long)c.dist_bits);
  }
}
__global__ void k_alias_pre_idx(GridDev g, uint64_t n, uint32_t ord0, unsigned long long before,
                                const unsigned long long* best_dist, unsigned long long* best_idx) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Candidate& c = g.cands[i];
    if (c.scan_idx == kCandEmpty || c.scan_idx >= before) continue;
    const uint32_t ord = alias_find(g, c.key);
    if (ord == ~0u || ord < ord0) continue;
    if (c.dist_bits == best_dist[ord - ord0]) atomicMin(best_idx + (ord - ord0), (unsigned long long)c.scan_idx);
  }
}
__global__ void k_alias_pre_take(GridDev g, uint64_t n, uint32_t ord0, unsigned long long before,
                                 const unsigned long long* best_dist, const unsigned long long* best_idx, AliasState* states) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Candidate& c = g.cands[i];
    if (c.scan_idx == kCandEmpty || c.scan_idx >= before) continue;
    const uint32_t ord = alias_find(g, c.key);
    if (ord == ~0u || ord < ord0) continue;
    if (c.dist_bits != best_dist[ord - ord0] || c.scan_idx != best_idx[ord - ord0]) continue;
    AliasState& s = states[ord];  // duplicates of the winner write identical bytes
#pragma unroll
    for (int b = 0; b < 31; ++b) s.point[b] = c.point[b];
    s.valid = 1;
  }
}

// ---- ordering the log -------------------------------------------------------------------------------------------
__global__ void k_alias_keys_scan(const Candidate* log, uint32_t n, unsigned long long* keys, uint32_t* vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = log[i].scan_idx;
    vals[i] = i;
  }
}
__global__ void k_alias_keys_cell(const Candidate* log, uint32_t n, const uint32_t* order, unsigned long long* keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = log[order[i]].key;
}

// ---- the reference's fold, one affected key per thread (grid_sampling.rs:72-102) -----------------------------
__global__ void k_alias_fold(GridDev g, uint32_t n, const unsigned long long* keys /* sorted */, const uint32_t* order,
                             AliasState* states) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  if (i > 0 && keys[i - 1] == key) return;  // not the head of a run
  const uint32_t ord = alias_find(g, key);
  if (ord == ~0u) return;  // cannot happen: every logged key is in the set by now
  AliasState st = states[ord];
  for (uint32_t j = i; j < n && keys[j] == key; ++j) {
    const Candidate& c = g.log[order[j]];
    bool take;
    if (!st.valid) {
      take = true;  // :73-76  None => insert
    } else {
      const PointView pn = pos_of(c.point);
      const PointView pc = pos_of(st.point);
      uint64_t cell[3];
      grid_cells(g, pn.x, pn.y, pn.z, cell);  // the centre is that of the NEW point's (unmasked) cell, :78-82
      const double cur = grid_dist2(g, cell, pc.x, pc.y, pc.z);
      const double nw = grid_dist2(g, cell, pn.x, pn.y, pn.z);
      take = nw < cur;  // :97  strictly smaller replaces
    }
    if (take) {
#pragma unroll
      for (int b = 0; b < 31; ++b) st.point[b] = c.point[b];
      st.valid = 1;
    }
  }
  states[ord] = st;
}

unsigned blocks_for(uint64_t n, int sm_count) {
  uint64_t b = (n + 255) / 256;
  const uint64_t cap = (uint64_t)sm_count * 8;
  if (b > cap) b = cap;
  return (unsigned)(b ? b : 1);
}

}  // namespace

int alias_prewinners(const GridDev& g, uint64_t n_cands, const unsigned long long* /*d_keys*/, uint32_t n_keys,
                     unsigned long long before_scan, AliasState* d_states, uint32_t ord0, int sm_count, void* stream) {
  if (n_keys == 0 || n_cands == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* scratch = nullptr;
  if (cudaMalloc(&scratch, 2ull * n_keys * sizeof(unsigned long long)) != cudaSuccess) return -1;
  cudaMemsetAsync(scratch, 0xFF, 2ull * n_keys * sizeof(unsigned long long), st);
  const unsigned b = blocks_for(n_cands, sm_count);
  k_alias_pre_dist<<<b, 256, 0, st>>>(g, n_cands, ord0, before_scan, scratch);
  k_alias_pre_idx<<<b, 256, 0, st>>>(g, n_cands, ord0, before_scan, scratch, scratch + n_keys);
  k_alias_pre_take<<<b, 256, 0, st>>>(g, n_cands, ord0, before_scan, scratch, scratch + n_keys, d_states);
  const cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(scratch);
  return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : -1;
}

int alias_replay(const GridDev& g, uint64_t n64, AliasState* d_states, int /*sm_count*/, void* stream) {
  if (n64 == 0) return 0;
  if (n64 > 0x7FFFFFFFull) return -2;
  const uint32_t n = (uint32_t)n64;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long *k0 = nullptr, *k1 = nullptr;
  uint32_t *v0 = nullptr, *v1 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = -1;
  do {
    if (cudaMalloc(&k0, n * 8ull) != cudaSuccess || cudaMalloc(&k1, n * 8ull) != cudaSuccess ||
        cudaMalloc(&v0, n * 4ull) != cudaSuccess || cudaMalloc(&v1, n * 4ull) != cudaSuccess)
      break;
    if (cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, (int)n, 0, 64, st) != cudaSuccess) break;
    if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1) != cudaSuccess) break;
    const unsigned b = (n + 255u) / 256u;
    // 1. by scan index
    k_alias_keys_scan<<<b, 256, 0, st>>>(g.log, n, k0, v0);
    if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (int)n, 0, 64, st) != cudaSuccess) break;
    // 2. stable by key: (key, scan index) order
    k_alias_keys_cell<<<b, 256, 0, st>>>(g.log, n, v1, k0);
    if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v1, v0, (int)n, 0, 64, st) != cudaSuccess) break;
    // 3. fold
    k_alias_fold<<<b, 256, 0, st>>>(g, n, k1, v0, d_states);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) break;
    rc = 0;
  } while (false);
  cudaFree(k0);
  cudaFree(k1);
  cudaFree(v0);
  cudaFree(v1);
  cudaFree(tmp);
  return rc;
}

}  // namespace pcq
