// index.cu — the on-the-fly chunk index (improvements.md:3-10): one 64-byte header per PCQ_INDEX_CHUNK_POINTS points
// of a resident file, holding the integer AABB of the raw x/y/z fields and the set of class bytes that occur.
//
// The build is one streaming pass over what the scans read anyway (LAS: every record; LAST: the position and the
// class column), so it is HBM-bound like them: N * record_len (LAS) or N * 13 (LAST) bytes in, 64 bytes per chunk out.
// One CTA owns a chunk at a time: 256 threads x 32 points, per-thread min/max, warp `redux`, shared-memory atomics
// across the 8 warps, one 64-byte header store.  The class set lives in shared memory: a lane only issues the atomic
// OR when the bit is not there yet, which after the first few points of a chunk is never.
#include <cuda_runtime.h>

#include <climits>

#include "pcq_device.h"

namespace pcq {
namespace {

constexpr uint32_t kChunk = PCQ_INDEX_CHUNK_POINTS;
constexpr uint32_t kIdxBlock = 256;
static_assert(kChunk % kIdxBlock == 0, "a full chunk is a whole number of rounds");

__device__ __forceinline__ int32_t field_i32(const uint8_t* p, int al) {
  if (al == 4) return __ldg(reinterpret_cast<const int32_t*>(p));
  if (al == 2) {
    const uint32_t a = __ldg(reinterpret_cast<const uint16_t*>(p));
    const uint32_t b = __ldg(reinterpret_cast<const uint16_t*>(p + 2));
    return (int32_t)(a | (b << 16));
  }
  return (int32_t)((uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) |
                   ((uint32_t)__ldg(p + 3) << 24));
}

struct Extent {
  int32_t lo[3], hi[3];
  __device__ __forceinline__ void add(int32_t x, int32_t y, int32_t z) {
    lo[0] = min(lo[0], x), hi[0] = max(hi[0], x);
    lo[1] = min(lo[1], y), hi[1] = max(hi[1], y);
    lo[2] = min(lo[2], z), hi[2] = max(hi[2], z);
  }
};

__device__ __forceinline__ void note_class(uint32_t* s_bits, uint32_t c) {
  const uint32_t w = c >> 5, b = 1u << (c & 31u);
  if ((*reinterpret_cast<volatile uint32_t*>(s_bits + w) & b) == 0u) atomicOr(s_bits + w, b);
}

__global__ void __launch_bounds__(kIdxBlock) k_chunk_index(ChunkIndexArgs A, pcq_chunk_header* __restrict__ out) {
  __shared__ int32_t s_lo[3], s_hi[3];
  __shared__ uint32_t s_bits[8];
  const uint32_t tid = threadIdx.x;
  const int al = A.align;

  for (uint64_t chunk = blockIdx.x; chunk < A.n_chunks; chunk += gridDim.x) {
    if (tid < 8) s_bits[tid] = 0u;
    if (tid < 3) s_lo[tid] = INT_MAX, s_hi[tid] = INT_MIN;
    __syncthreads();

    const uint64_t p0 = chunk * (uint64_t)kChunk;
    const uint64_t rem = A.n_points - p0;
    const uint32_t n = rem < (uint64_t)kChunk ? (uint32_t)rem : kChunk;
    Extent e{{INT_MAX, INT_MAX, INT_MAX}, {INT_MIN, INT_MIN, INT_MIN}};

    if (A.layout == PCQ_LAYOUT_LAS) {
      const uint8_t* base = A.rec + p0 * (uint64_t)A.record_len;
      if (n == kChunk) {
#pragma unroll 8
        for (uint32_t j = 0; j < kChunk / kIdxBlock; ++j) {
          const uint8_t* p = base + (uint64_t)(j * kIdxBlock + tid) * A.record_len;
          e.add(field_i32(p, al), field_i32(p + 4, al), field_i32(p + 8, al));
          note_class(s_bits, __ldg(p + A.cls_off));
        }
      } else {
        for (uint32_t i = tid; i < n; i += kIdxBlock) {
          const uint8_t* p = base + (uint64_t)i * A.record_len;
          e.add(field_i32(p, al), field_i32(p + 4, al), field_i32(p + 8, al));
          note_class(s_bits, __ldg(p + A.cls_off));
        }
      }
    } else {
      const uint8_t* base = A.rec + p0 * 12ull;
      const uint8_t* cls = A.cls + p0;
      if (A.parts & kIndexPartBox) {
        if (n == kChunk) {
#pragma unroll 8
          for (uint32_t j = 0; j < kChunk / kIdxBlock; ++j) {
            const uint8_t* p = base + (uint64_t)(j * kIdxBlock + tid) * 12u;
            e.add(field_i32(p, al), field_i32(p + 4, al), field_i32(p + 8, al));
          }
        } else {
          for (uint32_t i = tid; i < n; i += kIdxBlock) {
            const uint8_t* p = base + (uint64_t)i * 12u;
            e.add(field_i32(p, al), field_i32(p + 4, al), field_i32(p + 8, al));
          }
        }
      }
      if (A.parts & kIndexPartCls) {
        if (n == kChunk && (reinterpret_cast<uintptr_t>(cls) & 15u) == 0) {
          // class column of a full chunk: 512 16-byte words, two per thread
#pragma unroll
          for (uint32_t j = 0; j < kChunk / 16u / kIdxBlock; ++j) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(cls) + j * kIdxBlock + tid);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // equal neighbours are the common case in real class columns: one probe per distinct byte of the word
              const uint32_t b0 = w[k] & 0xFFu, b1 = (w[k] >> 8) & 0xFFu, b2 = (w[k] >> 16) & 0xFFu, b3 = w[k] >> 24;
              note_class(s_bits, b0);
              if (b1 != b0) note_class(s_bits, b1);
              if (b2 != b1) note_class(s_bits, b2);
              if (b3 != b2) note_class(s_bits, b3);
            }
          }
        } else {
          for (uint32_t i = tid; i < n; i += kIdxBlock) note_class(s_bits, __ldg(cls + i));
        }
      }
    }

#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int32_t lo = __reduce_min_sync(0xffffffffu, e.lo[a]);
      const int32_t hi = __reduce_max_sync(0xffffffffu, e.hi[a]);
      if ((tid & 31u) == 0) {
        atomicMin(&s_lo[a], lo);
        atomicMax(&s_hi[a], hi);
      }
    }
    __syncthreads();
    if (tid == 0 && A.parts != (kIndexPartBox | kIndexPartCls)) {
      // the part this pass has no column for: nothing is known, so nothing may be excluded
      if (!(A.parts & kIndexPartBox)) {
#pragma unroll
        for (int a = 0; a < 3; ++a) s_lo[a] = INT_MIN, s_hi[a] = INT_MAX;
      }
      if (!(A.parts & kIndexPartCls)) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s_bits[k] = 0xFFFFFFFFu;
      }
    }
    __syncthreads();
    if (tid < 4) {
      // 64-byte header as four 16-byte stores
      uint4 v;
      if (tid == 0) v = make_uint4((uint32_t)s_lo[0], (uint32_t)s_lo[1], (uint32_t)s_lo[2], (uint32_t)s_hi[0]);
      if (tid == 1) v = make_uint4((uint32_t)s_hi[1], (uint32_t)s_hi[2], s_bits[0], s_bits[1]);
      if (tid == 2) v = make_uint4(s_bits[2], s_bits[3], s_bits[4], s_bits[5]);
      if (tid == 3) v = make_uint4(s_bits[6], s_bits[7], n, 0u);
      reinterpret_cast<uint4*>(out + chunk)[tid] = v;
    }
    __syncthreads();
  }
}

}  // namespace

int launch_chunk_index(const ChunkIndexArgs& a, pcq_chunk_header* out, int sm_count, void* stream) {
  static_assert(sizeof(pcq_chunk_header) == 64, "chunk header must be 64 bytes");
  if (a.n_chunks == 0) return 0;
  const uint64_t want = (uint64_t)sm_count * 8u;  // 8 CTAs of 256 threads fill an SM
  const unsigned grid = (unsigned)(a.n_chunks < want ? a.n_chunks : want);
  k_chunk_index<<<grid, kIdxBlock, 0, static_cast<cudaStream_t>(stream)>>>(a, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace pcq
