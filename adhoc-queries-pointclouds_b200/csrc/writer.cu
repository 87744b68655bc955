// writer.cu — the `-o` output of the query (dump_points.rs:63-116) prepared on the device: the 31-byte Points a
// BUFFER or GRID collector holds are turned into LAS 1.2 point format 2 records (26 bytes) by two kernels — a min / max
// reduction of the positions (offset = min position, :74-80) and the quantisation round((p - offset) / scale) — so
// that 26 instead of 31 bytes per record cross PCIe and no host loop touches a point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>

#include "pcq_internal.hpp"

using namespace pcq;

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) return fail(PCQ_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
#define RC(call)             \
  do {                       \
    int rc_ = (call);        \
    if (rc_ != PCQ_OK) return rc_; \
  } while (0)

namespace {

constexpr int kWrThreads = 256;
constexpr int kWrWarps = kWrThreads / 32;
constexpr uint32_t kGroupBytes = 32u * 31u;  // 32 Points = 992 bytes = 62 16-byte words: groups of a 16-byte aligned array stay aligned

// doubles in an order-preserving u64 code (atomicMin / atomicMax have no f64 form)
__device__ __forceinline__ unsigned long long enc_f64(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
inline double dec_f64(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  double d;
  std::memcpy(&d, &b, 8);
  return d;
}

// the 32 Points of a warp's group: coalesced 16-byte loads into shared memory, then every lane takes the eight words
// of its Point (31-byte stride: aligned word loads + one funnel shift per word)
__device__ __forceinline__ bool load_point(const uint8_t* pts, uint64_t n, uint64_t group, uint32_t* sm /* 62 * 4 + 2 words */,
                                           uint32_t w[8]) {
  const uint32_t ln = threadIdx.x & 31u;
  const uint64_t first = group * 32ull;
  const uint64_t bytes = (n - first < 32ull ? n - first : 32ull) * 31ull;
  const uint4* src = reinterpret_cast<const uint4*>(pts + first * 31ull);
  uint4* dst = reinterpret_cast<uint4*>(sm);
  for (uint32_t k = ln; k < (bytes + 15u) / 16u; k += 32u) dst[k] = src[k];  // (the array is padded to a multiple of 16 bytes)
  __syncwarp();
  const bool valid = first + ln < n;
  const uint32_t o = ln * 31u, sh = (o & 3u) * 8u;
  const uint32_t* p = sm + (o >> 2);
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = valid ? __funnelshift_r(p[k], p[k + 1], sh) : 0u;
  __syncwarp();
  return valid;
}

__global__ void __launch_bounds__(kWrThreads) k_points_minmax(const uint8_t* pts, uint64_t n, unsigned long long* mm /* min x,y,z, max x,y,z */) {
  __shared__ __align__(16) uint32_t sm[kWrWarps][kGroupBytes / 4 + 4];
  const uint32_t wid = threadIdx.x >> 5;
  const uint64_t groups = (n + 31ull) / 32ull;
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
  for (uint64_t g = (uint64_t)blockIdx.x * kWrWarps + wid; g < groups; g += (uint64_t)gridDim.x * kWrWarps) {
    uint32_t w[8];
    if (load_point(pts, n, g, sm[wid], w)) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double v = __hiloint2double((int)w[2 * a + 1], (int)w[2 * a]);
        if (v == v) {  // (a NaN compares false in the fold of :77-80 and leaves the state as it is)
          const unsigned long long e = enc_f64(v);
          lo[a] = e < lo[a] ? e : lo[a];
          hi[a] = e > hi[a] ? e : hi[a];
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo[a], o), h2 = __shfl_xor_sync(0xffffffffu, hi[a], o);
      lo[a] = l2 < lo[a] ? l2 : lo[a];
      hi[a] = h2 > hi[a] ? h2 : hi[a];
    }
    if ((threadIdx.x & 31u) == 0) {
      atomicMin(mm + a, lo[a]);
      atomicMax(mm + 3 + a, hi[a]);
    }
  }
}

// LAS point format 2 record: x, y, z i32 @0/4/8, intensity u16 @12, return byte @14, classification @15, scan angle
// @16, user data @17, point source id u16 @18, r, g, b u16 @20/22/24
__global__ void __launch_bounds__(kWrThreads) k_points_to_las2(const uint8_t* pts, uint64_t n, double ox, double oy, double oz, double scale,
                                                             uint8_t* out) {
  __shared__ __align__(16) uint32_t sm[kWrWarps][kGroupBytes / 4 + 4];
  __shared__ __align__(16) uint32_t so[kWrWarps][32 * 26 / 4];
  const uint32_t wid = threadIdx.x >> 5, ln = threadIdx.x & 31u;
  const uint64_t groups = (n + 31ull) / 32ull;
  const double off[3] = {ox, oy, oz};
  for (uint64_t g = (uint64_t)blockIdx.x * kWrWarps + wid; g < groups; g += (uint64_t)gridDim.x * kWrWarps) {
    uint32_t w[8];
    const bool valid = load_point(pts, n, g, sm[wid], w);
    uint16_t* rec = reinterpret_cast<uint16_t*>(so[wid]) + ln * 13u;  // 26 bytes, 2-byte aligned
    int32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = __hiloint2double((int)w[2 * a + 1], (int)w[2 * a]);
      // las-rs Transform::inverse: round((p - offset) / scale) as i32, saturating (f64::round: half away from zero)
      const double r = round(__ddiv_rn(__dsub_rn(v, off[a]), scale));
      q[a] = r >= 2147483647.0 ? INT_MAX : (r <= -2147483648.0 ? INT_MIN : (r == r ? (int32_t)r : 0));
    }
    rec[0] = (uint16_t)(uint32_t)q[0];
    rec[1] = (uint16_t)((uint32_t)q[0] >> 16);
    rec[2] = (uint16_t)(uint32_t)q[1];
    rec[3] = (uint16_t)((uint32_t)q[1] >> 16);
    rec[4] = (uint16_t)(uint32_t)q[2];
    rec[5] = (uint16_t)((uint32_t)q[2] >> 16);
    rec[6] = 0;                                               // intensity
    rec[7] = (uint16_t)(0x09u | ((w[7] >> 16 & 0xFFu) << 8));  // return 1 of 1, classification (Point byte 30)
    rec[8] = 0;
    rec[9] = 0;
    rec[10] = (uint16_t)(w[6] & 0xFFFFu);  // r (Point bytes 24-25)
    rec[11] = (uint16_t)(w[6] >> 16);      // g
    rec[12] = (uint16_t)(w[7] & 0xFFFFu);  // b
    __syncwarp();
    // a group's 32 records = 832 bytes = 52 16-byte words, contiguous and 16-byte aligned in the output
    const uint64_t first = g * 32ull;
    const uint32_t bytes = (uint32_t)(n - first < 32ull ? n - first : 32ull) * 26u;
    uint4* dst = reinterpret_cast<uint4*>(out + first * 26ull);
    const uint4* srcw = reinterpret_cast<const uint4*>(so[wid]);
    for (uint32_t k = ln; k < (bytes + 15u) / 16u; k += 32u) dst[k] = srcw[k];  // (the output is padded as well)
    __syncwarp();
    (void)valid;
  }
}

}  // namespace

// :81-88  one scale for all axes: the next power of ten above max_extent / i32::MAX, at least a millimetre
double pcq::las_writer_scale(double max_extent) {
  const double min_scale = max_extent / (double)INT32_MAX;
  double scale = std::pow(10.0, std::ceil(std::log10(min_scale)));
  if (!(scale >= 0.001)) scale = 0.001;  // `if scale < 0.001` plus the NaN / 0 cases of a zero extent
  return scale;
}

extern "C" int pcq_collector_las_records(pcq_collector* c, double out_min[3], double out_max[3], double* out_scale,
                                         const uint8_t** out_records, uint64_t* out_n) {
  if (!c || !out_min || !out_max || !out_scale || !out_records || !out_n) return fail(PCQ_ERR_ARG, "null argument");
  *out_records = nullptr;
  *out_n = 0;
  const void* dptr = nullptr;
  uint64_t n = 0;
  RC(pcq_collector_points_device(c, &dptr, &n));
  if (n == 0) return PCQ_OK;  // :65-67: an empty buffer writes nothing
  pcq_ctx* ctx = c->ctx;
  uint8_t* d_tmp = nullptr;
  const size_t out_bytes = round_up((size_t)n * 26u, 16) + 16;
  if (cudaMalloc(&d_tmp, 64 + out_bytes) != cudaSuccess) {
    cudaGetLastError();
    return fail(PCQ_ERR_NOMEM, "cannot allocate %zu bytes of HBM for the LAS records", out_bytes);
  }
  unsigned long long* d_mm = reinterpret_cast<unsigned long long*>(d_tmp);
  uint8_t* d_rec = d_tmp + 64;
  auto bail = [&](int rc) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    return rc;
  };
  const unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  unsigned long long mm[6];
  const uint64_t groups = (n + 31) / 32;
  const unsigned grid = (unsigned)std::min<uint64_t>((groups + kWrWarps - 1) / kWrWarps, (uint64_t)ctx->sm_count * 8);
  if (cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return bail(fail(PCQ_ERR_CUDA, "upload of the min/max seeds failed"));
  k_points_minmax<<<grid, kWrThreads, 0, ctx->stream>>>(static_cast<const uint8_t*>(dptr), n, d_mm);
  if (cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    return bail(fail(PCQ_ERR_CUDA, "k_points_minmax failed: %s", cudaGetErrorString(cudaGetLastError())));
  double mn[3], mx[3];
  for (int a = 0; a < 3; ++a) {
    // (no finite position on an axis: the fold's seeds f64::MAX / f64::MIN survive, :74-76)
    mn[a] = mm[a] == ~0ull ? DBL_MAX : dec_f64(mm[a]);
    mx[a] = mm[3 + a] == 0ull ? -DBL_MAX : dec_f64(mm[3 + a]);
  }
  const double scale = pcq::las_writer_scale(std::max({mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]}));
  k_points_to_las2<<<grid, kWrThreads, 0, ctx->stream>>>(static_cast<const uint8_t*>(dptr), n, mn[0], mn[1], mn[2], scale, d_rec);
  ctx->launches += 2;
  if (c->h_cap * 31ull < n * 26ull) {  // the collector's pinned host array (capacity counted in Points)
    if (c->h_pts) cudaFreeHost(c->h_pts);
    c->h_pts = nullptr;
    c->h_cap = 0;
    const uint64_t cap = std::max<uint64_t>(n, 4096);
    if (cudaMallocHost(&c->h_pts, cap * 31ull) != cudaSuccess) {
      cudaGetLastError();
      return bail(fail(PCQ_ERR_NOMEM, "cannot pin %llu bytes of host memory", (unsigned long long)(cap * 31ull)));
    }
    c->h_cap = cap;
  }
  if (cudaMemcpyAsync(c->h_pts, d_rec, n * 26ull, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    return bail(fail(PCQ_ERR_CUDA, "k_points_to_las2 failed: %s", cudaGetErrorString(cudaGetLastError())));
  cudaFree(d_tmp);
  for (int a = 0; a < 3; ++a) {
    out_min[a] = mn[a];
    out_max[a] = mx[a];
  }
  *out_scale = scale;
  *out_records = static_cast<const uint8_t*>(c->h_pts);
  *out_n = n;
  return PCQ_OK;
}
