// synth.cu — seeded synthetic LAS / LAST data (include/pcq_synth.h).  Test / benchmark tooling.
//
// Every value of point i is a pure integer function of (seed, i), so the host loop and the device
// kernel write identical bytes.  Layouts: LAS = row-major records; LAST = the same record
// transposed field by field (column of the field at record offset k starts at k * N,
// readers/src/last_reader.rs:88-144; query/src/search/last.rs:80-90, 114, 245-250).
#include <cuda_runtime.h>

#include <climits>
#include <cstring>

#include <cstdarg>
#include <cstdio>

#include "../../include/pcq_synth.h"

// libpcq_synth.so is self-contained (it does not link libpcq.so): the CPU reference arm of bench.py generates its
// inputs with it without mapping the product library.
namespace pcq {
namespace {
thread_local char g_synth_err[512] = "";
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(g_synth_err, sizeof(g_synth_err), fmt, ap);
  va_end(ap);
  return code;
}
uint16_t format_record_len(uint8_t fmt) {  // ASPRS LAS 1.x point data record formats 0..3
  static const uint16_t len[4] = {20, 28, 26, 34};
  return fmt < 4 ? len[fmt] : 0xFFFF;
}
}  // namespace
}  // namespace pcq

#define HD __host__ __device__ __forceinline__

namespace {

constexpr int kMaxRecord = 96;

HD uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
HD uint64_t draw(uint64_t seed, uint64_t i, uint64_t k) {
  return splitmix(splitmix(seed ^ (k * 0xD6E8FEB86659FD93ULL)) + i * 0x9E3779B97F4A7C15ULL);
}
HD int32_t uniform_in(uint64_t r, int32_t lo, int32_t hi) {
  const uint64_t span = (uint64_t)((int64_t)hi - (int64_t)lo) + 1ull;
  return (int32_t)((int64_t)lo + (int64_t)(((r >> 32) * span) >> 32));
}
HD int32_t clampi(int64_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : (int32_t)v); }
HD uint64_t tri(uint64_t t, uint64_t period) {  // triangle wave in [0, period/2]
  t %= period;
  return t < period / 2 ? t : period - t;
}

struct Fields {
  int32_t x, y, z;
  uint8_t cls;
};

// fills rec[0 .. record_len) with the LAS record of point i
HD void make_record(const pcq_synth_spec& sp, uint64_t i, uint8_t* rec, Fields* f) {
  const uint64_t r0 = draw(sp.seed, i, 0), r1 = draw(sp.seed, i, 1), r2 = draw(sp.seed, i, 2);
  const uint64_t r3 = draw(sp.seed, i, 3), r4 = draw(sp.seed, i, 4), r5 = draw(sp.seed, i, 5);
  int32_t x = uniform_in(r0, sp.lo[0], sp.hi[0]);
  int32_t y = uniform_in(r1, sp.lo[1], sp.hi[1]);
  int32_t z = uniform_in(r2, sp.lo[2], sp.hi[2]);
  const uint64_t span_x = (uint64_t)((int64_t)sp.hi[0] - sp.lo[0]) + 1, span_y = (uint64_t)((int64_t)sp.hi[1] - sp.lo[1]) + 1,
                 span_z = (uint64_t)((int64_t)sp.hi[2] - sp.lo[2]) + 1;
  if (sp.shape == PCQ_SHAPE_TERRAIN) {
    const uint64_t s = (r2 & 0xFFFF) + ((r2 >> 16) & 0xFFFF) + ((r2 >> 32) & 0xFFFF) + ((r2 >> 48) & 0xFFFF);
    z = clampi((int64_t)sp.lo[2] + (int64_t)((s * span_z) / 262141ull), sp.lo[2], sp.hi[2]);
  } else if (sp.shape == PCQ_SHAPE_INDOOR) {
    const uint32_t choose = (uint32_t)(r3 & 0xFF);
    const int64_t noise = (int64_t)((r3 >> 8) % 5) - 2;
    if (choose < 102) {  // floors / ceilings
      const uint64_t k = (r3 >> 16) % 4;
      z = clampi((int64_t)sp.lo[2] + (int64_t)((span_z - 1) * k / 3) + noise, sp.lo[2], sp.hi[2]);
    } else if (choose < 179) {  // walls across x
      const uint64_t k = (r3 >> 16) % 8;
      x = clampi((int64_t)sp.lo[0] + (int64_t)((span_x - 1) * k / 7) + noise, sp.lo[0], sp.hi[0]);
    } else {  // walls across y
      const uint64_t k = (r3 >> 16) % 8;
      y = clampi((int64_t)sp.lo[1] + (int64_t)((span_y - 1) * k / 7) + noise, sp.lo[1], sp.hi[1]);
    }
  } else if (sp.shape == PCQ_SHAPE_RELIEF) {
    const uint64_t px = span_x / 4 > 2 ? span_x / 4 : 2, py = span_y / 3 > 2 ? span_y / 3 : 2;
    const uint64_t tx = tri((uint64_t)((int64_t)x - sp.lo[0]), px), ty = tri((uint64_t)((int64_t)y - sp.lo[1]), py);
    const uint64_t relief = (tx * (span_z / 3)) / px + (ty * (span_z / 3)) / py;
    const uint64_t noise = (r3 >> 20) % (span_z / 16 + 1);
    z = clampi((int64_t)sp.lo[2] + (int64_t)relief + (int64_t)noise, sp.lo[2], sp.hi[2]);
  }
  // classification byte: categorical class, a few with flag bits set (only non-extended formats)
  const uint32_t c16 = (uint32_t)(r4 & 0xFFFF);
  uint8_t cls = sp.class_val[sp.n_classes ? sp.n_classes - 1 : 0];
  for (int k = 0; k < 8; ++k) {
    if (k < sp.n_classes && c16 <= sp.class_cum[k]) {
      cls = sp.class_val[k];
      break;
    }
  }
  if ((uint32_t)((r4 >> 16) & 0xFFFF) < sp.flag_per_64k) cls |= (uint8_t)(0x20u << (uint32_t)((r4 >> 32) % 3));

  for (int b = 0; b < kMaxRecord; ++b) rec[b] = 0;
  const uint32_t ux = (uint32_t)x, uy = (uint32_t)y, uz = (uint32_t)z;
  for (int b = 0; b < 4; ++b) {
    rec[b] = (uint8_t)(ux >> (8 * b));
    rec[4 + b] = (uint8_t)(uy >> (8 * b));
    rec[8 + b] = (uint8_t)(uz >> (8 * b));
  }
  rec[12] = (uint8_t)(r4 >> 48);  // intensity
  rec[13] = (uint8_t)(r4 >> 56);
  rec[14] = (uint8_t)(0x09u | ((uint32_t)(r5 & 1) << 6));  // return 1 of 1, scan direction flag
  rec[15] = cls;
  rec[16] = (uint8_t)(r5 >> 8);   // scan angle rank
  rec[17] = (uint8_t)(r5 >> 16);  // user data
  rec[18] = (uint8_t)(r5 >> 24);  // point source id
  rec[19] = (uint8_t)(r5 >> 32);
  int fmt_len = 20;
  if (sp.format == 1 || sp.format == 3) {
    const double gps = (double)(1000000ull + i) * 0.0001;  // one rounding: identical on host and device
    unsigned long long gb;
#ifdef __CUDA_ARCH__
    gb = (unsigned long long)__double_as_longlong(gps);
#else
    std::memcpy(&gb, &gps, 8);
#endif
    for (int b = 0; b < 8; ++b) rec[20 + b] = (uint8_t)(gb >> (8 * b));
    fmt_len = 28;
  }
  if (sp.format == 2 || sp.format == 3) {
    const uint64_t r6 = draw(sp.seed, i, 6);
    const int o = sp.format == 2 ? 20 : 28;
    for (int b = 0; b < 6; ++b) rec[o + b] = (uint8_t)(r6 >> (8 * b));
    fmt_len = o + 6;
  }
  if (sp.record_len > fmt_len) {
    for (int b = fmt_len; b < sp.record_len && b < kMaxRecord; ++b) rec[b] = (uint8_t)(draw(sp.seed, i, 7 + (uint64_t)(b >> 3)) >> (8 * (b & 7)));
  }
  f->x = x;
  f->y = y;
  f->z = z;
  f->cls = cls;
}

struct FieldDef {
  int off, size;
};

// fields of a record in file order (each becomes one LAST column)
HD int record_fields(uint8_t format, int record_len, FieldDef* out) {
  int n = 0;
  out[n++] = {0, 12};   // position
  out[n++] = {12, 2};   // intensity
  out[n++] = {14, 1};   // return bits
  out[n++] = {15, 1};   // classification
  out[n++] = {16, 1};   // scan angle rank
  out[n++] = {17, 1};   // user data
  out[n++] = {18, 2};   // point source id
  int len = 20;
  if (format == 1 || format == 3) {
    out[n++] = {20, 8};  // gps time
    len = 28;
  }
  if (format == 2 || format == 3) {
    out[n++] = {len, 6};  // colour
    len += 6;
  }
  if (record_len > len) out[n++] = {len, record_len - len};  // extra bytes
  return n;
}

// `i` = index of the record inside the block at dst, `n_col` = points per LAST column of that block
HD void store_record(const pcq_synth_spec& sp, uint64_t i, uint64_t n_col, const uint8_t* rec, uint8_t* dst) {
  if (sp.layout == PCQ_LAYOUT_LAS) {
    uint8_t* p = dst + i * (uint64_t)sp.record_len;
    for (int b = 0; b < sp.record_len; ++b) p[b] = rec[b];
    return;
  }
  FieldDef fd[12];
  const int nf = record_fields(sp.format, sp.record_len, fd);
  for (int k = 0; k < nf; ++k) {
    uint8_t* p = dst + (uint64_t)fd[k].off * n_col + i * (uint64_t)fd[k].size;
    for (int b = 0; b < fd[k].size; ++b) p[b] = rec[fd[k].off + b];
  }
}

__global__ void k_synth(pcq_synth_spec sp, uint64_t first, uint64_t n, uint8_t* dst, int* minmax) {
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint8_t rec[kMaxRecord];
    Fields f;
    make_record(sp, first + i, rec, &f);
    store_record(sp, i, n, rec, dst);
    mn[0] = min(mn[0], f.x);
    mn[1] = min(mn[1], f.y);
    mn[2] = min(mn[2], f.z);
    mx[0] = max(mx[0], f.x);
    mx[1] = max(mx[1], f.y);
    mx[2] = max(mx[2], f.z);
  }
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(minmax + a, mn[a]);
      atomicMax(minmax + 3 + a, mx[a]);
    }
  }
}

int check_spec(const pcq_synth_spec* sp) {
  if (!sp) return pcq::fail(PCQ_ERR_ARG, "null spec");
  if (sp->format > 3) return pcq::fail(PCQ_ERR_ARG, "synthetic data covers point formats 0..3");
  if (sp->layout > 1) return pcq::fail(PCQ_ERR_ARG, "bad layout");
  if (sp->record_len < pcq::format_record_len(sp->format) || sp->record_len > kMaxRecord)
    return pcq::fail(PCQ_ERR_ARG, "record_len %u out of range for format %u", sp->record_len, sp->format);
  if (sp->n_classes < 1 || sp->n_classes > 8) return pcq::fail(PCQ_ERR_ARG, "n_classes must be 1..8");
  for (int a = 0; a < 3; ++a)
    if (sp->lo[a] > sp->hi[a]) return pcq::fail(PCQ_ERR_ARG, "lo > hi on axis %d", a);
  return PCQ_OK;
}

template <typename T>
void put(uint8_t* p, T v) {
  std::memcpy(p, &v, sizeof(T));
}

}  // namespace

extern "C" {

size_t pcq_synth_file_size(const pcq_synth_spec* sp) { return sp ? 227 + (size_t)sp->n_points * sp->record_len : 0; }

int pcq_synth_header(const pcq_synth_spec* sp, const int32_t minmax[6], void* out227) {
  int rc = check_spec(sp);
  if (rc != PCQ_OK) return rc;
  if (!minmax || !out227) return pcq::fail(PCQ_ERR_ARG, "null argument");
  if (sp->n_points > 0xFFFFFFFFull) return pcq::fail(PCQ_ERR_ARG, "a LAS 1.2 header holds at most 2^32-1 points");
  uint8_t* h = static_cast<uint8_t*>(out227);
  std::memset(h, 0, 227);
  std::memcpy(h, "LASF", 4);
  h[24] = 1;
  h[25] = 2;
  std::strncpy(reinterpret_cast<char*>(h + 26), "pcq-b200 synthetic", 32);
  std::strncpy(reinterpret_cast<char*>(h + 58), "pcq_synth", 32);
  put<uint16_t>(h + 90, 291);
  put<uint16_t>(h + 92, 2026);
  put<uint16_t>(h + 94, 227);
  put<uint32_t>(h + 96, 227);
  put<uint32_t>(h + 100, 0);
  h[104] = sp->format;
  put<uint16_t>(h + 105, sp->record_len);
  put<uint32_t>(h + 107, (uint32_t)sp->n_points);
  put<uint32_t>(h + 111, (uint32_t)sp->n_points);
  for (int a = 0; a < 3; ++a) {
    put<double>(h + 131 + 8 * a, sp->scale[a]);
    put<double>(h + 155 + 8 * a, sp->offset[a]);
    const double mx = (double)minmax[3 + a] * sp->scale[a];
    const double mn = (double)minmax[a] * sp->scale[a];
    put<double>(h + 179 + 16 * a, mx + sp->offset[a]);
    put<double>(h + 187 + 16 * a, mn + sp->offset[a]);
  }
  return PCQ_OK;
}

int pcq_synth_desc(const pcq_synth_spec* sp, const int32_t minmax[6], pcq_file_desc* out) {
  int rc = check_spec(sp);
  if (rc != PCQ_OK) return rc;
  if (!minmax || !out) return pcq::fail(PCQ_ERR_ARG, "null argument");
  std::memset(out, 0, sizeof(*out));
  out->layout = sp->layout;
  out->format = sp->format;
  out->record_len = sp->record_len;
  out->point_data_off = 227;
  out->n_points = sp->n_points;
  for (int a = 0; a < 3; ++a) {
    out->scale[a] = sp->scale[a];
    out->offset[a] = sp->offset[a];
    const double mx = (double)minmax[3 + a] * sp->scale[a];
    const double mn = (double)minmax[a] * sp->scale[a];
    out->hdr_max[a] = mx + sp->offset[a];
    out->hdr_min[a] = mn + sp->offset[a];
  }
  return PCQ_OK;
}

int pcq_synth_host_points(const pcq_synth_spec* sp, uint64_t first_point, uint64_t n_points, void* out_points,
                          int32_t minmax[6]) {
  int rc = check_spec(sp);
  if (rc != PCQ_OK) return rc;
  if (!minmax || (!out_points && n_points)) return pcq::fail(PCQ_ERR_ARG, "null argument");
  if (first_point > sp->n_points || n_points > sp->n_points - first_point) return pcq::fail(PCQ_ERR_ARG, "point range outside the file");
  uint8_t* base = static_cast<uint8_t*>(out_points);
  int32_t mm[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  for (uint64_t i = 0; i < n_points; ++i) {
    uint8_t rec[kMaxRecord];
    Fields f;
    make_record(*sp, first_point + i, rec, &f);
    store_record(*sp, i, n_points, rec, base);
    const int32_t v[3] = {f.x, f.y, f.z};
    for (int a = 0; a < 3; ++a) {
      if (v[a] < mm[a]) mm[a] = v[a];
      if (v[a] > mm[3 + a]) mm[3 + a] = v[a];
    }
  }
  for (int a = 0; a < 6; ++a) minmax[a] = mm[a];
  return PCQ_OK;
}

int pcq_synth_host(const pcq_synth_spec* sp, void* out, size_t cap) {
  int rc = check_spec(sp);
  if (rc != PCQ_OK) return rc;
  if (!out || cap < pcq_synth_file_size(sp)) return pcq::fail(PCQ_ERR_ARG, "output buffer too small");
  int32_t mm[6];
  rc = pcq_synth_host_points(sp, 0, sp->n_points, static_cast<uint8_t*>(out) + 227, mm);
  if (rc != PCQ_OK) return rc;
  if (sp->n_points == 0)
    for (int a = 0; a < 6; ++a) mm[a] = 0;
  return pcq_synth_header(sp, mm, out);
}

int pcq_synth_device_points(int device, const pcq_synth_spec* sp, uint64_t first_point, uint64_t n_points,
                            void* dev_point_data, int32_t minmax[6]) {
  int rc = check_spec(sp);
  if (rc != PCQ_OK) return rc;
  if (!minmax || (!dev_point_data && n_points)) return pcq::fail(PCQ_ERR_ARG, "null argument");
  if (first_point > sp->n_points || n_points > sp->n_points - first_point) return pcq::fail(PCQ_ERR_ARG, "point range outside the file");
  if (cudaSetDevice(device) != cudaSuccess) return pcq::fail(PCQ_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(cudaGetLastError()));
  int* d_mm = nullptr;
  if (cudaMalloc(&d_mm, 6 * sizeof(int)) != cudaSuccess) return pcq::fail(PCQ_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
  int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  cudaMemcpy(d_mm, init, sizeof(init), cudaMemcpyHostToDevice);
  if (n_points) {
    uint64_t blocks = (n_points + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    k_synth<<<(unsigned)blocks, 256>>>(*sp, first_point, n_points, static_cast<uint8_t*>(dev_point_data), d_mm);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(minmax, d_mm, sizeof(init), cudaMemcpyDeviceToHost);
  cudaFree(d_mm);
  if (e != cudaSuccess) return pcq::fail(PCQ_ERR_CUDA, "synthetic data kernel failed: %s", cudaGetErrorString(e));
  return PCQ_OK;
}

int pcq_synth_device(int device, const pcq_synth_spec* sp, void* dev_point_data, int32_t minmax[6]) {
  if (!sp) return pcq::fail(PCQ_ERR_ARG, "null spec");
  const int rc = pcq_synth_device_points(device, sp, 0, sp->n_points, dev_point_data, minmax);
  if (rc == PCQ_OK && sp->n_points == 0)
    for (int a = 0; a < 6; ++a) minmax[a] = 0;
  return rc;
}

const char* pcq_synth_last_error(void) { return pcq::g_synth_err; }

}  // extern "C"
