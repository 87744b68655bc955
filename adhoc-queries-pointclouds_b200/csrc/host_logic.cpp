// host_logic.cpp — the part of the path that runs once per file / once per query on the host:
// LAS header parsing, query bounds -> local integer bounds, AABB overlap, SparseGrid geometry.
// No CUDA in this file; these entry points work without a GPU.
//
// Reference: query/src/search/las.rs:33-36, 59-99; query/src/search/last.rs:36-39, 53-109, 220-250;
// query/src/grid_sampling.rs:18-47.  The header byte layout is the ASPRS LAS 1.x public header
// block that `las` 0.7.4 (un-vendored) reads; field order is corroborated in-repo by
// query/src/las.rs:6-40.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "host_logic.hpp"

namespace pcq {

static thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
const char* last_error() { return g_last_error.c_str(); }

// Rust `as` casts: truncate toward zero, saturate, NaN -> 0
int64_t f64_as_i64(double v) {
  if (std::isnan(v)) return 0;
  if (v >= 9223372036854775808.0) return INT64_MAX;
  if (v <= -9223372036854775808.0) return INT64_MIN;
  return static_cast<int64_t>(v);
}
uint64_t f64_as_u64(double v) {
  if (std::isnan(v) || v <= 0.0) return 0;
  if (v >= 18446744073709551616.0) return UINT64_MAX;
  return static_cast<uint64_t>(v);
}

template <typename T>
static T le(const uint8_t* p) {
  T v;
  std::memcpy(&v, p, sizeof(T));  // x86-64 / aarch64 hosts are little-endian, as is LAS
  return v;
}

static const uint16_t kFormatLen[11] = {20, 28, 26, 34, 57, 63, 30, 36, 38, 59, 67};

uint16_t format_record_len(uint8_t format) { return format <= 10 ? kFormatLen[format] : 0; }

int parse_header(const void* bytes, size_t n, int layout, int mask_format, pcq_file_desc* out, uint8_t* raw_format) {
  if (!bytes || !out) return fail(PCQ_ERR_ARG, "pcq_parse_header: null argument");
  if (layout != PCQ_LAYOUT_LAS && layout != PCQ_LAYOUT_LAST) return fail(PCQ_ERR_ARG, "pcq_parse_header: bad layout %d", layout);
  const uint8_t* b = static_cast<const uint8_t*>(bytes);
  if (n < 227) return fail(PCQ_ERR_IO, "LAS header needs 227 bytes, buffer has %zu", n);
  if (std::memcmp(b, "LASF", 4) != 0) return fail(PCQ_ERR_FORMAT, "invalid LAS file signature");
  const uint8_t vmaj = b[24], vmin = b[25];
  const uint16_t header_size = le<uint16_t>(b + 94);
  const bool v13 = vmaj > 1 || (vmaj == 1 && vmin >= 3);
  const bool v14 = vmaj > 1 || (vmaj == 1 && vmin >= 4);
  size_t need = 227 + (v13 ? 8 : 0) + (v14 ? 140 : 0);
  if (n < need) return fail(PCQ_ERR_IO, "LAS %u.%u header needs %zu bytes, buffer has %zu", vmaj, vmin, need, n);
  if (header_size > need && n < header_size) return fail(PCQ_ERR_IO, "LAS header_size %u exceeds buffer", header_size);

  std::memset(out, 0, sizeof(*out));
  out->layout = static_cast<uint8_t>(layout);
  out->point_data_off = le<uint32_t>(b + 96);
  uint8_t fmt = b[104];
  if (raw_format) *raw_format = fmt;
  if (mask_format) fmt &= 0x0F;  // last.rs:222, last_reader.rs:76-79
  out->format = fmt;
  out->record_len = le<uint16_t>(b + 105);
  const uint32_t legacy = le<uint32_t>(b + 107);
  for (int i = 0; i < 3; ++i) out->scale[i] = le<double>(b + 131 + 8 * i);
  for (int i = 0; i < 3; ++i) out->offset[i] = le<double>(b + 155 + 8 * i);
  for (int i = 0; i < 3; ++i) {
    out->hdr_max[i] = le<double>(b + 179 + 16 * i);
    out->hdr_min[i] = le<double>(b + 187 + 16 * i);
  }
  // Header::from_raw
  if (fmt > 10) return fail(PCQ_ERR_FORMAT, "Invalid LAS format %u", fmt);
  if (out->record_len < kFormatLen[fmt])
    return fail(PCQ_ERR_FORMAT, "point data record length %u too small for format %u", out->record_len, fmt);
  if (fmt >= 6 && !v14) return fail(PCQ_ERR_FORMAT, "point format %u needs LAS 1.4, file is %u.%u", fmt, vmaj, vmin);
  out->n_points = legacy > 0 ? static_cast<uint64_t>(legacy) : (v14 ? le<uint64_t>(b + 247) : 0);
  return PCQ_OK;
}

int local_bounds(const pcq_file_desc* d, const double qmin[3], const double qmax[3], int64_t lo[3], int64_t hi[3]) {
  // las.rs:88-99 / last.rs:98-109.  min.y and min.z are divided by the X scale factor, exactly as
  // the reference does.
  lo[0] = f64_as_i64((qmin[0] - d->offset[0]) / d->scale[0]);
  lo[1] = f64_as_i64((qmin[1] - d->offset[1]) / d->scale[0]);
  lo[2] = f64_as_i64((qmin[2] - d->offset[2]) / d->scale[0]);
  hi[0] = f64_as_i64((qmax[0] - d->offset[0]) / d->scale[0]);
  hi[1] = f64_as_i64((qmax[1] - d->offset[1]) / d->scale[1]);
  hi[2] = f64_as_i64((qmax[2] - d->offset[2]) / d->scale[2]);
  for (int i = 0; i < 3; ++i)
    if (lo[i] > hi[i])
      return fail(PCQ_ERR_PANIC, "AABB::from_min_max: local query bounds have min > max on axis %d (%lld > %lld)", i,
                  (long long)lo[i], (long long)hi[i]);
  return PCQ_OK;
}

int file_intersects(const pcq_file_desc* d, const double qmin[3], const double qmax[3], int* out) {
  for (int i = 0; i < 3; ++i)
    if (d->hdr_min[i] > d->hdr_max[i])
      return fail(PCQ_ERR_PANIC, "AABB::from_min_max: header bounds have min > max on axis %d", i);
  int hit = 1;
  for (int i = 0; i < 3; ++i)
    if (!(d->hdr_min[i] <= qmax[i] && d->hdr_max[i] >= qmin[i])) hit = 0;
  *out = hit;
  return PCQ_OK;
}

int grid_params(const double gmin[3], const double gmax[3], double cell, uint64_t dims[3], uint64_t bits[3]) {
  uint64_t sum = 0;
  for (int i = 0; i < 3; ++i) {
    const double extent = gmax[i] - gmin[i];          // grid_sampling.rs:19-23
    const double ncells = std::ceil(extent / cell);   // :24-28
    bits[i] = f64_as_u64(std::ceil(std::log2(ncells)));  // :29-31
    dims[i] = f64_as_u64(ncells);                     // :39-43
    sum += bits[i];                                   // wrapping add, like a release build
  }
  if (sum > 64)  // :32-34
    return fail(PCQ_ERR_GRID, "Too many cells (%llu*%llu*%llu) in SparseGrid! The number of cells exceeds the capacity of a u64 index!",
                (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2]);
  return PCQ_OK;
}

}  // namespace pcq

extern "C" {

const char* pcq_last_error(void) { return pcq::last_error(); }
const char* pcq_version(void) { return "pcq-b200 0.1 (sm_100a)"; }

int pcq_parse_header(const void* bytes, size_t n_bytes, int layout, int mask_format, pcq_file_desc* out) {
  return pcq::parse_header(bytes, n_bytes, layout, mask_format, out, nullptr);
}

int pcq_local_bounds(const pcq_file_desc* desc, const double qmin[3], const double qmax[3], int64_t lo[3], int64_t hi[3]) {
  if (!desc || !qmin || !qmax || !lo || !hi) return pcq::fail(PCQ_ERR_ARG, "pcq_local_bounds: null argument");
  return pcq::local_bounds(desc, qmin, qmax, lo, hi);
}

int pcq_file_intersects(const pcq_file_desc* desc, const double qmin[3], const double qmax[3], int* out) {
  if (!desc || !qmin || !qmax || !out) return pcq::fail(PCQ_ERR_ARG, "pcq_file_intersects: null argument");
  return pcq::file_intersects(desc, qmin, qmax, out);
}

int pcq_grid_params(const double gmin[3], const double gmax[3], double cell_size, uint64_t dims[3], uint64_t bits[3]) {
  if (!gmin || !gmax || !dims || !bits) return pcq::fail(PCQ_ERR_ARG, "pcq_grid_params: null argument");
  return pcq::grid_params(gmin, gmax, cell_size, dims, bits);
}

}  // extern "C"
