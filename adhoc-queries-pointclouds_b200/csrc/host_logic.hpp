// host_logic.hpp — host-only pieces of the scan path shared by the C ABI and the C++ host mirror.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/pcq.h"

namespace pcq {

// records a thread-local message (pcq_last_error) and returns `code`
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
const char* last_error();

int64_t f64_as_i64(double v);
uint64_t f64_as_u64(double v);
uint16_t format_record_len(uint8_t format);

// raw_format (optional) receives point_data_record_format before masking
int parse_header(const void* bytes, size_t n, int layout, int mask_format, pcq_file_desc* out, uint8_t* raw_format);
int local_bounds(const pcq_file_desc* d, const double qmin[3], const double qmax[3], int64_t lo[3], int64_t hi[3]);
int file_intersects(const pcq_file_desc* d, const double qmin[3], const double qmax[3], int* out);
int grid_params(const double gmin[3], const double gmax[3], double cell, uint64_t dims[3], uint64_t bits[3]);

}  // namespace pcq
