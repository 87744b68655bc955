/*
 * pcq_synth.h — seeded synthetic LAS / LAST datasets of the navvis / doc / ca13 shapes
 * (SURVEY.md §8d).  Test and benchmark tooling, not part of the drop-in boundary: the reference's
 * datasets are private files (readers/src/last_reader.rs:406-407), so parity and throughput runs
 * use counter-based, integer-only generators that produce identical bytes on the host (small
 * files for the oracle) and on the device (billions of points for the benchmark).
 */
#ifndef PCQ_SYNTH_H
#define PCQ_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#include "pcq.h"

#ifdef __cplusplus
extern "C" {
#endif

enum pcq_synth_shape {
  PCQ_SHAPE_UNIFORM = 0, /* x, y, z uniform in [lo, hi]                                   */
  PCQ_SHAPE_TERRAIN = 1, /* x, y uniform; z bell-shaped over [lo.z, hi.z] (doc-like tiles)   */
  PCQ_SHAPE_INDOOR = 2,  /* points on floors / walls with +-2 raw units of noise (navvis)   */
  PCQ_SHAPE_RELIEF = 3   /* x, y uniform; z follows an integer relief plus noise (ca13)     */
};

typedef struct pcq_synth_spec {
  uint64_t seed;
  uint64_t n_points;
  uint8_t layout;     /* pcq_layout */
  uint8_t format;     /* 0..3 */
  uint8_t shape;      /* pcq_synth_shape */
  uint8_t n_classes;  /* 1..8 */
  uint16_t record_len; /* >= size of `format`; extra bytes are filled from the hash */
  uint16_t flag_per_64k; /* of 65536 points, how many get one of the flag bits 0x20/0x40/0x80 ORed into byte 15 */
  int32_t lo[3];      /* raw coordinate range, inclusive */
  int32_t hi[3];
  double scale[3];
  double offset[3];
  uint8_t class_val[8];
  uint16_t class_cum[8]; /* cumulative thresholds out of 65536; class k if r16 < class_cum[k] (last must be 65535) */
} pcq_synth_spec;

/* libpcq_synth.so is self-contained (no dependency on libpcq.so); errors: PCQ_OK or a negative pcq_status,
 * message through pcq_synth_last_error(). */
const char* pcq_synth_last_error(void);

/* size of the whole file image (227-byte LAS 1.2 header + point data) */
size_t pcq_synth_file_size(const pcq_synth_spec* spec);

/* Writes the whole file image (header with true min/max + points) into host memory. */
int pcq_synth_host(const pcq_synth_spec* spec, void* out, size_t cap);

/* Point data of the range [first_point, first_point + n_points) of the file (no header) into host memory, as a block
 * of its own (LAS: n_points records; LAST: columns of n_points entries each); raw coordinate minima / maxima of the
 * range in minmax[0..2] / minmax[3..5].  Thread-safe: callers split a file over threads. */
int pcq_synth_host_points(const pcq_synth_spec* spec, uint64_t first_point, uint64_t n_points, void* out_points,
                          int32_t minmax[6]);

/* The same bytes written into device memory by a kernel on `device` (synchronous). */
int pcq_synth_device_points(int device, const pcq_synth_spec* spec, uint64_t first_point, uint64_t n_points,
                            void* dev_point_data, int32_t minmax[6]);
/* whole file: first_point = 0, n_points = spec->n_points */
int pcq_synth_device(int device, const pcq_synth_spec* spec, void* dev_point_data, int32_t minmax[6]);

/* 227-byte LAS 1.2 header for `spec` whose bounds are minmax * scale + offset. */
int pcq_synth_header(const pcq_synth_spec* spec, const int32_t minmax[6], void* out227);

/* file descriptor equal to what pcq_parse_header would return for that header */
int pcq_synth_desc(const pcq_synth_spec* spec, const int32_t minmax[6], pcq_file_desc* out);

#ifdef __cplusplus
}
#endif
#endif
