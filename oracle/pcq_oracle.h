/*
 * pcq_oracle.h — CPU oracle: a plain-C restatement of the reference's `--optimized` scan path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (include/, adhoc-queries-pointclouds_b200/)
 * may include, link, call or execute this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity status: the reference is Rust and cannot be built in this environment (no
 * cargo/rustc, no crate sources), so there is no oracle/_ref.  The only reference tests on this path
 * are the three SparseGrid tests (query/src/grid_sampling.rs:121-208); the oracle is pinned against
 * those (tests/test_oracle.py).  Nothing in the reference pins bbox / class scan results and the
 * header / AABB / distance arithmetic lives in un-vendored crates (las 0.7.4, pasture-core 0.1.0,
 * nalgebra 0.23.2): for the scans themselves this oracle is "PARITY UNPINNED" — a line-by-line
 * restatement anchored on the reference's call sites, cross-checked by an independent numpy
 * restatement (oracle/np_oracle.py) and the survey's known-answer vectors.
 */
#ifndef PCQ_ORACLE_H
#define PCQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_ERR_IO -1      /* read past end of file (UnexpectedEof) */
#define ORC_ERR_FORMAT -2  /* anyhow error: bad signature / format */
#define ORC_ERR_PANIC -3   /* the reference panics here */
#define ORC_ERR_GRID -6    /* SparseGrid::new error */

/* readers::Point, readers/src/lib.rs:10-19 */
typedef struct __attribute__((packed)) orc_point {
  double pos[3];
  uint16_t rgb[3];
  uint8_t cls;
} orc_point;

/* las::raw::Header fields the path touches (+ what Header::from_raw derives) */
typedef struct orc_header {
  uint8_t version_major, version_minor;
  uint16_t header_size;
  uint32_t offset_to_point_data;
  uint32_t n_vlrs;
  uint8_t format;          /* raw point_data_record_format */
  uint16_t record_len;
  uint32_t legacy_count;
  double scale[3], offset[3];
  double max[3], min[3];
  uint64_t large_count;
  int has_large;
  uint64_t n_points;       /* header.number_of_points() */
} orc_header;

/* SparseGrid, query/src/grid_sampling.rs:9-15 */
typedef struct orc_grid {
  double bmin[3], bmax[3];
  double cell_size;
  uint64_t dims[3];
  uint64_t bits[3];
  /* HashMap<u64, Point> as open addressing */
  uint64_t* keys;
  orc_point* vals;
  uint8_t* used;
  size_t cap, len;
} orc_grid;

enum { ORC_COLLECT_COUNT = 0, ORC_COLLECT_BUFFER = 1, ORC_COLLECT_GRID = 2 };

/* ResultCollector implementations, query/src/collect_points.rs */
typedef struct orc_collector {
  int kind;
  size_t count;
  orc_point* buf;
  size_t buf_len, buf_cap;
  orc_grid* grid;
} orc_collector;

/* Rust `as` casts */
int64_t orc_f64_as_i64(double v);
uint64_t orc_f64_as_u64(double v);

int orc_parse_header(const uint8_t* bytes, size_t n, int mask_format, orc_header* out);
int orc_local_bounds(const orc_header* h, const double qmin[3], const double qmax[3], int64_t lo[3],
                     int64_t hi[3]);

int orc_grid_new(const double bmin[3], const double bmax[3], double cell_size, orc_grid** out);
void orc_grid_free(orc_grid* g);
int orc_grid_insert_point(orc_grid* g, const orc_point* p); /* returns 1 if stored/replaced, 0 if not */
size_t orc_grid_len(const orc_grid* g);
size_t orc_grid_cells(const orc_grid* g, uint64_t* keys_out, size_t cap);
size_t orc_grid_points(const orc_grid* g, orc_point* out, size_t cap);
/* the key insert_point would use, and whether any axis cell exceeds its mask (aliasing) */
uint64_t orc_grid_key(const orc_grid* g, const double pos[3], int* aliased);

int orc_collector_new(int kind, const double gmin[3], const double gmax[3], double cell, orc_collector** out);
void orc_collector_free(orc_collector* c);
void orc_collect_one(orc_collector* c, const orc_point* p);
size_t orc_collector_point_count(const orc_collector* c);
/* copies points() into out (BUFFER: scan order; GRID: arbitrary order); returns number copied */
size_t orc_collector_points(const orc_collector* c, orc_point* out, size_t cap);

int orc_search_las_file_by_bounds_optimized(const uint8_t* file, size_t n, const double qmin[3],
                                            const double qmax[3], orc_collector* c);
int orc_search_las_file_by_classification_optimized(const uint8_t* file, size_t n, uint8_t cls,
                                                    orc_collector* c);
int orc_search_last_file_by_bounds_optimized(const uint8_t* file, size_t n, const double qmin[3],
                                             const double qmax[3], orc_collector* c);
int orc_search_last_file_by_classification_optimized(const uint8_t* file, size_t n, uint8_t cls,
                                                     orc_collector* c);

/* Searcher::search_file dispatch on extension, Optimized arm (searcher.rs:43-91, 104-152) */
int orc_search_file(const uint8_t* file, size_t n, const char* ext, int query_kind /*0 bounds,1 class*/,
                    const double qmin[3], const double qmax[3], uint8_t cls, orc_collector* c);

/* run_search_parallel with CountCollectors (main.rs:146-183): one task per file on n_threads
 * workers (rayon par_iter ~ min(files, cores)); per-file counts out.  Returns first error. */
int orc_count_parallel(const uint8_t* const* files, const size_t* sizes, const char* const* exts,
                       size_t n_files, int query_kind, const double qmin[3], const double qmax[3],
                       uint8_t cls, int n_threads, uint64_t* per_file_counts);

/* The chunk headers improvements.md:3-10 proposes (the reference does not implement them): per chunk of
 * `chunk_points` consecutive points of the range [first, first + count) of a las / last file image, the min / max
 * of the raw x, y, z fields and the set of the class bytes the class search compares (las.rs:202-212,
 * last.rs:245-259).  Checker of pcq_file_build_index.  out: ceil(count / chunk_points) headers of 16 u32 each
 * (lo[3], hi[3] as i32, cls_bits[8], n_points, 0). */
int orc_chunk_headers(const uint8_t* file, size_t n, const char* ext, uint64_t first, uint64_t count,
                      uint32_t chunk_points, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
