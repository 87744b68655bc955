/*
 * pcq.h — C ABI of the B200-native full-scan point-cloud query path.
 *
 * This is the drop-in boundary for ONE path of igd-geo/adhoc-queries-pointclouds: the `--optimized`
 * full scan of uncompressed LAS (row-major records) and LAST (columnar) files that evaluates
 * bounding-box / class / max-density predicates on every point.  The reference has no FFI; the seam
 * this library sits behind is the Rust trait pair
 *
 *     Searcher::search_file(&self, path, &SearchImplementation, &mut dyn ResultCollector)
 *                                                           query/src/search/searcher.rs:24-31
 *     ResultCollector::{collect_one, points, points_ref, point_count}
 *                                                           query/src/collect_points.rs:7-12
 *
 * and the four free functions it dispatches to for SearchImplementation::Optimized
 * (query/src/search/las.rs:52, 192; query/src/search/last.rs:46, 213).  A per-point `collect_one`
 * callback cannot cross a device boundary, so the three collector kinds of collect_points.rs are
 * exported as device-resident objects that are filled in bulk; a Rust `impl Searcher` shim binds
 * these entry points 1:1.
 *
 * Conventions: plain pointers and sizes only; every function returns PCQ_OK (0) or a negative
 * pcq_status and records a thread-local message readable through pcq_last_error(); nothing throws or
 * aborts across the boundary; the caller owns every input buffer, the library owns device memory and
 * every array it hands back until the owning object is destroyed.  There is NO CPU fallback: every
 * entry point that scans points runs hand-written sm_100a CUDA kernels and fails with PCQ_ERR_CUDA
 * when no usable GPU is present.
 */
#ifndef PCQ_H
#define PCQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Status codes.  PCQ_ERR_FORMAT  ~ an `Err(anyhow!(..))` of the reference (bad header / bad format),
 *                PCQ_ERR_PANIC   ~ a place where the reference panics (AABB::from_min_max with
 *                                  min > max: las.rs:61, 88; last.rs:55, 98; main.rs:80),
 *                PCQ_ERR_GRID    ~ SparseGrid::new's "Too many cells" error (grid_sampling.rs:32-34),
 *                PCQ_ERR_ALIASED ~ no counterpart in the reference.  SparseGrid keys that suffer key aliasing
 *                                  (grid_sampling.rs:62-70 vs 78-82: the result for such a key is a sequential
 *                                  fold in scan order) are replayed exactly on one GPU and, by the pcq_group_*
 *                                  searches, across the GPUs of a box (their points are routed to the key's owner);
 *                                  the code is returned where that order cannot be honoured: point ranges fed to
 *                                  one collector out of scan order, a hashed table that must be re-hashed in the
 *                                  launch that meets the aliased key, or the bare pcq_grid_export_candidates on a
 *                                  collector that holds such keys.
 * ---------------------------------------------------------------------------------------------- */
enum pcq_status {
  PCQ_OK = 0,
  PCQ_ERR_ARG = -1,
  PCQ_ERR_FORMAT = -2,
  PCQ_ERR_PANIC = -3,
  PCQ_ERR_CUDA = -4,
  PCQ_ERR_NOMEM = -5,
  PCQ_ERR_GRID = -6,
  PCQ_ERR_ALIASED = -7,
  PCQ_ERR_IO = -8
};

/* readers::Point (readers/src/lib.rs:10-19): #[repr(C, packed)], 31 bytes,
 * position f64x3 @0, color u16x3 @24, classification u8 @30. */
typedef struct __attribute__((packed)) pcq_point {
  double pos[3];
  uint16_t rgb[3];
  uint8_t cls;
} pcq_point;

enum pcq_layout { PCQ_LAYOUT_LAS = 0, PCQ_LAYOUT_LAST = 1 };

/* What the scan needs from las::raw::Header / las::Header (las.rs:59-60, 73-78, 101-103;
 * last.rs:53-54, 80-90, 220-225): LAS 1.x public header block fields. */
typedef struct pcq_file_desc {
  uint8_t layout;          /* pcq_layout */
  uint8_t format;          /* point_data_record_format after optional `&= 0b1111` (last.rs:222) */
  uint16_t record_len;     /* point_data_record_length */
  uint32_t point_data_off; /* offset_to_point_data */
  uint64_t n_points;       /* header.number_of_points() */
  double scale[3];
  double offset[3];
  double hdr_min[3];       /* header.bounds().min */
  double hdr_max[3];       /* header.bounds().max */
} pcq_file_desc;

/* BoundsSearcher / ClassSearcher (searcher.rs:33-91, 94-152). */
enum pcq_query_kind { PCQ_QUERY_BOUNDS = 0, PCQ_QUERY_CLASS = 1 };
typedef struct pcq_query {
  uint8_t kind;   /* pcq_query_kind */
  uint8_t cls;    /* ClassSearcher::class; whole-byte compare (las.rs:229, last.rs:260) */
  uint8_t pad_[6];
  double qmin[3]; /* BoundsSearcher::bounds.min() */
  double qmax[3]; /* BoundsSearcher::bounds.max() */
} pcq_query;

/* CountCollector / BufferCollector / GridSampledCollector (collect_points.rs:72-98, 14-44, 100-127). */
enum pcq_collector_kind { PCQ_COLLECT_COUNT = 0, PCQ_COLLECT_BUFFER = 1, PCQ_COLLECT_GRID = 2 };

typedef struct pcq_ctx pcq_ctx;             /* one GPU + stream + scratch arenas (one per process/rank) */
typedef struct pcq_file pcq_file;           /* a file (or a point range of one) resident in HBM       */
typedef struct pcq_collector pcq_collector; /* device-resident ResultCollector                        */

/* ---- library --------------------------------------------------------------------------------- */
const char* pcq_last_error(void);
const char* pcq_version(void);

/* ---- host-side logic of the path (no GPU needed) ----------------------------------------------- */

/* parse_las_header + Header::from_raw (las.rs:33-36, 59-60; last.rs:36-39, 53-54, 220-223).
 * `mask_format` != 0 applies `point_data_record_format &= 0b1111` first, as the LAST class search and
 * LASTReader do (last.rs:222, last_reader.rs:76-79); the two bounds searches do not (last.rs:53-54). */
int pcq_parse_header(const void* bytes, size_t n_bytes, int layout, int mask_format, pcq_file_desc* out);

/* Query bounds -> integer coordinates in the local space of the file (las.rs:88-99, last.rs:98-109),
 * including the reference's use of x_scale_factor for min.y / min.z.  PCQ_ERR_PANIC when any
 * lo > hi (AABB::<i64>::from_min_max panics). */
int pcq_local_bounds(const pcq_file_desc* desc, const double qmin[3], const double qmax[3],
                     int64_t lo[3], int64_t hi[3]);

/* file_bounds.intersects(bounds) (las.rs:82, last.rs:92): closed-interval overlap on all axes.
 * Writes 0/1 to *out.  PCQ_ERR_PANIC when the header bounds have min > max (las.rs:61). */
int pcq_file_intersects(const pcq_file_desc* desc, const double qmin[3], const double qmax[3], int* out);

/* SparseGrid::new (grid_sampling.rs:18-47): cells per dimension and bits per dimension. */
int pcq_grid_params(const double gmin[3], const double gmax[3], double cell_size,
                    uint64_t dims[3], uint64_t bits[3]);

/* ---- device context ---------------------------------------------------------------------------- */
int pcq_ctx_create(int device, pcq_ctx** out);
void pcq_ctx_destroy(pcq_ctx* ctx);
/* Use a caller-owned stream (a cudaStream_t passed as void*) for all work of this context. */
int pcq_ctx_set_stream(pcq_ctx* ctx, void* cuda_stream);
int pcq_ctx_synchronize(pcq_ctx* ctx);
/* Scan kernel variant: 0 = auto, 1 = direct (vectorised global loads), 2 = staged (bulk-async tiles
 * into shared memory behind an mbarrier pipeline).  For measurement; results are identical. */
int pcq_ctx_set_scan_variant(pcq_ctx* ctx, int variant);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t pcq_ctx_launch_count(const pcq_ctx* ctx);

/* ---- files in HBM ------------------------------------------------------------------------------ */

/* Replaces open_file_reader's mmap (las.rs:24-31): copies the point range
 * [first_point, first_point + n_points) of a whole LAS/LAST file image from host memory to HBM
 * (256-byte aligned; for LAST only the position / classification / colour columns travel).
 * n_points == UINT64_MAX means "to the end of the file".  `ext` is "las" or "last". */
int pcq_file_stage_host(pcq_ctx* ctx, const void* file_bytes, size_t n_bytes, const char* ext,
                        uint64_t first_point, uint64_t n_points, pcq_file** out);

/* Wraps point data that is already resident in HBM.  For LAS `dev_point_data` points at record 0;
 * for LAST it points at the start of the transposed record block (column of the field at record
 * offset k starts at dev_point_data + k * desc->n_points, last_reader.rs:88-144).  The memory stays
 * owned by the caller, must not be rewritten while the file object (and its chunk index) is in use, and must come
 * from an allocator with at least 16-byte granularity (cudaMalloc, a framework's
 * caching allocator, ...): record tiles move with 16-byte bulk copies, so up to 15 bytes past the last record may be
 * read.  `first_point_index` is the scan index of record 0 inside its file. */
int pcq_file_wrap_device(pcq_ctx* ctx, const pcq_file_desc* desc, const void* dev_point_data,
                         uint64_t first_point_index, pcq_file** out);

/* Scan index of this range's record 0 within its collector (ties in the density fold keep the point
 * with the smaller scan index, i.e. the one the sequential reference loop meets first).  Default:
 * the number of points of all files fed to the collector before this one.  Set it explicitly when
 * files or point ranges of one logical scan are sharded over several GPUs. */
int pcq_file_set_scan_base(pcq_file* f, uint64_t scan_base);

int pcq_file_desc_get(const pcq_file* f, pcq_file_desc* out);
void pcq_file_release(pcq_file* f);

/* ---- collectors -------------------------------------------------------------------------------- */

/* kind = PCQ_COLLECT_GRID uses gmin/gmax/cell_size exactly as GridSampledCollector::new(bounds,
 * cell_size) (collect_points.rs:104-108; main.rs:253-264); other kinds ignore them (may be NULL).
 * A grid collector's cell table is allocated by the first search that feeds it, out of a third of the free HBM shared
 * by the grid collectors of that call, and covers the cells under the header boxes of the collector's files (cut by the
 * query box) when that is much less than the whole grid — creating 64 per-file collectors costs nothing until they
 * are used, and a file whose header understates its bounds only costs a second launch. */
int pcq_collector_create(pcq_ctx* ctx, int kind, const double gmin[3], const double gmax[3],
                         double cell_size, pcq_collector** out);
void pcq_collector_destroy(pcq_collector* c);
/* Forget everything collected so far (keeps allocations). */
int pcq_collector_reset(pcq_collector* c);
/* Host only (no GPU needed): the cells of SparseGrid(gmin, gmax, cell_size) that positions inside the box
 * [box_min, box_max] can fall into — per axis the first cell lo[a] and the number of cells n[a] (one cell of margin on
 * both sides, clamped to the axis' key mask).  This is the sub-box a grid collector's dense table is made for when it
 * is fed files whose header boxes (cut by the query box) lie inside the box; every position p with box_min <= p <=
 * box_max has lo[a] <= cell_a(p) < lo[a] + n[a] unless the cell exceeds the mask (such points are aliased and bypass
 * the table, grid_sampling.rs:62-70). */
int pcq_grid_cells_under_box(const double gmin[3], const double gmax[3], double cell_size, const double box_min[3],
                             const double box_max[3], uint64_t lo[3], uint64_t n[3]);
/* pcq_collector_reset for a list of collectors (e.g. the per-file collectors of run_search_parallel before the next
 * query): the count and buffer collectors of a context are cleared by ONE kernel launch instead of one stream operation
 * each; grid collectors are reset one by one. */
int pcq_collectors_reset(pcq_collector* const* collectors, uint32_t n);
/* ResultCollector::point_count (collect_points.rs:11). */
int pcq_collector_point_count(pcq_collector* c, uint64_t* out);
/* ResultCollector::points / points_ref (collect_points.rs:9-10): host array owned by the collector,
 * valid until the next call on it.  BUFFER: scan order.  GRID: arbitrary order (HashMap::values).
 * COUNT: *out_points = NULL, *out_n = 0 and the call returns PCQ_OK (`None`). */
int pcq_collector_points(pcq_collector* c, const pcq_point** out_points, uint64_t* out_n);
/* Same records, left in HBM (device pointer, 31-byte stride). */
int pcq_collector_points_device(pcq_collector* c, const void** out_dev_points, uint64_t* out_n);
/* The `-o` output of the query (FileDumper::dump_points, dump_points.rs:63-116) prepared on the device: the
 * collector's points as LAS 1.2 point format 2 records (26 bytes each: x, y, z = round((p - min) / scale) as i32,
 * return byte 0x09, classification, r, g, b) together with the header values of :74-88 — out_min = offset = smallest
 * position, out_max, out_scale = max(10^ceil(log10(max_extent / i32::MAX)), 0.001).  A min / max reduction and a
 * quantisation kernel replace the per-point host loops, and 26 instead of 31 bytes per record cross PCIe.
 * BUFFER: scan order; GRID: the order of pcq_collector_points.  *out_records is pinned host memory owned by the
 * collector, valid until the next call on it; *out_n == 0 (nothing to write, :65-67) leaves the other outputs untouched.
 * COUNT collectors: *out_n = 0. */
int pcq_collector_las_records(pcq_collector* c, double out_min[3], double out_max[3], double* out_scale,
                              const uint8_t** out_records, uint64_t* out_n);

/* ---- the scan ---------------------------------------------------------------------------------- */

/* Searcher::search_file over a batch of resident files with SearchImplementation::Optimized.
 * n_collectors == 1: every file feeds collectors[0] in `files` order (run_search_sequential,
 * main.rs:122-144).  n_collectors == n_files: file i feeds collectors[i] (run_search_parallel,
 * main.rs:146-183).  All collectors of one call must be of one kind.  Asynchronous on the
 * context's stream; collector getters synchronise. */
int pcq_search_files(pcq_ctx* ctx, pcq_file* const* files, uint32_t n_files, const pcq_query* query,
                     pcq_collector* const* collectors, uint32_t n_collectors);

/* Host-staged variant: whole file images in (ideally pinned) host memory are streamed through a ring
 * of HBM chunk buffers, copies overlapped with the scan; nothing stays resident.  A count search returns
 * while its last copies and scans are still in flight: `file_bytes` must stay valid (and unchanged) until a
 * collector of the call has been read or pcq_ctx_synchronize has returned. */
int pcq_search_host_files(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes,
                          const char* const* exts, uint32_t n_files, const pcq_query* query,
                          pcq_collector* const* collectors, uint32_t n_collectors);

/* Same, for a batch of queries over the same files: every chunk that crosses PCIe is scanned by each
 * query that needs its file while it is resident.  collectors[q * n_collectors_per_query + lane] is
 * the collector of query q (lane = 0 for sequential mode, = file index for parallel mode). */
int pcq_search_host_files_multi(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes,
                                const char* const* exts, uint32_t n_files, const pcq_query* queries, uint32_t n_queries,
                                pcq_collector* const* collectors, uint32_t n_collectors_per_query);

/* ---- on-the-fly chunk index ---------------------------------------------------------------------
 * The reference's own first idea for going faster (improvements.md:3-10): one header per chunk of
 * points holding the min/max of the queried attributes, consulted by later scans to find the chunks
 * that can hold a match.  Here a header covers PCQ_INDEX_CHUNK_POINTS consecutive points of a
 * resident file and holds the integer AABB of their x/y/z fields and a 256-bit set of the class
 * bytes that occur (the byte the class search compares, las.rs:202-212 / last.rs:245-259).  A search
 * over an indexed file launches the scan kernels over the surviving runs of chunks only; scan order,
 * scan indices and therefore every result are exactly those of the full scan.                     */
#define PCQ_INDEX_CHUNK_POINTS 8192u

typedef struct pcq_chunk_header {
  int32_t lo[3], hi[3];  /* min / max of the raw i32 x, y, z of the chunk                          */
  uint32_t cls_bits[8];  /* bit c set iff some point of the chunk has class byte c                 */
  uint32_t n_points;     /* points in the chunk (the last one may be short)                        */
  uint32_t pad_;
} pcq_chunk_header;      /* 64 bytes */

/* Builds the chunk headers of a resident file in one pass over its point data (k_chunk_index) and
 * keeps a host copy for the per-search filter.  Synchronises the context's stream.  Idempotent.   */
int pcq_file_build_index(pcq_file* f);
void pcq_file_drop_index(pcq_file* f);
/* Host copy of the headers (owned by the file, valid until it is released or the index dropped);
 * *out_n == 0 when the file has no index. */
int pcq_file_index(const pcq_file* f, const pcq_chunk_header** out_headers, uint64_t* out_n);
/* The filter the searches apply, as a host-only function (no GPU): the runs [first_chunk, end_chunk) of chunks of a
 * file in which `query` can find a match, runs less than `join_gap` chunks apart joined.  Writes the first `cap_runs`
 * runs as pairs into `runs`, their total number into *n_runs and the number of chunks that may match into *n_may.
 * Zero runs when the file's header box already excludes the query. */
int pcq_index_filter(const pcq_chunk_header* headers, uint64_t n_chunks, const pcq_file_desc* desc,
                     const pcq_query* query, uint64_t join_gap, uint64_t* runs, uint64_t cap_runs,
                     uint64_t* n_runs, uint64_t* n_may);
/* n > 0: pcq_search_files builds the index of a file by itself when it scans it for the (n+1)-th
 * time ("while scanning first (without an index) …, upon further scans …").  0 (default): never.  */
int pcq_ctx_set_auto_index(pcq_ctx* ctx, uint32_t after_n_scans);

/* What the last pcq_search_files / pcq_search_host_files* call of the context skipped (host-staged:
 * points_scanned = points that crossed PCIe). */
typedef struct pcq_scan_stats {
  uint64_t points_total;    /* points of the files that passed the per-file checks                 */
  uint64_t points_scanned;  /* points the kernels were launched over                               */
  uint64_t chunks_total;    /* chunks of the indexed files among them                              */
  uint64_t chunks_skipped;
  uint32_t segments;        /* point ranges of the launch                                          */
  uint32_t pad_;
} pcq_scan_stats;
int pcq_ctx_last_scan_stats(const pcq_ctx* ctx, pcq_scan_stats* out);

/* The same idea for file images that stay in host memory (pcq_search_host_files*): the headers are a
 * by-product of the first pass — every piece is indexed while it is resident for the scan, only for
 * the attributes whose columns that pass copies (LAS: both; LAST: positions for bounds queries, the
 * class column for class queries) — and later passes copy only the runs of chunks in which some
 * query of the batch can find a match, so fewer bytes cross PCIe.  `index` remembers the files by
 * position in the list: pass the same list every time.  Results equal pcq_search_host_files_multi. */
typedef struct pcq_host_index pcq_host_index;
int pcq_host_index_create(pcq_ctx* ctx, pcq_host_index** out);
void pcq_host_index_destroy(pcq_host_index* index);
int pcq_host_index_info(pcq_host_index* index, uint32_t file, uint64_t* n_chunks, int* has_box, int* has_cls);
int pcq_search_host_files_indexed(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes,
                                  const char* const* exts, uint32_t n_files, const pcq_query* queries, uint32_t n_queries,
                                  pcq_collector* const* collectors, uint32_t n_collectors_per_query,
                                  pcq_host_index* index);

/* Pinned host memory helpers for callers that want full-speed staging. */
int pcq_host_alloc(size_t n_bytes, void** out);
void pcq_host_free(void* p);

/* ---- multi-GPU density exchange (one process per GPU; the caller moves the bytes, e.g. with an
 *      NCCL all-to-all) ------------------------------------------------------------------------- */

/* One 64-byte candidate per locally occupied cell. */
typedef struct pcq_cell_candidate {
  uint64_t key;       /* SparseGrid cell key (grid_sampling.rs:68-70) */
  uint64_t dist_bits; /* f64 bits of the squared distance to the cell centre */
  uint64_t scan_idx;  /* global scan index: ties keep the first point, as the strict `<` does */
  pcq_point point;    /* 31 bytes */
  uint8_t pad_[9];
} pcq_cell_candidate;

/* Emits the local winners of a GRID collector partitioned by owner = mix64(key) % n_parts into one
 * device array; counts[p] candidates for part p, parts stored back to back. */
int pcq_grid_export_candidates(pcq_collector* c, uint32_t n_parts, const void** out_dev_candidates,
                               uint64_t* counts /* n_parts */);
/* Folds candidates received from peers (device pointer) into this collector's grid. */
int pcq_grid_import_candidates(pcq_collector* c, const void* dev_candidates, uint64_t n);

/* ---- multi-GPU: a group of GPUs of one box -----------------------------------------------------------------------
 * The reference's only parallelism is one rayon task per file (run_search_parallel, main.rs:146-183).  Here files AND
 * point ranges of files shard across the GPUs of one box.  Count queries need no exchange (the per-file counts are
 * summed on the host, or with one ncclAllReduce of n_files integers when every GPU has its own process); select
 * queries concatenate the per-(file, GPU) record streams in GPU order, which is the order one BufferCollector would
 * have seen; a max-density query has one exchange step: one candidate per locally occupied cell travels to the cell's
 * owner (mix64(key) % holders) with a grouped ncclSend / ncclRecv all-to-all over NVLink, ties break on the global
 * scan index, and the points of SparseGrid keys that suffer key aliasing (whose result is a sequential fold in scan
 * order, grid_sampling.rs:62-102) are routed to the key's owner and folded there in global scan order — so every
 * result equals what the reference computes over the whole dataset.
 *
 * A group is ONE process driving n GPUs (pcq_group_create; ncclCommInitAll) or one process per GPU
 * (pcq_group_create_rank; the launcher hands the id of pcq_group_unique_id to every rank).  In the second form every
 * entry point below is collective: all ranks call it with the same arguments.  NCCL is resolved at run time
 * (dlopen of libnccl.so.2, or the path in PCQ_NCCL_LIB); a group of one GPU does not need it.                   */
typedef struct pcq_group pcq_group;
typedef struct pcq_dataset pcq_dataset; /* files sharded over the members of a group, resident in HBM          */
typedef struct pcq_result pcq_result;   /* what the collectors of one group search hold, per lane               */

#define PCQ_GROUP_ID_BYTES 128

int pcq_group_create(const int* devices /* NULL: 0 .. n-1 */, uint32_t n_devices, pcq_group** out);
int pcq_group_unique_id(void* id_out /* PCQ_GROUP_ID_BYTES */);
int pcq_group_create_rank(int device, uint32_t rank, uint32_t world, const void* id, pcq_group** out);
void pcq_group_destroy(pcq_group* g);
uint32_t pcq_group_world(const pcq_group* g);
uint32_t pcq_group_local_count(const pcq_group* g);                      /* members driven by this process      */
uint32_t pcq_group_local_rank(const pcq_group* g, uint32_t local_index); /* their ranks in the group            */
pcq_ctx* pcq_group_ctx(pcq_group* g, uint32_t local_index);              /* their contexts (owned by the group) */
uint64_t pcq_group_launch_count(const pcq_group* g);
int pcq_group_synchronize(pcq_group* g);

/* How a dataset is cut (host only, no GPU).  PCQ_SHARD_RANGES: every file is cut into `world` contiguous ranges of
 * whole index chunks and rank r takes the r-th range of every file, so that a query which touches few files (doc-S:
 * 5 of 64 tiles) still spreads over all GPUs.  PCQ_SHARD_FILES: whole files, largest first onto the least loaded
 * rank (the reference's unit of parallelism).  Writes the first `cap` shards, ordered by (file, rank).          */
enum pcq_shard_mode { PCQ_SHARD_RANGES = 0, PCQ_SHARD_FILES = 1 };
typedef struct pcq_shard {
  uint32_t file;
  uint32_t rank;
  uint64_t first_point;
  uint64_t n_points;
} pcq_shard;
int pcq_shard_plan(const uint64_t* points_per_file, uint32_t n_files, uint32_t world, int mode, pcq_shard* out,
                   uint64_t cap, uint64_t* n_out);

/* Stages every member's shards of whole file images (host memory) into its HBM. */
int pcq_group_stage_host_files(pcq_group* g, const void* const* file_bytes, const size_t* n_bytes,
                               const char* const* exts, uint32_t n_files, int shard_mode, pcq_dataset** out);
/* A dataset of point ranges that are already resident (pcq_file_wrap_device / pcq_file_stage_host on the contexts of
 * pcq_group_ctx): piece i is files[i], a range of file file_index[i], held by local member local_member[i].  A
 * member holds at most one range of a file, pieces of a member in ascending file order, ranges of a file ascending
 * with the rank.  points_per_file describes the WHOLE dataset (all n_files files).  The files stay the caller's. */
int pcq_group_wrap_files(pcq_group* g, const uint64_t* points_per_file, uint32_t n_files, pcq_file* const* files,
                         const uint32_t* local_member, const uint32_t* file_index, uint32_t n_local,
                         pcq_dataset** out);
void pcq_dataset_release(pcq_dataset* ds);

/* run_search_sequential (per_file == 0: ONE collector over all files in dataset order, main.rs:122-144) or
 * run_search_parallel (per_file != 0: one collector per file, main.rs:146-183) of a batch of queries over a resident
 * dataset, with collectors of `collector_kind` (gmin / gmax / cell_size as in pcq_collector_create).  out[q] receives
 * the result of query q.  Count searches return before their kernels finish; pcq_result_counts waits.            */
int pcq_group_search(pcq_group* g, pcq_dataset* ds, const pcq_query* queries, uint32_t n_queries, int collector_kind,
                     const double gmin[3], const double gmax[3], double cell_size, int per_file, pcq_result** out);
/* The same over file images in host memory (pcq_search_host_files_multi per member, restricted to its shards): every
 * member streams its point ranges over its own PCIe link.  A process only reads the headers and the ranges of its own
 * members, so with one process per GPU an image needs to be populated (and pinned, pcq_host_register) only there.
 * Synchronous: the images may be released when the call returns.                                                  */
int pcq_group_search_host_files(pcq_group* g, const void* const* file_bytes, const size_t* n_bytes,
                                const char* const* exts, uint32_t n_files, const pcq_query* queries, uint32_t n_queries,
                                int collector_kind, const double gmin[3], const double gmax[3], double cell_size,
                                int per_file, int shard_mode, pcq_result** out);

/* ResultCollector::point_count per lane (1 lane, or one per file), group-wide. */
int pcq_result_counts(pcq_result* r, const uint64_t** counts, uint32_t* n_lanes);
/* ResultCollector::points of a lane: BUFFER in scan order, GRID in arbitrary order; COUNT: NULL / 0.  With one
 * process per GPU the records are gathered on rank 0 (other ranks get NULL / 0; the counts are known everywhere).
 * Pinned host memory that the result borrows from its group: release results before their group.               */
int pcq_result_points(pcq_result* r, uint32_t lane, const pcq_point** out_points, uint64_t* out_n);
void pcq_result_release(pcq_result* r);

/* Where the time of the last select / density search of the group went (host clock around synchronised phases, this
 * process's members; a count search does not synchronise and leaves them untouched). */
typedef struct pcq_group_stats {
  double scan_ms;            /* the members' scans of their ranges (first pass)                                   */
  double rescan_ms;          /* density: gathering the affected keys + the log-only second pass (0 when none)     */
  double export_ms;          /* density: one candidate per locally occupied cell, partitioned by owner            */
  double exchange_ms;        /* density: part sizes + the all-to-all over NVLink                                  */
  double import_ms;          /* density: owners fold what they received (+ the ordered replay of affected keys)   */
  double finalize_ms;        /* winners -> 31-byte records -> host lanes                                          */
  uint64_t cells_sent;       /* candidates that left their GPU                                                    */
  uint64_t log_entries_sent; /* points of affected keys that left their GPU                                       */
  uint64_t bytes_sent;       /* 64 bytes each                                                                     */
  uint64_t affected_keys;    /* SparseGrid keys with an order-dependent result, group-wide                        */
} pcq_group_stats;
int pcq_group_last_stats(const pcq_group* g, pcq_group_stats* out);

/* Pins (page-locks) a range of host memory the caller owns — e.g. the mapping of a file — so that the host-staged
 * searches copy from it at link speed instead of through the bounce ring. */
int pcq_host_register(void* p, size_t n_bytes);
int pcq_host_unregister(void* p);
/* Binds the calling thread (and the memory it touches first from now on) to the NUMA node the context's GPU hangs
 * off, so that pinned staging buffers are local to the GPU's PCIe root.  *out_node: the node, or -1 when the platform
 * does not say.  Linux only; a no-op elsewhere. */
int pcq_ctx_bind_host_thread(pcq_ctx* ctx, int* out_node);

#ifdef __cplusplus
}
#endif
#endif /* PCQ_H */
