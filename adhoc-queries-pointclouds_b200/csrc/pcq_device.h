// pcq_device.h — structures shared between the host side of libpcq and its CUDA kernels.
//
// Vocabulary: a *segment* is one point range of one file as the scan kernels see it; a *tile* is
// kTilePts consecutive records of one segment (the unit of scheduling, staging and of the
// decoupled look-back); a *lane* is one ResultCollector (collect_points.rs) — one lane for
// run_search_sequential, one lane per file for run_search_parallel (main.rs:122-183).
#pragma once
#include <cstdint>

#include "../../include/pcq.h"

namespace pcq {

constexpr int kTilePts = 512;  // records per tile
constexpr int kBlock = 256;    // threads per CTA of the scan kernels
constexpr int kPPT = kTilePts / kBlock;

// MODE_GRIDQ: MODE_GRID for sparse matches — matching points go through a CTA match queue (kernels.cu)
enum Mode : int { MODE_COUNT = 0, MODE_SELECT = 1, MODE_GRID = 2, MODE_GRIDQ = 3 };

// One point range of one file, as resident in HBM.
struct alignas(16) Segment {
  const uint8_t* rec;   // LAS: record 0 of the range.  LAST: positions column (12-byte stride)
  const uint8_t* cls;   // LAST: classification column of the range (LAS: unused)
  const uint8_t* rgb;   // LAST: colour column of the range, or nullptr (LAS: unused)
  uint64_t n_points;
  uint64_t first_tile;  // index of this segment's first tile within the launch
  uint64_t lane_first_tile;  // first tile of this segment's lane within the launch (look-back stops here)
  uint64_t scan_base;   // collector-wide scan index of point 0 of the range
  double scale[3];      // raw_header.{x,y,z}_scale_factor
  double offset[3];     // raw_header.{x,y,z}_offset
  int32_t lo[3];        // query_bounds_local.min(), clamped into i32 (las.rs:88-99)
  int32_t hi[3];        // query_bounds_local.max(), clamped into i32
  uint32_t record_len;  // LAS: point_data_record_length.  LAST: 12
  uint32_t lane;
  uint16_t cls_off;     // LAS: offset of the classification byte inside a record (15 / 16)
  int16_t rgb_off;      // LAS: offset of the colour inside a record, -1 if the format has none
  uint8_t layout;       // pcq_layout
  uint8_t align;        // 4, 2 or 1: alignment every x/y/z field of the range is guaranteed to have
  uint8_t rgb_align2;   // LAST: 1 when the colour column of the range is 2-byte aligned
  uint8_t pad_[1];
};

// 64-byte density candidate == pcq_cell_candidate of the C ABI.
struct alignas(16) Candidate {
  uint64_t key;
  uint64_t dist_bits;
  uint64_t scan_idx;
  uint8_t point[31];
  uint8_t pad_[9];
};
static_assert(sizeof(Candidate) == 64, "candidate must be 64 bytes");
static_assert(sizeof(pcq_cell_candidate) == 64, "ABI candidate must be 64 bytes");

// SparseGrid (grid_sampling.rs:9-15) as the insert kernel sees it.
struct GridDev {
  double bmin[3];
  double bmax[3];
  double dims_f[3];     // self.dimensions.{x,y,z} as f64
  double cell_size;
  uint64_t mask[3];     // (1 << bits) - 1
  uint32_t shift_y;     // bits.x
  uint32_t shift_z;     // bits.x + bits.y
  unsigned long long* table;  // per cell: f64 bits of the smallest squared distance seen (init ~0)
  uint64_t table_slots;       // dense: 1 << (bits.x + bits.y + bits.z).  hashed: power of two
  unsigned long long* hkeys;  // hashed only: cell key per slot (init ~0), nullptr when dense
  Candidate* cands;
  unsigned long long* cand_count;
  uint64_t cand_cap;
  uint32_t* flags;      // bit 0: (unused), bit 1: candidate arena overflow, bit 2: hash full, bit 3: replay log overflow
  // Key aliasing (grid_sampling.rs:62-70 vs 78-82): a point whose cell index exceeds its bit mask on some axis shares
  // the key of a low cell while its centre lies elsewhere, which makes the reference's fold for that key depend on
  // insertion order.  Such keys are "affected": their points bypass the atomic-min table, are logged, and are
  // replayed in scan order (k_alias_fold).  alias_keys is an open-addressing set of affected keys (~0 = empty),
  // alias_ord[slot] the ordinal of the key in the collector's state array.
  const unsigned long long* alias_keys;
  const uint32_t* alias_ord;
  uint64_t alias_slots;  // power of two, 0 when no key is affected yet
  uint32_t log_only;     // 1: second pass of a launch that met new affected keys — log their points, insert nothing
  uint32_t own_parts;    // finalisation: > 1 = only cells with mix64(key) % own_parts == own_me are finalists (multi-GPU owner)
  uint32_t own_me;
  // dense table over a sub-box of the grid (kernels.cu grid_slot): cells [sub_lo, sub_lo + sub_n) per axis, x fastest
  uint32_t sub_on;
  uint32_t pad_;
  uint64_t sub_lo[3];
  uint64_t sub_n[3];
  uint32_t pad2_;
  uint32_t fast_div;     // 1: every axis extent is a normal number in [2^-500, 2^500] — division by reciprocal (grid_math.cuh)
  double ext[3];         // bmax - bmin, the divisor of :51-57
  double inv_ext[3];     // RN(1 / ext), IEEE division on the host
  Candidate* log;        // replay log: {key, -, scan index, point} of every point of an affected key / aliased point
  unsigned long long* log_count;
  uint64_t log_cap;
};

// candidates are appended through warp-private chunks of this many arena slots (kernels.cu); a launch can leave up
// to one partly used chunk per warp and lane behind, which the host adds to every capacity estimate
constexpr uint32_t kCandChunk = 64;
constexpr uint32_t kGridCtasPerSm = 4;  // upper bound of resident scan CTAs per SM in grid mode

// k_select_ring: rows of 32 records per consumer warp (16 of them) and unit, chosen so that a unit is 80-93 KB
// (an even number of rows: the emit works in rounds of 64 records)
constexpr uint32_t sel_ring_rows(uint32_t R) { return R <= 12 ? 14u : (R <= 20 ? 8u : (R <= 28 ? 6u : 4u)); }
constexpr uint32_t sel_ring_unit_points(uint32_t R) { return 16u * 32u * sel_ring_rows(R); }

// Look-back descriptors are spread out: one descriptor every kDescStride 8-byte words, so that the few hundred
// descriptors every look-back warp of the GPU is polling at any moment do not all live in a handful of L2 lines.
#ifndef PCQ_DESC_STRIDE
#define PCQ_DESC_STRIDE 4
#endif
constexpr uint32_t kDescStride = PCQ_DESC_STRIDE;

constexpr uint32_t kFlagCandOverflow = 2u;
constexpr uint32_t kFlagHashFull = 4u;
constexpr uint32_t kFlagLogOverflow = 8u;

// state of one affected key: what SparseGrid's HashMap holds for it after the points replayed so far
struct AliasState {
  uint8_t point[31];
  uint8_t valid;
};
static_assert(sizeof(AliasState) == 32, "alias state must be 32 bytes");

// Per-lane (per-collector) device state for one launch.
struct LaneDev {
  unsigned long long* count;  // matches (COUNT / BUFFER collectors)
  uint8_t* out;               // BUFFER: 31-byte records
  uint64_t out_base;          // BUFFER: records already in `out` before this launch
  uint64_t out_cap;           // BUFFER: capacity of `out` in records
  GridDev grid;               // GRID
};

struct ScanParams {
  const Segment* segs;
  uint32_t n_segs;
  uint32_t query_kind;  // pcq_query_kind
  uint32_t cls;         // class byte of a class query
  uint32_t tile_pts;    // records per scheduling unit: kTilePts, or a whole look-back group for MODE_SELECT
  uint64_t n_tiles;
  const LaneDev* lanes;
  unsigned long long* tile_state;  // MODE_SELECT: decoupled look-back descriptors (n_tiles, zeroed)
  unsigned long long* ticket;      // MODE_SELECT: tile ticket counter (zeroed)
  uint32_t grid_sparse;            // MODE_GRID: 1 = few points are expected to match: use the match-queue kernels
  uint32_t sel_bytes;              // MODE_SELECT: 1 = k_select_bytes (LAST class query, 32768-point units)
  uint32_t sel_ring;               // MODE_SELECT: record length when the launch qualifies for k_select_ring, else 0
  uint32_t one_grid;               // MODE_GRID: 1 = the launch has one collector; its grid is grid0 (constant bank operands)
  GridDev grid0;
  uint32_t debug;                  // measurement hooks, only in -DPCQ_DEBUG_HOOKS builds: 1 = skip the look-back, 2 = skip the emit
};

// launch wrappers implemented in kernels.cu (stream is a cudaStream_t); 0 = ok, < 0 = CUDA error
bool staged_supports(uint32_t record_len);
// select_bytes: MODE_SELECT launch of a class query over LAST segments whose class columns are 16-byte aligned
uint32_t tile_points(int variant, int mode, uint32_t record_len, bool select_bytes);
int launch_scan(int variant, int mode, const ScanParams& p, uint32_t uniform_record_len, int min_align, int sm_count,
                void* stream);
int launch_class_count_soa(const ScanParams& p, int sm_count, void* stream);
int launch_grid_prune(const GridDev& g, uint64_t n_in, Candidate* dst, unsigned long long* dst_count, int sm_count,
                      void* stream);
// in-place finalisation of a density table (kernels.cu): [finalists] list the candidates that sit at their cell's
// minimum distance, [final_min] the smallest scan index among them wins the cell (the table then holds winner codes
// instead of distances), [emit], [final_restore] the finalists put the distances back
int launch_grid_finalists(const GridDev& g, uint64_t n, uint32_t* list, unsigned long long* list_count, int sm_count, void* stream);
int launch_grid_final_min(const GridDev& g, uint64_t n_max, const uint32_t* list, const unsigned long long* list_count, int sm_count,
                          void* stream);
int launch_grid_final_restore(const GridDev& g, uint64_t n_max, const uint32_t* list, const unsigned long long* list_count, int sm_count,
                              void* stream);
// over the finalists; mode 0: count winners per owner part, 1: write winners as candidates into their parts, 2: write 31-byte points
int launch_grid_emit(const GridDev& g, uint64_t n_max, const uint32_t* list, const unsigned long long* list_count, int mode,
                     uint32_t n_parts, unsigned long long* part_counts, unsigned long long* part_cursor, Candidate* out_cands,
                     uint8_t* out_points, unsigned long long* out_count, int sm_count, void* stream);
int launch_gather_blocks(const LaneDev* lanes, uint32_t n, uint32_t block_bytes, void* out, void* stream);
int launch_grid_import(const GridDev& g, const Candidate* in, uint64_t n, int sm_count, void* stream);

// alias replay (alias.cu).  All asynchronous on `stream` unless stated otherwise.
// best earlier candidate (scan index < before_scan) of each key listed in `keys` -> states[ord0 + i]
int alias_prewinners(const GridDev& g, uint64_t n_cands, const unsigned long long* d_keys, uint32_t n_keys,
                     unsigned long long before_scan, AliasState* d_states, uint32_t ord0, int sm_count, void* stream);
// sort the first n log entries by (key, scan index) and fold them into the states, one affected key per thread;
// synchronises the stream (temporary storage is freed before returning)
int alias_replay(const GridDev& g, uint64_t n, AliasState* d_states, int sm_count, void* stream);

// chunk index build (index.cu): one pcq_chunk_header per PCQ_INDEX_CHUNK_POINTS points of a resident file range
struct ChunkIndexArgs {
  const uint8_t* rec;   // LAS: record 0.  LAST: positions column
  const uint8_t* cls;   // LAST: class column (LAS: unused)
  uint64_t n_points;
  uint64_t n_chunks;
  uint32_t record_len;  // LAS only
  uint32_t cls_off;     // LAS only: the byte the class search compares (15 / 16)
  uint8_t layout;
  uint8_t align;        // alignment of the x/y/z fields: 4, 2 or 1
  uint8_t parts;        // kIndexPartBox | kIndexPartCls: a LAST pass may hold only one of the two columns; the part
                        // that is not computed is written as "anything may be here" (full i32 range / every class)
};
constexpr uint8_t kIndexPartBox = 1, kIndexPartCls = 2;
int launch_chunk_index(const ChunkIndexArgs& a, pcq_chunk_header* out, int sm_count, void* stream);

}  // namespace pcq
