// pcq_internal.hpp — host-side objects behind the opaque handles of include/pcq.h, shared by api.cu (one GPU) and
// group.cu (a group of GPUs).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "host_logic.hpp"
#include "pcq_device.h"

namespace pcq {

constexpr int kUploadSlots = 8;
constexpr int kChunkBuffers = 3;

struct UploadSlot {
  void* host = nullptr;
  void* dev = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
};

// device-side scalars of one collector
struct DevBlock {
  unsigned long long count;       // matches (COUNT / BUFFER)
  unsigned long long cand_count;  // GRID: candidates appended (may exceed capacity on overflow)
  unsigned long long out_count;   // GRID finalisation: winners emitted
  uint32_t flags;
  uint32_t pad_;
  unsigned long long log_count;   // GRID: replay-log entries written by the last launch (affected keys, alias.cu)
  unsigned long long fin_count;   // GRID finalisation: entries of the finalist list
};

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace pcq

using pcq::AliasState;
using pcq::Candidate;
using pcq::DevBlock;
using pcq::GridDev;
using pcq::UploadSlot;
using pcq::kChunkBuffers;
using pcq::kUploadSlots;

struct pcq_ctx {
  // files and collectors keep their context alive: pcq_ctx_destroy only marks it closing while any exist
  int refs = 0;
  bool closing = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int variant = 0;
  uint64_t launches = 0;
  UploadSlot slots[kUploadSlots];
  int next_slot = 0;
  std::vector<uint8_t> upload_tmp;  // staging of upload2
  // MODE_SELECT scratch
  unsigned long long* tile_state = nullptr;  // [0] = ticket, [1..] = descriptors
  uint64_t tile_state_cap = 0;
  // scalar blocks of many collectors travel in one copy
  void* d_gather = nullptr;
  void* h_gather = nullptr;
  size_t gather_cap = 0;
  // GRID export scratch (one export at a time)
  unsigned long long* part_scratch = nullptr;  // 2 * n_parts counters
  uint32_t part_scratch_cap = 0;
  // host-staged streaming
  cudaStream_t copy_stream = nullptr;
  void* chunk[kChunkBuffers] = {nullptr, nullptr, nullptr};
  size_t chunk_cap = 0;
  cudaEvent_t chunk_copied[kChunkBuffers] = {nullptr, nullptr, nullptr};
  cudaEvent_t chunk_free[kChunkBuffers] = {nullptr, nullptr, nullptr};
  // pinned bounce ring for file images in pageable memory (mmap'ed files): host threads copy a piece in, the copy
  // engine takes it from there at link speed
  void* bounce[kChunkBuffers] = {nullptr, nullptr, nullptr};
  size_t bounce_cap = 0;
  cudaEvent_t bounce_done[kChunkBuffers] = {nullptr, nullptr, nullptr};
  bool bounce_busy[kChunkBuffers] = {false, false, false};  // bounce_done[b] was recorded and not yet waited for
  // chunk index
  void* index_scratch = nullptr;  // device headers of the file being indexed (grow-only)
  void* index_bounce = nullptr;   // pinned landing buffer of their copy to the host
  size_t index_scratch_cap = 0;
  uint32_t auto_index_after = 0;  // 0 = never build one unasked
  pcq_scan_stats stats{};
};

struct pcq_file {
  pcq_ctx* ctx = nullptr;
  pcq_file_desc desc{};
  uint8_t raw_format = 0;
  uint64_t first_point = 0;  // index inside the file of record 0 of this range
  uint64_t n_points = 0;     // points in this range
  void* owned = nullptr;     // device allocation owned by this object (staged files)
  const uint8_t* rec = nullptr;
  const uint8_t* cls = nullptr;
  const uint8_t* rgb = nullptr;
  bool has_scan_base = false;
  uint64_t scan_base = 0;
  // chunk index (index.cu): host copy of the headers (what the per-search filter walks), scans seen so far
  std::vector<pcq_chunk_header> index;
  uint32_t scans = 0;
};

// Chunk headers of a list of file images that live in host memory (pcq_search_host_files_indexed).  The two parts of
// a header are kept apart because a pass only sees the columns its queries made it copy: `box` carries lo/hi, `cls`
// the class set.  A building pass writes the headers of a file into its device array behind the scans; the next
// indexed search fetches them into the host vectors the filter walks.
struct pcq_host_index {
  pcq_ctx* ctx = nullptr;
  struct File {
    uint64_t n_points = 0;
    uint64_t n_chunks = 0;
    uint64_t key = 0;                       // hash of (image address, length, parsed header): what the headers were made from
    pcq_chunk_header* d_headers = nullptr;  // n_chunks headers, written by k_chunk_index
    uint8_t unfetched = 0;                  // parts (kIndexPartBox | kIndexPartCls) of d_headers not yet in the vectors
    std::vector<pcq_chunk_header> box, cls;
    bool has_box = false, has_cls = false;
  };
  std::vector<File> files;
  bool pending = false;  // some file has unfetched headers (their kernels may still be running)
};

struct pcq_collector {
  pcq_ctx* ctx = nullptr;
  int kind = 0;
  DevBlock* dev = nullptr;
  uint64_t scan_total = 0;  // points of all files fed so far (scan index of the next file)
  // BUFFER
  uint8_t* d_out = nullptr;
  uint64_t out_len = 0, out_cap = 0;
  // GRID
  double gmin[3]{}, gmax[3]{}, cell = 0;
  uint64_t dims[3]{}, bits[3]{};
  GridDev grid{};
  uint64_t cand_len = 0;
  uint8_t* d_final = nullptr;
  uint64_t final_cap = 0, final_n = 0;
  bool final_valid = false;
  uint64_t total_bits = 0;           // GRID: key bits (bits.x + bits.y + bits.z)
  bool table_holds_winners = false;  // finalised in place: the cells of the winners hold scan indices until grid_restore
  // key-aliasing replay (alias.cu): affected keys in ordinal order, their fold states (host copy is authoritative
  // between launches), the device-side set and the replay log
  std::vector<uint64_t> akeys;
  std::vector<AliasState> astates;
  unsigned long long* d_akeys = nullptr;
  uint32_t* d_aord = nullptr;
  uint64_t a_slots = 0, a_slots_cap = 0;
  AliasState* d_astates = nullptr;
  uint64_t d_astates_cap = 0;
  Candidate* d_log = nullptr;
  uint64_t log_cap = 0;
  uint64_t scan_hi = 0;       // end of the highest point range fed so far: the replay needs launches in scan order
  uint64_t prune_epoch = 0;   // bumped whenever candidates are dropped (prune / rehash)
  // group-wide replay of affected keys (group.cu): log-only passes append their launch logs here
  int pass_mode = 0;  // 0 normal, 1 log-only (points of affected keys -> d_rawlog, nothing inserted), 2 sit this pass out
  Candidate* d_rawlog = nullptr;
  uint64_t rawlog_len = 0, rawlog_cap = 0;
  // finalisation keeps only the cells this collector owns: mix64(key) % own_parts == own_me (own_parts <= 1: all)
  uint32_t own_parts = 0, own_me = 0;
  // export scratch
  Candidate* d_export = nullptr;
  uint64_t export_cap = 0;
  // finalisation: indices of the candidates that sit at their cell's minimum distance (count: dev->fin_count)
  uint32_t* d_finlist = nullptr;
  uint64_t finlist_cap = 0;
  uint64_t fin_max = 0;  // candidates the list was made from (upper bound of its length, sizes the launches)
  // host copy of points()
  void* h_pts = nullptr;
  uint64_t h_cap = 0;
};


// ---- internals of api.cu that group.cu drives (one call per member context) ------------------------------------
namespace pcq {

struct HostRange {
  uint64_t first_point;  // multiple of PCQ_INDEX_CHUNK_POINTS
  uint64_t n_points;     // 0: this member holds nothing of the file
  uint64_t scan_base;    // scan index of the FILE's point 0 within its collector
};

int use_device(pcq_ctx* ctx);
int upload(pcq_ctx* ctx, const void* src, size_t bytes, void** dev_out);
int upload2(pcq_ctx* ctx, const void* a, size_t na, const void* b, size_t nb, void** dev_a, void** dev_b);
GridDev grid_view(const pcq_collector* c);
int grid_restore(pcq_collector* c);
int ensure_grid_tables(pcq_collector* const* cols, uint32_t n, const double* boxes);
int grow_log(pcq_collector* c, uint64_t need);
int alias_upload(pcq_collector* c);
int grid_finalize(pcq_collector* c);
// writer.cu: scale of the -o output (dump_points.rs:81-88)
double las_writer_scale(double max_extent);
int reset_collectors_batched(pcq_ctx* ctx, pcq_collector* const* cols, size_t n);
// the host-staged scan of pcq_search_host_files*; `ranges` (one per file, or nullptr) restricts it to point ranges
int search_host_multi(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes, const char* const* exts,
                      uint32_t n_files, const pcq_query* queries, uint32_t n_queries, pcq_collector* const* collectors,
                      uint32_t n_collectors, pcq_host_index* hix, const HostRange* ranges);

}  // namespace pcq
