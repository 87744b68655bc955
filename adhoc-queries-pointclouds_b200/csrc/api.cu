// api.cu — C ABI of libpcq: device context, files resident in HBM, device-resident collectors and
// the orchestration of one Searcher::search_file batch (searcher.rs:24-31; main.rs:122-183).
//
// Host work per call is O(files): validate like the reference does before its per-point loop
// (las.rs:59-99, 199-219; last.rs:53-109, 220-250), build one Segment per file, upload the table,
// launch ONE kernel for the whole batch.  All per-point work is in kernels.cu; there is no CPU scan.
#include <cuda_runtime.h>

#ifdef __linux__
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif

#include <algorithm>
#include <cctype>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <unordered_set>
#include <vector>

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

namespace pcq {  // (external linkage: group.cu drives several contexts through these, see pcq_internal.hpp)

int use_device(pcq_ctx* ctx) {
  CU(cudaSetDevice(ctx->device));
  return PCQ_OK;
}

// copy a small parameter table to the device through a ring of pinned slots (fully asynchronous)
// two tables of one launch (its segments and its collectors) in ONE host-to-device copy: a small search is a handful of
// microseconds of kernel behind a sequence of stream operations, each of which costs about as much
int upload2(pcq_ctx* ctx, const void* a, size_t na, const void* b, size_t nb, void** dev_a, void** dev_b) {
  const size_t off_b = round_up(na, 256);
  std::vector<uint8_t>& tmp = ctx->upload_tmp;
  tmp.resize(off_b + nb);
  std::memcpy(tmp.data(), a, na);
  std::memcpy(tmp.data() + off_b, b, nb);
  void* d = nullptr;
  RC(upload(ctx, tmp.data(), tmp.size(), &d));
  *dev_a = d;
  *dev_b = static_cast<uint8_t*>(d) + off_b;
  return PCQ_OK;
}

int upload(pcq_ctx* ctx, const void* src, size_t bytes, void** dev_out) {
  UploadSlot& s = ctx->slots[ctx->next_slot];
  ctx->next_slot = (ctx->next_slot + 1) % kUploadSlots;
  if (s.pending) {
    CU(cudaEventSynchronize(s.ev));
    s.pending = false;
  }
  if (s.cap < bytes) {
    if (s.host) cudaFreeHost(s.host);
    if (s.dev) cudaFree(s.dev);
    s.host = s.dev = nullptr;
    s.cap = 0;
    size_t cap = round_up(std::max<size_t>(bytes, 16384), 4096);
    CU(cudaMallocHost(&s.host, cap));
    CU(cudaMalloc(&s.dev, cap));
    s.cap = cap;
  }
  if (!s.ev) CU(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  std::memcpy(s.host, src, bytes);
  CU(cudaMemcpyAsync(s.dev, s.host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaEventRecord(s.ev, ctx->stream));
  s.pending = true;
  *dev_out = s.dev;
  return PCQ_OK;
}

int read_devblock(pcq_collector* c, DevBlock* out) {
  CU(cudaMemcpyAsync(out, c->dev, sizeof(DevBlock), cudaMemcpyDeviceToHost, c->ctx->stream));
  CU(cudaStreamSynchronize(c->ctx->stream));
  return PCQ_OK;
}

// the scalar blocks of all collectors of a launch: one gather kernel + one copy when there are many lanes
int read_devblocks(pcq_ctx* ctx, pcq_collector* const* collectors, uint32_t n, const LaneDev* d_lanes, std::vector<DevBlock>& out) {
  out.resize(n);
  if (n <= 4 || d_lanes == nullptr) {
    for (uint32_t l = 0; l < n; ++l)
      CU(cudaMemcpyAsync(&out[l], collectors[l]->dev, sizeof(DevBlock), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PCQ_OK;
  }
  const size_t bytes = (size_t)n * sizeof(DevBlock);
  if (ctx->gather_cap < bytes) {
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
    ctx->d_gather = ctx->h_gather = nullptr;
    ctx->gather_cap = 0;
    const size_t cap = round_up(bytes, 4096);
    CU(cudaMalloc(&ctx->d_gather, cap));
    CU(cudaMallocHost(&ctx->h_gather, cap));
    ctx->gather_cap = cap;
  }
  if (launch_gather_blocks(d_lanes, n, (uint32_t)sizeof(DevBlock), ctx->d_gather, ctx->stream) != 0)
    return fail(PCQ_ERR_CUDA, "k_gather_blocks launch failed");
  CU(cudaMemcpyAsync(ctx->h_gather, ctx->d_gather, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  std::memcpy(out.data(), ctx->h_gather, bytes);
  return PCQ_OK;
}

// host-side copy of a piece into the pinned bounce ring, split over a few threads (one thread moves ~10 GB/s)
void parallel_memcpy(void* dst, const void* src, size_t n) {
  unsigned hw = std::thread::hardware_concurrency();
  size_t threads = std::min<size_t>({8, hw ? hw / 2 : 2, n / (8u << 20) + 1});
  if (threads <= 1) {
    std::memcpy(dst, src, n);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = round_up((n + threads - 1) / threads, 4096);
  for (size_t t = 1; t < threads; ++t) {
    const size_t a = t * per;
    if (a >= n) break;
    const size_t len = std::min(per, n - a);
    pool.emplace_back([=] { std::memcpy(static_cast<uint8_t*>(dst) + a, static_cast<const uint8_t*>(src) + a, len); });
  }
  std::memcpy(dst, src, std::min(per, n));
  for (std::thread& t : pool) t.join();
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

int layout_of_ext(const char* ext) {
  if (!ext) return -1;
  if (std::strcmp(ext, "las") == 0) return PCQ_LAYOUT_LAS;
  if (std::strcmp(ext, "last") == 0) return PCQ_LAYOUT_LAST;
  return -1;
}

uint32_t cls_offset_in_record(uint8_t format) { return format <= 5 ? 15u : 16u; }  // las.rs:202-205, last.rs:69-71
int rgb_offset_in_record(uint8_t format) {                                         // las.rs:38-45
  switch (format) {
    case 2: return 20;
    case 3: return 28;
    case 5: return 28;
    default: return -1;
  }
}

uint8_t field_alignment(const void* base, uint32_t stride) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(base);
  if ((a & 3u) == 0 && (stride & 3u) == 0) return 4;
  if ((a & 1u) == 0 && (stride & 1u) == 0) return 2;
  return 1;
}

struct SegmentPlan {
  bool skip = false;  // file cannot contribute (early-out of las.rs:82-84 or empty integer range)
  int32_t lo[3]{}, hi[3]{};
};

// Everything the reference does between opening a file and entering its per-point loop.
int plan_file(const pcq_file_desc& d, uint8_t raw_format, const pcq_query* q, SegmentPlan* plan) {
  plan->skip = false;
  if (q->kind == PCQ_QUERY_BOUNDS) {
    // neither bounds search masks the format byte (las.rs:59-60, last.rs:53-54)
    if (raw_format > 10) return fail(PCQ_ERR_FORMAT, "Invalid LAS format %u", raw_format);
    int hit = 0;
    RC(file_intersects(&d, q->qmin, q->qmax, &hit));
    if (!hit) {
      plan->skip = true;
      return PCQ_OK;
    }
    int64_t lo[3], hi[3];
    RC(local_bounds(&d, q->qmin, q->qmax, lo, hi));
    for (int i = 0; i < 3; ++i) {
      if (lo[i] > INT32_MAX || hi[i] < INT32_MIN) plan->skip = true;  // no i32 coordinate can match
      plan->lo[i] = (int32_t)std::max<int64_t>(lo[i], INT32_MIN);
      plan->hi[i] = (int32_t)std::min<int64_t>(hi[i], INT32_MAX);
    }
    return PCQ_OK;
  }
  if (q->kind == PCQ_QUERY_CLASS) {
    // LAS class search does not mask (las.rs:199-212); LAST class search does (last.rs:222)
    if (d.layout == PCQ_LAYOUT_LAS && raw_format > 10) return fail(PCQ_ERR_FORMAT, "Invalid LAS format %u", raw_format);
    return PCQ_OK;
  }
  return fail(PCQ_ERR_ARG, "unknown query kind %u", q->kind);
}

void fill_segment(Segment* s, const pcq_file_desc& d, const uint8_t* rec, const uint8_t* cls, const uint8_t* rgb,
                  uint64_t n_points, const SegmentPlan& plan, uint32_t lane, uint64_t scan_base, uint32_t query_kind) {
  std::memset(s, 0, sizeof(*s));
  s->rec = rec;
  s->cls = cls;
  s->rgb = rgb;
  s->n_points = n_points;
  s->scan_base = scan_base;
  for (int i = 0; i < 3; ++i) {
    s->scale[i] = d.scale[i];
    s->offset[i] = d.offset[i];
    s->lo[i] = plan.lo[i];
    s->hi[i] = plan.hi[i];
  }
  s->lane = lane;
  s->layout = d.layout;
  if (d.layout == PCQ_LAYOUT_LAS) {
    s->record_len = d.record_len;
    // the LAS bounds search skips 3 bytes after z and reads "the" classification at +15 whatever the
    // format (las.rs:80, 121-124); the LAS class search uses +16 for formats 6..10 (las.rs:202-205)
    s->cls_off = (uint16_t)(query_kind == PCQ_QUERY_BOUNDS ? 15u : cls_offset_in_record(d.format));
    s->rgb_off = (int16_t)rgb_offset_in_record(d.format);
    s->align = field_alignment(rec, d.record_len);
  } else {
    s->record_len = 12;
    s->cls_off = 0;
    s->rgb_off = -1;
    s->align = field_alignment(rec, 12);
    s->rgb_align2 = (reinterpret_cast<uintptr_t>(rgb) & 1u) == 0 ? 1 : 0;
  }
}

int ensure_tile_state(pcq_ctx* ctx, uint64_t n_tiles) {
  const uint64_t need = n_tiles * kDescStride + 2 * kDescStride;
  if (ctx->tile_state_cap < need) {
    if (ctx->tile_state) cudaFree(ctx->tile_state);
    ctx->tile_state = nullptr;
    ctx->tile_state_cap = 0;
    uint64_t cap = std::max<uint64_t>(need, 1u << 16);
    CU(cudaMalloc(&ctx->tile_state, cap * sizeof(unsigned long long)));
    ctx->tile_state_cap = cap;
  }
  CU(cudaMemsetAsync(ctx->tile_state, 0, need * sizeof(unsigned long long), ctx->stream));
  return PCQ_OK;
}

int grow_out(pcq_collector* c, uint64_t need_records) {
  if (c->out_cap >= need_records) return PCQ_OK;
  uint64_t cap = std::max<uint64_t>(need_records, 4096);
  uint8_t* nb = nullptr;
  CU(cudaMalloc(&nb, cap * 31ull + 64));
  if (c->d_out) {
    if (c->out_len) CU(cudaMemcpyAsync(nb, c->d_out, c->out_len * 31ull, cudaMemcpyDeviceToDevice, c->ctx->stream));
    CU(cudaStreamSynchronize(c->ctx->stream));
    cudaFree(c->d_out);
  }
  c->d_out = nb;
  c->out_cap = cap;
  return PCQ_OK;
}

GridDev grid_view(const pcq_collector* c) {
  GridDev g = c->grid;
  g.cand_count = &c->dev->cand_count;
  g.flags = &c->dev->flags;
  g.alias_keys = c->a_slots ? c->d_akeys : nullptr;
  g.alias_ord = c->a_slots ? c->d_aord : nullptr;
  g.alias_slots = c->a_slots;
  g.log_only = c->pass_mode == 1 ? 1u : 0u;
  g.own_parts = c->own_parts;
  g.own_me = c->own_me;
  g.log = c->d_log;
  g.log_count = &c->dev->log_count;
  g.log_cap = c->log_cap;
  return g;
}

// a finalisation leaves the winners' scan indices in the table; before anything reads or updates distances again
// the finalists write theirs back (phase 3 of k_grid_final_phase)
int grid_restore(pcq_collector* c) {
  if (!c->table_holds_winners) return PCQ_OK;
  GridDev g = grid_view(c);
  if (launch_grid_final_restore(g, c->fin_max, c->d_finlist, &c->dev->fin_count, c->ctx->sm_count, c->ctx->stream) != 0)
    return fail(PCQ_ERR_CUDA, "k_grid_final_restore launch failed");
  c->ctx->launches++;
  c->table_holds_winners = false;
  return PCQ_OK;
}

// finalisation, first half: list the finalists of the first n candidates and let the smallest scan index win each cell.
// Afterwards the table holds winner codes (table_holds_winners) until grid_restore.
int grid_pick_winners(pcq_collector* c, const GridDev& g, uint64_t n) {
  pcq_ctx* ctx = c->ctx;
  if (n >> 32) return fail(PCQ_ERR_NOMEM, "density finalisation: %llu candidates exceed the 32-bit finalist list", (unsigned long long)n);
  if (c->finlist_cap < n) {
    if (c->d_finlist) cudaFree(c->d_finlist);
    c->d_finlist = nullptr;
    c->finlist_cap = 0;
    const uint64_t cap = n + n / 4 + 4096;
    if (cudaMalloc(&c->d_finlist, cap * sizeof(uint32_t)) != cudaSuccess) {
      cudaGetLastError();
      return fail(PCQ_ERR_NOMEM, "cannot allocate the finalist list of %llu entries", (unsigned long long)cap);
    }
    c->finlist_cap = cap;
  }
  CU(cudaMemsetAsync(&c->dev->fin_count, 0, sizeof(unsigned long long), ctx->stream));
  c->fin_max = n;
  if (launch_grid_finalists(g, n, c->d_finlist, &c->dev->fin_count, ctx->sm_count, ctx->stream) != 0)
    return fail(PCQ_ERR_CUDA, "k_grid_finalists launch failed");
  c->table_holds_winners = true;  // from here on (also if a later launch fails); the distances go back lazily
                                  // (grid_restore): usually nothing follows
  if (launch_grid_final_min(g, n, c->d_finlist, &c->dev->fin_count, ctx->sm_count, ctx->stream) != 0)
    return fail(PCQ_ERR_CUDA, "k_grid_final_min launch failed");
  ctx->launches += 2;
  return PCQ_OK;
}

int grow_log(pcq_collector* c, uint64_t need) {
  if (c->log_cap >= need) return PCQ_OK;
  CU(cudaStreamSynchronize(c->ctx->stream));
  if (c->d_log) cudaFree(c->d_log);
  c->d_log = nullptr;
  c->log_cap = 0;
  const uint64_t cap = std::max<uint64_t>(need, 1u << 14);
  if (cudaMalloc(&c->d_log, cap * sizeof(Candidate)) != cudaSuccess) {
    cudaGetLastError();
    return fail(PCQ_ERR_NOMEM, "cannot allocate a replay log of %llu entries", (unsigned long long)cap);
  }
  c->log_cap = cap;
  return PCQ_OK;
}

// (re)build the device-side set of affected keys and upload the host copy of their states
int alias_upload(pcq_collector* c) {
  pcq_ctx* ctx = c->ctx;
  const uint64_t n = c->akeys.size();
  if (n == 0) {
    c->a_slots = 0;
    return PCQ_OK;
  }
  uint64_t slots = 64;
  while (slots < 4 * n) slots <<= 1;
  CU(cudaStreamSynchronize(ctx->stream));
  if (c->a_slots_cap < slots) {
    if (c->d_akeys) cudaFree(c->d_akeys);
    if (c->d_aord) cudaFree(c->d_aord);
    c->d_akeys = nullptr;
    c->d_aord = nullptr;
    c->a_slots_cap = 0;
    CU(cudaMalloc(&c->d_akeys, slots * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_aord, slots * sizeof(uint32_t)));
    c->a_slots_cap = slots;
  }
  if (c->d_astates_cap < n) {
    if (c->d_astates) cudaFree(c->d_astates);
    c->d_astates = nullptr;
    c->d_astates_cap = 0;
    const uint64_t cap = std::max<uint64_t>(2 * n, 256);
    CU(cudaMalloc(&c->d_astates, cap * sizeof(AliasState)));
    c->d_astates_cap = cap;
  }
  std::vector<unsigned long long> hk(slots, ~0ull);
  std::vector<uint32_t> ho(slots, 0u);
  const uint64_t mask = slots - 1;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t x = c->akeys[i];  // mix64 of kernels.cu
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    uint64_t s = x & mask;
    while (hk[s] != ~0ull) s = (s + 1) & mask;
    hk[s] = c->akeys[i];
    ho[s] = (uint32_t)i;
  }
  CU(cudaMemcpy(c->d_akeys, hk.data(), slots * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->d_aord, ho.data(), slots * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->d_astates, c->astates.data(), n * sizeof(AliasState), cudaMemcpyHostToDevice));
  c->a_slots = slots;
  return PCQ_OK;
}

int grow_cands(pcq_collector* c, uint64_t need) {
  if (c->grid.cand_cap >= need) return PCQ_OK;
  uint64_t cap = std::max<uint64_t>(need, 1u << 16);
  Candidate* nb = nullptr;
  CU(cudaMalloc(&nb, cap * sizeof(Candidate)));
  if (c->grid.cands) {
    if (c->cand_len)
      CU(cudaMemcpyAsync(nb, c->grid.cands, c->cand_len * sizeof(Candidate), cudaMemcpyDeviceToDevice, c->ctx->stream));
    CU(cudaStreamSynchronize(c->ctx->stream));
    cudaFree(c->grid.cands);
  }
  c->grid.cands = nb;
  c->grid.cand_cap = cap;
  return PCQ_OK;
}

// drop candidates that no longer hold their cell's minimum; compacts in place through a scratch arena
int prune_cands(pcq_collector* c) {
  pcq_ctx* ctx = c->ctx;
  RC(grid_restore(c));
  c->prune_epoch++;
  if (c->cand_len == 0) return PCQ_OK;
  Candidate* tmp = nullptr;
  CU(cudaMalloc(&tmp, std::max<uint64_t>(c->cand_len, 1) * sizeof(Candidate)));
  CU(cudaMemsetAsync(&c->dev->out_count, 0, sizeof(unsigned long long), ctx->stream));
  GridDev g = grid_view(c);
  if (launch_grid_prune(g, c->cand_len, tmp, &c->dev->out_count, ctx->sm_count, ctx->stream) != 0) {
    cudaFree(tmp);
    return fail(PCQ_ERR_CUDA, "k_grid_prune launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  ctx->launches++;
  DevBlock b;
  int rc = read_devblock(c, &b);
  if (rc != PCQ_OK) {
    cudaFree(tmp);
    return rc;
  }
  const uint64_t kept = b.out_count;
  if (kept) cudaMemcpyAsync(c->grid.cands, tmp, kept * sizeof(Candidate), cudaMemcpyDeviceToDevice, ctx->stream);
  cudaMemcpyAsync(&c->dev->cand_count, &kept, sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  c->cand_len = kept;
  return PCQ_OK;
}

int alloc_grid_tables(pcq_collector* c, uint64_t slots, bool hashed) {
  pcq_ctx* ctx = c->ctx;
  unsigned long long* table = nullptr;
  unsigned long long* hkeys = nullptr;
  if (cudaMalloc(&table, slots * 8ull) != cudaSuccess) {
    cudaGetLastError();
    return fail(PCQ_ERR_NOMEM, "cannot allocate density table of %llu cells", (unsigned long long)slots);
  }
  if (hashed && cudaMalloc(&hkeys, slots * 8ull) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(table);
    return fail(PCQ_ERR_NOMEM, "cannot allocate density key table of %llu slots", (unsigned long long)slots);
  }
  CU(cudaMemsetAsync(table, 0xFF, slots * 8ull, ctx->stream));
  if (hkeys) CU(cudaMemsetAsync(hkeys, 0xFF, slots * 8ull, ctx->stream));
  c->grid.table = table;
  c->grid.hkeys = hkeys;
  c->grid.table_slots = slots;
  return PCQ_OK;
}

// Cell tables of the grid collectors of one call that do not have one yet.  `boxes` (6 doubles per collector: min xyz,
// max xyz; may be null) bounds the positions the collector can be fed in this call — the header boxes of its files,
// cut by the query box.  In order of preference:
//   dense over the sub-box of cells under `boxes` — when that is at most half of the grid and fits the budget;
//   dense over the whole grid (slot == key) — when the key has at most 30 bits and it fits the budget;
//   an open-addressing hash that starts small and is quadrupled on demand.
// The budget is a third of the free HBM shared by the collectors of the call that need a table: `query --parallel
// --density` makes one grid per file, and 64 dense ca13-XL tables would be 512 GB.  Results do not depend on the choice.
static bool cell_range(const GridDev& g, int a, double lo, double hi, uint64_t* c_lo, uint64_t* c_n) {
  if (!(lo <= hi) || !std::isfinite(lo) || !std::isfinite(hi) || !(g.ext[a] > 0.0) || !std::isfinite(g.ext[a])) return false;
  auto cell = [&](double p) -> double {
    const double r = ((p - g.bmin[a]) * g.dims_f[a]) / g.ext[a];  // :51-57 (any rounding difference is inside the margin)
    return r > 0.0 ? std::floor(r) : 0.0;
  };
  // (positions outside the grid's own box still have cells — 0 below it, up to the mask above it, beyond the mask they
  // are aliased and bypass the table: the range is NOT cut to the grid box)
  const double cmax = (double)g.mask[a];
  double l = cell(lo) - 1.0, h = cell(hi) + 1.0;
  l = std::min(std::max(l, 0.0), cmax);
  h = std::min(std::max(h, l), cmax);
  *c_lo = (uint64_t)l;
  *c_n = (uint64_t)(h - l) + 1ull;
  return true;
}

int ensure_grid_tables(pcq_collector* const* cols, uint32_t n, const double* boxes) {
  // an EMPTY collector (fresh, or reset) whose sub-box table does not cover the box of this call gets a new table
  for (uint32_t i = 0; i < n && boxes; ++i) {
    pcq_collector* c = cols[i];
    GridDev& g = c->grid;
    if (c->kind != PCQ_COLLECT_GRID || !g.table || g.hkeys || !g.sub_on || c->cand_len != 0 || !c->akeys.empty()) continue;
    const double* bx = boxes + 6 * (size_t)i;
    bool inside = true;
    for (int a = 0; a < 3 && inside; ++a) {
      uint64_t lo = 0, nn = 0;
      inside = cell_range(g, a, bx[a], bx[3 + a], &lo, &nn) && lo >= g.sub_lo[a] && lo + nn <= g.sub_lo[a] + g.sub_n[a];
    }
    if (inside) continue;
    CU(cudaStreamSynchronize(c->ctx->stream));
    cudaFree(g.table);
    g.table = nullptr;
    g.sub_on = 0;
    c->table_holds_winners = false;
  }
  uint32_t missing = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (cols[i]->kind == PCQ_COLLECT_GRID && !cols[i]->grid.table) ++missing;
  if (!missing) return PCQ_OK;
  uint64_t dense_bits = 30;
  if (const char* e = std::getenv("PCQ_DENSE_MAX_BITS")) dense_bits = (uint64_t)std::atoi(e);
  bool use_sub = true;
  if (const char* e = std::getenv("PCQ_GRID_SUBBOX")) use_sub = std::atoi(e) != 0;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const uint64_t budget = (uint64_t)free_b / 3 / missing;
  for (uint32_t i = 0; i < n; ++i) {
    pcq_collector* c = cols[i];
    if (c->kind != PCQ_COLLECT_GRID || c->grid.table) continue;
    GridDev& g = c->grid;
    const bool dense_full = c->total_bits <= dense_bits && (8ull << c->total_bits) <= budget;
    g.sub_on = 0;
    if (use_sub && boxes) {
      const double* bx = boxes + 6 * (size_t)i;
      uint64_t lo[3], nn[3];
      bool ok = true;
      for (int a = 0; a < 3 && ok; ++a) ok = cell_range(g, a, bx[a], bx[3 + a], &lo[a], &nn[a]);
      if (ok) {
        const long double cells = (long double)nn[0] * (long double)nn[1] * (long double)nn[2];
        const long double full = std::ldexp(1.0L, (int)c->total_bits);
        if (cells * 8.0L <= (long double)budget && (!dense_full || cells * 2.0L <= full)) {
          for (int a = 0; a < 3; ++a) {
            g.sub_lo[a] = lo[a];
            g.sub_n[a] = nn[a];
          }
          g.sub_on = 1;
          RC(alloc_grid_tables(c, nn[0] * nn[1] * nn[2], false));
          continue;
        }
      }
    }
    uint64_t slots = dense_full ? (1ull << c->total_bits) : (1ull << 22);
    if (const char* e = std::getenv("PCQ_HASH_SLOTS_LOG2"))
      if (!dense_full) slots = 1ull << std::atoi(e);
    RC(alloc_grid_tables(c, slots, !dense_full));
  }
  return PCQ_OK;
}

// hashed table full: quadruple it and re-insert the surviving candidates (a sub-box table that met a cell outside its
// box: replace it by a table over the whole grid)
int rehash_grid(pcq_collector* c) {
  pcq_ctx* ctx = c->ctx;
  const bool was_sub = c->grid.table && c->grid.hkeys == nullptr && c->grid.sub_on;
  // (a sub-box table only moves: every candidate of the earlier launches is folded into the new table, nothing is
  // pruned, so an ordered replay of aliased keys met in the same launch still finds what it needs)
  if (was_sub) RC(grid_restore(c));
  else RC(prune_cands(c));
  const uint64_t kept = c->cand_len;
  Candidate* tmp = nullptr;
  if (kept) {
    CU(cudaMalloc(&tmp, kept * sizeof(Candidate)));
    CU(cudaMemcpyAsync(tmp, c->grid.cands, kept * sizeof(Candidate), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  const uint64_t old_slots = c->grid.table_slots;
  cudaFree(c->grid.table);
  cudaFree(c->grid.hkeys);
  c->grid.table = c->grid.hkeys = nullptr;
  int rc;
  if (was_sub) {
    // a cell outside the sub-box the table was made for: from now on a table over the whole grid
    c->grid.sub_on = 0;
    pcq_collector* one = c;
    rc = ensure_grid_tables(&one, 1, nullptr);
  } else {
    rc = alloc_grid_tables(c, old_slots * 4ull, true);
  }
  if (rc != PCQ_OK) {
    cudaFree(tmp);
    return rc;
  }
  const unsigned long long zero = 0;
  cudaMemcpyAsync(&c->dev->cand_count, &zero, sizeof(zero), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemsetAsync(&c->dev->flags, 0, sizeof(uint32_t), ctx->stream);
  c->cand_len = 0;
  if (kept) {
    GridDev g = grid_view(c);
    if (launch_grid_import(g, tmp, kept, ctx->sm_count, ctx->stream) != 0) {
      cudaFree(tmp);
      return fail(PCQ_ERR_CUDA, "k_grid_import launch failed");
    }
    ctx->launches++;
  }
  DevBlock b;
  rc = read_devblock(c, &b);
  cudaFree(tmp);
  if (rc != PCQ_OK) return rc;
  c->cand_len = std::min<uint64_t>(b.cand_count, c->grid.cand_cap);
  return PCQ_OK;
}

int grid_finalize(pcq_collector* c) {
  if (c->final_valid) return PCQ_OK;
  pcq_ctx* ctx = c->ctx;
  RC(grid_restore(c));
  const uint64_t n = c->cand_len;
  // winners of affected keys come from the ordered replay, not from the table
  std::vector<uint8_t> replayed;
  for (const AliasState& s : c->astates)
    if (s.valid) replayed.insert(replayed.end(), s.point, s.point + 31);
  const uint64_t n_replayed = replayed.size() / 31;
  c->final_n = 0;
  if (n == 0 && n_replayed == 0) {
    c->final_valid = true;
    return PCQ_OK;
  }
  if (c->final_cap < n + n_replayed) {
    // (with slack: the number of candidates of the same query varies from run to run with the order in which the
    // GPU happens to meet the points, and a cudaFree + cudaMalloc of a few hundred megabytes costs 20-40 ms)
    if (c->d_final) cudaFree(c->d_final);
    c->d_final = nullptr;
    c->final_cap = 0;
    const uint64_t cap = n + n_replayed + (n + n_replayed) / 4 + 4096;
    if (cudaMalloc(&c->d_final, cap * 31ull + 64) != cudaSuccess) {
      cudaGetLastError();
      return fail(PCQ_ERR_NOMEM, "cannot allocate %llu bytes of HBM for the density result", (unsigned long long)(cap * 31ull));
    }
    c->final_cap = cap;
  }
  uint64_t from_table = 0;
  if (n) {
    CU(cudaMemsetAsync(&c->dev->out_count, 0, sizeof(unsigned long long), ctx->stream));
    GridDev g = grid_view(c);
    RC(grid_pick_winners(c, g, n));
    if (launch_grid_emit(g, n, c->d_finlist, &c->dev->fin_count, 2, 1, nullptr, nullptr, nullptr, c->d_final, &c->dev->out_count,
                         ctx->sm_count, ctx->stream) != 0)
      return fail(PCQ_ERR_CUDA, "k_grid_emit launch failed");
    ctx->launches += 1;
    DevBlock b;
    RC(read_devblock(c, &b));
    from_table = b.out_count;
  }
  if (n_replayed) {
    CU(cudaMemcpyAsync(c->d_final + from_table * 31ull, replayed.data(), replayed.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  c->final_n = from_table + n_replayed;
  c->final_valid = true;
  return PCQ_OK;
}

// A GRID launch logged points of affected keys (key aliasing, see alias.cu).  New affected keys get their fold state
// from the earlier launches, a second (log-only) pass collects every point of theirs in this launch, and the log is
// replayed in scan order.
int alias_slow_path(pcq_ctx* ctx, const std::vector<Segment>& segs, ScanParams P, int variant, uint32_t R, int min_align,
                    pcq_collector* const* collectors, uint32_t n_collectors, std::vector<DevBlock>& blocks,
                    const std::vector<uint64_t>& epoch0, const std::vector<uint64_t>& lane_lo) {
  bool pass2 = false;
  for (uint32_t l = 0; l < n_collectors; ++l) {
    pcq_collector* c = collectors[l];
    const uint64_t n_log = blocks[l].log_count;
    if (n_log == 0) continue;
    std::vector<Candidate> h(n_log);
    CU(cudaMemcpy(h.data(), c->d_log, n_log * sizeof(Candidate), cudaMemcpyDeviceToHost));
    std::unordered_set<uint64_t> known(c->akeys.begin(), c->akeys.end());
    std::vector<uint64_t> fresh;
    for (const Candidate& e : h)
      if (known.insert(e.key).second) fresh.push_back(e.key);
    if (fresh.empty()) continue;
    if (lane_lo[l] < c->scan_hi)
      return fail(PCQ_ERR_ALIASED,
                  "density grid: key aliasing (grid_sampling.rs:62-70 vs 78-82) makes the result depend on insertion order; "
                  "the ordered replay needs point ranges fed in scan order, but this launch starts at scan index %llu "
                  "after %llu was already fed",
                  (unsigned long long)lane_lo[l], (unsigned long long)c->scan_hi);
    if (c->prune_epoch != epoch0[l])
      return fail(PCQ_ERR_ALIASED,
                  "density grid: key aliasing met in a launch that also had to drop candidates (table rehash / arena "
                  "overflow); the earlier winner of the aliased key is no longer available for the ordered replay");
    const uint32_t ord0 = (uint32_t)c->akeys.size();
    c->akeys.insert(c->akeys.end(), fresh.begin(), fresh.end());
    c->astates.resize(c->akeys.size());
    for (size_t k = ord0; k < c->astates.size(); ++k) std::memset(&c->astates[k], 0, sizeof(AliasState));
    RC(alias_upload(c));
    GridDev g = grid_view(c);
    if (alias_prewinners(g, c->cand_len, nullptr, (uint32_t)fresh.size(), lane_lo[l], c->d_astates, ord0, ctx->sm_count,
                         ctx->stream) != 0)
      return fail(PCQ_ERR_CUDA, "alias_prewinners failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->launches += 3;
    CU(cudaMemcpy(c->astates.data() + ord0, c->d_astates + ord0, fresh.size() * sizeof(AliasState), cudaMemcpyDeviceToHost));
    pass2 = true;
  }
  if (pass2) {
    // the first pass could not know the new keys: collect every point of an affected key again, insert nothing
    bool done = false;
    for (int attempt = 0; attempt < 8 && !done; ++attempt) {
      std::vector<LaneDev> lanes(n_collectors);
      for (uint32_t l = 0; l < n_collectors; ++l) {
        pcq_collector* c = collectors[l];
        CU(cudaMemsetAsync(&c->dev->log_count, 0, sizeof(unsigned long long), ctx->stream));
        CU(cudaMemsetAsync(&c->dev->flags, 0, sizeof(uint32_t), ctx->stream));
        std::memset(&lanes[l], 0, sizeof(LaneDev));
        lanes[l].count = &c->dev->count;
        lanes[l].grid = grid_view(c);
        lanes[l].grid.log_only = 1;
      }
      void* d_segs = nullptr;
      void* d_lanes = nullptr;
      RC(upload2(ctx, segs.data(), segs.size() * sizeof(Segment), lanes.data(), lanes.size() * sizeof(LaneDev), &d_segs, &d_lanes));
      P.segs = static_cast<const Segment*>(d_segs);
      P.lanes = static_cast<const LaneDev*>(d_lanes);
      if (P.one_grid) P.grid0 = lanes[0].grid;
      if (launch_scan(variant, MODE_GRID, P, R, min_align, ctx->sm_count, ctx->stream) != 0)
        return fail(PCQ_ERR_CUDA, "scan kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      ctx->launches++;
      for (uint32_t l = 0; l < n_collectors; ++l)
        CU(cudaMemcpyAsync(&blocks[l], collectors[l]->dev, sizeof(DevBlock), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      done = true;
      for (uint32_t l = 0; l < n_collectors; ++l) {
        if (blocks[l].flags & kFlagLogOverflow) {
          RC(grow_log(collectors[l], blocks[l].log_count + blocks[l].log_count / 4 + 1024));
          done = false;
        }
      }
    }
    if (!done) return fail(PCQ_ERR_NOMEM, "replay log capacity did not converge");
  }
  for (uint32_t l = 0; l < n_collectors; ++l) {
    pcq_collector* c = collectors[l];
    const uint64_t n_log = blocks[l].log_count;
    if (n_log == 0) continue;
    GridDev g = grid_view(c);
    const int rc = alias_replay(g, n_log, c->d_astates, ctx->sm_count, ctx->stream);
    if (rc != 0) return fail(rc == -2 ? PCQ_ERR_NOMEM : PCQ_ERR_CUDA, "alias replay of %llu points failed", (unsigned long long)n_log);
    ctx->launches += 5;
    CU(cudaMemcpy(c->astates.data(), c->d_astates, c->astates.size() * sizeof(AliasState), cudaMemcpyDeviceToHost));
    c->final_valid = false;
  }
  return PCQ_OK;
}

// one kernel launch for a prepared batch of segments; handles BUFFER / GRID capacity retries
// Share of a file's points a bounds query can be expected to match if points were spread evenly over the header box:
// volume of (query box ∩ header box) / volume of the header box.  Only steers kernel choice (density insert: match
// queue or direct), never results.  Class queries: unknown (-1).
double expected_match_fraction(const pcq_file_desc& d, const pcq_query* q) {
  if (q->kind != PCQ_QUERY_BOUNDS) return -1.0;
  double f = 1.0;
  for (int i = 0; i < 3; ++i) {
    const double ext = d.hdr_max[i] - d.hdr_min[i];
    if (!(ext > 0.0)) continue;  // flat or unusable header axis: no information
    const double lo = std::max(q->qmin[i], d.hdr_min[i]), hi = std::min(q->qmax[i], d.hdr_max[i]);
    f *= std::min(1.0, std::max(0.0, (hi - lo) / ext));
  }
  return f;
}

// k_class_count_soa walks the segments one after the other with the whole grid: fine for files, not for the thousands
// of chunk runs an indexed launch can hold (those take the tile-scheduled scan)
constexpr size_t kSoaCountMaxSegs = 256;

// Where the positions a collector is fed in one call can lie: the union of the header boxes of its files, cut by the
// query box of a bounds query (6 doubles per collector: min xyz, max xyz; a collector without a file keeps an empty box).
struct LaneBoxes {
  std::vector<double> v;
  explicit LaneBoxes(uint32_t n) : v(6 * (size_t)n) {
    for (uint32_t l = 0; l < n; ++l)
      for (int a = 0; a < 3; ++a) {
        v[6 * (size_t)l + a] = INFINITY;
        v[6 * (size_t)l + 3 + a] = -INFINITY;
      }
  }
  void add(uint32_t lane, const pcq_file_desc& d) {
    for (int a = 0; a < 3; ++a) {
      v[6 * (size_t)lane + a] = std::min(v[6 * (size_t)lane + a], d.hdr_min[a]);
      v[6 * (size_t)lane + 3 + a] = std::max(v[6 * (size_t)lane + 3 + a], d.hdr_max[a]);
    }
  }
  void cut(const pcq_query* q) {
    if (q->kind != PCQ_QUERY_BOUNDS) return;
    for (size_t l = 0; l < v.size() / 6; ++l)
      for (int a = 0; a < 3; ++a) {
        v[6 * l + a] = std::max(v[6 * l + a], q->qmin[a]);
        v[6 * l + 3 + a] = std::min(v[6 * l + 3 + a], q->qmax[a]);
      }
  }
};

// lane_box (optional, 6 doubles per collector): where the positions fed to the collector in this call can lie
int run_batch(pcq_ctx* ctx, std::vector<Segment>& segs, const pcq_query* q, pcq_collector* const* collectors,
              uint32_t n_collectors, const std::vector<uint64_t>& lane_points, double match_fraction = -1.0,
              const std::vector<double>* lane_box = nullptr) {
  const int kind = collectors[0]->kind;
  if (segs.empty()) return PCQ_OK;

  // kernel variant: staged needs one record length, 16-byte aligned ranges, and records that carry
  // the predicate's field (LAST class queries read the class column -> direct)
  uint32_t R = segs[0].record_len;
  bool staged_ok = staged_supports(R);
  bool all_last = true;
  int min_align = 4;
  for (const Segment& s : segs) {
    min_align = std::min<int>(min_align, s.align);
    if (s.record_len != R) staged_ok = false;
    if ((reinterpret_cast<uintptr_t>(s.rec) & 15u) != 0) staged_ok = false;
    if (s.layout == PCQ_LAYOUT_LAST && q->kind == PCQ_QUERY_CLASS) staged_ok = false;
    if (s.layout != PCQ_LAYOUT_LAST) all_last = false;
  }
  int variant = ctx->variant;
  if (variant == 0) variant = staged_ok ? 2 : 1;
  if (variant == 2 && !staged_ok) variant = 1;

  std::vector<LaneDev> lanes(n_collectors);
  const int mode = kind == PCQ_COLLECT_COUNT ? MODE_COUNT : (kind == PCQ_COLLECT_BUFFER ? MODE_SELECT : MODE_GRID);

  // number the scheduling units ("tiles") across the segments of the launch; lanes are contiguous runs
  bool select_bytes = mode == MODE_SELECT && q->kind == PCQ_QUERY_CLASS && all_last;
  for (const Segment& s : segs)
    if ((reinterpret_cast<uintptr_t>(s.cls) & 15u) != 0) select_bytes = false;
  const uint32_t tile_pts = tile_points(variant, mode, R, select_bytes);
  uint64_t n_tiles = 0;
  {
    std::vector<uint64_t> lane_first(n_collectors, ~0ull);
    for (Segment& s : segs) {
      s.first_tile = n_tiles;
      if (lane_first[s.lane] == ~0ull) lane_first[s.lane] = n_tiles;
      s.lane_first_tile = lane_first[s.lane];
      n_tiles += (s.n_points + tile_pts - 1) / tile_pts;
    }
  }
  if (n_tiles == 0) return PCQ_OK;

  // first-attempt capacities are guesses (an overflowing launch is re-run with the exact size); on multi-billion-point
  // inputs they must not eat the HBM the exact size may need later: at most a sixth of what is free, shared by the lanes.
  // cudaMemGetInfo is only asked when some collector really has to grow: its cost is anything from microseconds to
  // milliseconds, and a collector that is reused (reset, next query) already owns what it needs.
  uint64_t budget = 0;
  auto growth_budget = [&]() -> uint64_t {
    if (budget == 0) {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      budget = std::max<uint64_t>((uint64_t)free_b / 6 / std::max<uint32_t>(n_collectors, 1u), 64u << 20);
    }
    return budget;
  };
  if (kind == PCQ_COLLECT_BUFFER) {
    for (uint32_t l = 0; l < n_collectors; ++l) {
      pcq_collector* c = collectors[l];
      if (lane_points[l] == 0) continue;
      uint64_t guess = std::min<uint64_t>(lane_points[l], (1u << 20) + lane_points[l] / 8);
      if (c->out_cap >= c->out_len + guess) continue;
      guess = std::min<uint64_t>(guess, std::max<uint64_t>(growth_budget() / 31, 1u << 20));
      RC(grow_out(c, c->out_len + guess));
    }
  } else if (kind == PCQ_COLLECT_GRID) {
    {  // (collectors whose files the query does not touch need no table)
      std::vector<pcq_collector*> active;
      std::vector<double> boxes;
      for (uint32_t l = 0; l < n_collectors; ++l) {
        if (lane_points[l] == 0) continue;
        active.push_back(collectors[l]);
        if (lane_box) boxes.insert(boxes.end(), lane_box->begin() + 6 * (size_t)l, lane_box->begin() + 6 * (size_t)l + 6);
      }
      RC(ensure_grid_tables(active.data(), (uint32_t)active.size(), lane_box ? boxes.data() : nullptr));
    }
    for (uint32_t l = 0; l < n_collectors; ++l) {
      pcq_collector* c = collectors[l];
      if (lane_points[l] == 0) continue;
      RC(grid_restore(c));
      const uint64_t slack = (uint64_t)ctx->sm_count * kGridCtasPerSm * (kBlock / 32) * kCandChunk;
      uint64_t guess = std::min<uint64_t>(lane_points[l], (1u << 22) + lane_points[l] / 16);
      if (c->cand_len + guess + slack > c->grid.cand_cap)
        guess = std::min<uint64_t>(guess, std::max<uint64_t>(growth_budget() / sizeof(Candidate), 1u << 22));
      guess += slack;
      // Candidates are only ever dropped BETWEEN launches: what survives is every cell's current winner, which is
      // exactly what a later launch needs if one of its points makes the cell's key an aliased one (alias.cu).
      if (c->cand_len + guess > c->grid.cand_cap && c->cand_len > (1u << 20)) RC(prune_cands(c));
      RC(grow_cands(c, c->cand_len + guess));
      RC(grow_log(c, 1u << 14));
      c->final_valid = false;
    }
  }
  // GRID: scan-index range of every lane in this launch and the prune epochs (ordered alias replay, alias.cu)
  std::vector<uint64_t> lane_lo(n_collectors, ~0ull), lane_hi(n_collectors, 0), epoch0(n_collectors, 0);
  if (kind == PCQ_COLLECT_GRID) {
    for (const Segment& s : segs) {
      lane_lo[s.lane] = std::min<uint64_t>(lane_lo[s.lane], s.scan_base);
      lane_hi[s.lane] = std::max<uint64_t>(lane_hi[s.lane], s.scan_base + s.n_points);
    }
    for (uint32_t l = 0; l < n_collectors; ++l) epoch0[l] = collectors[l]->prune_epoch;
  }

  for (int attempt = 0; attempt < 8; ++attempt) {
    for (uint32_t l = 0; l < n_collectors; ++l) {
      pcq_collector* c = collectors[l];
      LaneDev& L = lanes[l];
      std::memset(&L, 0, sizeof(L));
      L.count = &c->dev->count;
      L.out = c->d_out;
      L.out_base = c->out_len;
      L.out_cap = c->out_cap;
      if (kind == PCQ_COLLECT_GRID) L.grid = grid_view(c);
    }
    void* d_segs = nullptr;
    void* d_lanes = nullptr;
    RC(upload2(ctx, segs.data(), segs.size() * sizeof(Segment), lanes.data(), lanes.size() * sizeof(LaneDev), &d_segs, &d_lanes));

    ScanParams P{};
    P.segs = static_cast<const Segment*>(d_segs);
    P.n_segs = (uint32_t)segs.size();
    P.query_kind = q->kind;
    P.cls = q->cls;
    P.n_tiles = n_tiles;
    P.tile_pts = tile_pts;
    P.sel_bytes = select_bytes ? 1u : 0u;
    // density insert: a box that covers less than a quarter of the points' volume goes through the match queue
    P.grid_sparse = (mode == MODE_GRID && match_fraction >= 0.0 && match_fraction < 0.25) ? 1u : 0u;
    if (const char* e = std::getenv("PCQ_GRID_SPARSE")) P.grid_sparse = std::atoi(e) ? 1u : 0u;
    // uniform record length + 16-byte aligned ranges (staged_ok) and a predicate that lives in the staged records
    P.sel_ring = (mode == MODE_SELECT && variant == 2 && !select_bytes) ? R : 0u;
    P.lanes = static_cast<const LaneDev*>(d_lanes);
    if (mode == MODE_GRID && n_collectors == 1) {
      P.one_grid = 1u;
      P.grid0 = lanes[0].grid;
    }
    if (mode == MODE_SELECT) {
#ifdef PCQ_DEBUG_HOOKS
      if (const char* e = std::getenv("PCQ_SELECT_DEBUG")) P.debug = (uint32_t)std::atoi(e);
#endif
      RC(ensure_tile_state(ctx, n_tiles));
      P.ticket = ctx->tile_state;
      P.tile_state = ctx->tile_state + kDescStride;
    }

    int lrc;
    if (mode == MODE_COUNT && q->kind == PCQ_QUERY_CLASS && all_last && segs.size() <= kSoaCountMaxSegs) {
      lrc = launch_class_count_soa(P, ctx->sm_count, ctx->stream);
    } else {
      lrc = launch_scan(variant, mode, P, R, min_align, ctx->sm_count, ctx->stream);
    }
    if (lrc != 0) return fail(PCQ_ERR_CUDA, "scan kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->launches++;

    if (kind == PCQ_COLLECT_COUNT) return PCQ_OK;  // fully asynchronous

    bool retry = false;
    if (kind == PCQ_COLLECT_BUFFER) {
      std::vector<DevBlock> blocks;
      RC(read_devblocks(ctx, collectors, n_collectors, static_cast<const LaneDev*>(d_lanes), blocks));
      for (uint32_t l = 0; l < n_collectors; ++l)
        if (blocks[l].count > collectors[l]->out_cap) retry = true;
      if (!retry) {
        for (uint32_t l = 0; l < n_collectors; ++l) collectors[l]->out_len = blocks[l].count;
        return PCQ_OK;
      }
      // some lane overflowed its buffer: grow to the exact size, rewind every counter, run again
      for (uint32_t l = 0; l < n_collectors; ++l) {
        pcq_collector* c = collectors[l];
        RC(grow_out(c, blocks[l].count));
        const unsigned long long len = c->out_len;
        CU(cudaMemcpyAsync(&c->dev->count, &len, sizeof(len), cudaMemcpyHostToDevice, ctx->stream));
      }
      CU(cudaStreamSynchronize(ctx->stream));
      continue;
    }

    // GRID
    std::vector<DevBlock> blocks;
    RC(read_devblocks(ctx, collectors, n_collectors, static_cast<const LaneDev*>(d_lanes), blocks));
    for (uint32_t l = 0; l < n_collectors; ++l) {
      pcq_collector* c = collectors[l];
      const DevBlock& b = blocks[l];
      if (b.flags & kFlagHashFull) {
        RC(rehash_grid(c));
        retry = true;
        continue;
      }
      if (b.flags & kFlagCandOverflow) {
        // Rewind to the candidates of the earlier launches and run the launch again into a larger arena.  (The table
        // keeps what the failed attempt wrote: a cell's true winner still qualifies, non-winners no longer do.)
        const uint64_t attempted = b.cand_count;  // what this launch wanted in total
        RC(grow_cands(c, attempted + attempted / 4 + (uint64_t)ctx->sm_count * kGridCtasPerSm * (kBlock / 32) * kCandChunk));
        const unsigned long long keep = c->cand_len;
        CU(cudaMemcpyAsync(&c->dev->cand_count, &keep, sizeof(keep), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemsetAsync(&c->dev->flags, 0, sizeof(uint32_t), ctx->stream));
        retry = true;
        continue;
      }
      if (b.flags & kFlagLogOverflow) {
        RC(grow_log(c, b.log_count + b.log_count / 4 + 1024));
        CU(cudaMemsetAsync(&c->dev->flags, 0, sizeof(uint32_t), ctx->stream));
        retry = true;
        continue;
      }
      c->cand_len = b.cand_count;
    }
    if (retry) {
      for (uint32_t l = 0; l < n_collectors; ++l)
        CU(cudaMemsetAsync(&collectors[l]->dev->log_count, 0, sizeof(unsigned long long), ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      continue;
    }
    bool any_logged = false;
    // group-wide ordered replay (group.cu): a log-only pass only collects the points of the affected keys; the
    // launch's log is appended to the collector's raw log, nothing is inserted or folded here
    any_logged = false;
    for (uint32_t l = 0; l < n_collectors; ++l) {
      pcq_collector* c = collectors[l];
      const uint64_t n_log = blocks[l].log_count;
      if (c->pass_mode != 1) {
        any_logged |= n_log != 0;
        continue;
      }
      if (n_log == 0) continue;
      if (c->rawlog_cap < c->rawlog_len + n_log) {
        const uint64_t cap = std::max<uint64_t>((c->rawlog_len + n_log) * 2, 1u << 14);
        Candidate* nb = nullptr;
        if (cudaMalloc(&nb, cap * sizeof(Candidate)) != cudaSuccess) {
          cudaGetLastError();
          return fail(PCQ_ERR_NOMEM, "cannot allocate a raw replay log of %llu entries", (unsigned long long)cap);
        }
        if (c->rawlog_len) CU(cudaMemcpyAsync(nb, c->d_rawlog, c->rawlog_len * sizeof(Candidate), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (c->d_rawlog) cudaFree(c->d_rawlog);
        c->d_rawlog = nb;
        c->rawlog_cap = cap;
      }
      CU(cudaMemcpyAsync(c->d_rawlog + c->rawlog_len, c->d_log, n_log * sizeof(Candidate), cudaMemcpyDeviceToDevice, ctx->stream));
      c->rawlog_len += n_log;
      blocks[l].log_count = 0;
    }
    if (any_logged) RC(alias_slow_path(ctx, segs, P, variant, R, min_align, collectors, n_collectors, blocks, epoch0, lane_lo));
    for (uint32_t l = 0; l < n_collectors; ++l) {
      pcq_collector* c = collectors[l];
      if (lane_hi[l] > c->scan_hi) c->scan_hi = lane_hi[l];
      CU(cudaMemsetAsync(&c->dev->log_count, 0, sizeof(unsigned long long), ctx->stream));
    }
    return PCQ_OK;
  }
  return fail(PCQ_ERR_NOMEM, "collector capacity did not converge");
}

}  // namespace pcq

// =================================================================================================
extern "C" {

int pcq_ctx_create(int device, pcq_ctx** out) {
  if (!out) return fail(PCQ_ERR_ARG, "pcq_ctx_create: null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(PCQ_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= n) return fail(PCQ_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(PCQ_ERR_CUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
  pcq_ctx* ctx = new (std::nothrow) pcq_ctx();
  if (!ctx) return fail(PCQ_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return fail(PCQ_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(cudaGetLastError()));
  }
  ctx->own_stream = true;
  const char* v = std::getenv("PCQ_SCAN_VARIANT");
  if (v) ctx->variant = std::atoi(v);
  *out = ctx;
  return PCQ_OK;
}

static void ctx_free(pcq_ctx* ctx);

void pcq_ctx_destroy(pcq_ctx* ctx) {
  if (!ctx) return;
  if (ctx->refs > 0) {
    ctx->closing = true;  // freed when the last file / collector goes away
    return;
  }
  ctx_free(ctx);
}

static void ctx_unref(pcq_ctx* ctx) {
  if (--ctx->refs == 0 && ctx->closing) ctx_free(ctx);
}

static void ctx_free(pcq_ctx* ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (UploadSlot& s : ctx->slots) {
    if (s.host) cudaFreeHost(s.host);
    if (s.dev) cudaFree(s.dev);
    if (s.ev) cudaEventDestroy(s.ev);
  }
  if (ctx->tile_state) cudaFree(ctx->tile_state);
  if (ctx->part_scratch) cudaFree(ctx->part_scratch);
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
  for (int i = 0; i < kChunkBuffers; ++i) {
    if (ctx->chunk[i]) cudaFree(ctx->chunk[i]);
    if (ctx->chunk_copied[i]) cudaEventDestroy(ctx->chunk_copied[i]);
    if (ctx->chunk_free[i]) cudaEventDestroy(ctx->chunk_free[i]);
    if (ctx->bounce[i]) cudaFreeHost(ctx->bounce[i]);
    if (ctx->bounce_done[i]) cudaEventDestroy(ctx->bounce_done[i]);
  }
  if (ctx->index_scratch) cudaFree(ctx->index_scratch);
  if (ctx->index_bounce) cudaFreeHost(ctx->index_bounce);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int pcq_ctx_set_stream(pcq_ctx* ctx, void* cuda_stream) {
  if (!ctx) return fail(PCQ_ERR_ARG, "null ctx");
  RC(use_device(ctx));
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return PCQ_OK;
}

int pcq_ctx_synchronize(pcq_ctx* ctx) {
  if (!ctx) return fail(PCQ_ERR_ARG, "null ctx");
  RC(use_device(ctx));
  CU(cudaStreamSynchronize(ctx->stream));
  return PCQ_OK;
}

int pcq_ctx_set_scan_variant(pcq_ctx* ctx, int variant) {
  if (!ctx || variant < 0 || variant > 2) return fail(PCQ_ERR_ARG, "bad scan variant");
  ctx->variant = variant;
  return PCQ_OK;
}

uint64_t pcq_ctx_launch_count(const pcq_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- files ---------------------------------------------------------------------------------------

int pcq_file_stage_host(pcq_ctx* ctx, const void* file_bytes, size_t n_bytes, const char* ext, uint64_t first_point,
                        uint64_t n_points, pcq_file** out) {
  if (!ctx || !file_bytes || !out) return fail(PCQ_ERR_ARG, "pcq_file_stage_host: null argument");
  const int layout = layout_of_ext(ext);
  if (layout < 0) return fail(PCQ_ERR_FORMAT, "Unsupported file extension \"%s\" (this path serves las and last)", ext ? ext : "");
  RC(use_device(ctx));
  pcq_file_desc d;
  uint8_t raw = 0;
  RC(parse_header(file_bytes, n_bytes, layout, 1, &d, &raw));
  if (first_point > d.n_points) return fail(PCQ_ERR_ARG, "first_point %llu beyond %llu points", (unsigned long long)first_point, (unsigned long long)d.n_points);
  uint64_t n = std::min<uint64_t>(n_points, d.n_points - first_point);
  const uint64_t N = d.n_points;
  if ((uint64_t)d.point_data_off > (uint64_t)n_bytes || d.record_len == 0 ||
      N > ((uint64_t)n_bytes - d.point_data_off) / d.record_len)  // (division: N * record_len may wrap)
    return fail(PCQ_ERR_IO, "file image holds %zu bytes but its header promises %llu points of %u bytes at offset %u",
                n_bytes, (unsigned long long)N, d.record_len, d.point_data_off);
  pcq_file* f = new (std::nothrow) pcq_file();
  if (!f) return fail(PCQ_ERR_NOMEM, "out of host memory");
  f->ctx = ctx;
  f->desc = d;
  f->raw_format = raw;
  f->first_point = first_point;
  f->n_points = n;
  const uint8_t* src = static_cast<const uint8_t*>(file_bytes) + d.point_data_off;
  cudaError_t copy_err = cudaSuccess;
  if (layout == PCQ_LAYOUT_LAS) {
    const size_t bytes = (size_t)n * d.record_len;
    if (cudaMalloc(&f->owned, round_up(bytes, 256) + 256) != cudaSuccess) {
      cudaGetLastError();
      delete f;
      return fail(PCQ_ERR_NOMEM, "cannot allocate %zu bytes of HBM", bytes);
    }
    if (bytes) copy_err = cudaMemcpyAsync(f->owned, src + first_point * d.record_len, bytes, cudaMemcpyHostToDevice, ctx->stream);
    f->rec = static_cast<const uint8_t*>(f->owned);
  } else {
    // only the columns the path reads travel: positions, classification, colour (last.rs:80-90, 114)
    const int rgb_k = rgb_offset_in_record(d.format);
    const size_t pos_b = round_up((size_t)n * 12, 256);
    const size_t cls_b = round_up((size_t)n, 256);
    const size_t rgb_b = rgb_k >= 0 ? round_up((size_t)n * 6, 256) : 0;
    if (cudaMalloc(&f->owned, pos_b + cls_b + rgb_b + 256) != cudaSuccess) {
      cudaGetLastError();
      delete f;
      return fail(PCQ_ERR_NOMEM, "cannot allocate %zu bytes of HBM", pos_b + cls_b + rgb_b);
    }
    uint8_t* base = static_cast<uint8_t*>(f->owned);
    if (n) {
      copy_err = cudaMemcpyAsync(base, src + first_point * 12, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream);
      if (copy_err == cudaSuccess)
        copy_err = cudaMemcpyAsync(base + pos_b, src + (uint64_t)cls_offset_in_record(d.format) * N + first_point, (size_t)n,
                                   cudaMemcpyHostToDevice, ctx->stream);
      if (copy_err == cudaSuccess && rgb_k >= 0)
        copy_err = cudaMemcpyAsync(base + pos_b + cls_b, src + (uint64_t)rgb_k * N + first_point * 6, (size_t)n * 6,
                                   cudaMemcpyHostToDevice, ctx->stream);
    }
    f->rec = base;
    f->cls = base + pos_b;
    f->rgb = rgb_k >= 0 ? base + pos_b + cls_b : nullptr;
  }
  // the caller may reuse its buffer as soon as we return
  if (copy_err == cudaSuccess) copy_err = cudaStreamSynchronize(ctx->stream);
  if (copy_err != cudaSuccess) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(f->owned);
    delete f;
    return fail(PCQ_ERR_CUDA, "staging copy failed: %s", cudaGetErrorString(copy_err));
  }
  ctx->refs++;
  *out = f;
  return PCQ_OK;
}

int pcq_file_wrap_device(pcq_ctx* ctx, const pcq_file_desc* desc, const void* dev_point_data, uint64_t first_point_index,
                         pcq_file** out) {
  if (!ctx || !desc || !out || (!dev_point_data && desc->n_points)) return fail(PCQ_ERR_ARG, "pcq_file_wrap_device: null argument");
  if (desc->layout != PCQ_LAYOUT_LAS && desc->layout != PCQ_LAYOUT_LAST) return fail(PCQ_ERR_ARG, "bad layout");
  if (desc->format > 10) return fail(PCQ_ERR_FORMAT, "Invalid LAS format %u", desc->format);
  if (desc->record_len < format_record_len(desc->format))
    return fail(PCQ_ERR_FORMAT, "point data record length %u too small for format %u", desc->record_len, desc->format);
  pcq_file* f = new (std::nothrow) pcq_file();
  if (!f) return fail(PCQ_ERR_NOMEM, "out of host memory");
  f->ctx = ctx;
  f->desc = *desc;
  f->raw_format = desc->format;
  f->first_point = first_point_index;
  f->n_points = desc->n_points;
  const uint8_t* base = static_cast<const uint8_t*>(dev_point_data);
  const uint64_t N = desc->n_points;
  if (desc->layout == PCQ_LAYOUT_LAS) {
    f->rec = base;
  } else {
    f->rec = base;
    f->cls = base + (uint64_t)cls_offset_in_record(desc->format) * N;
    const int rgb_k = rgb_offset_in_record(desc->format);
    f->rgb = rgb_k >= 0 ? base + (uint64_t)rgb_k * N : nullptr;
  }
  ctx->refs++;
  *out = f;
  return PCQ_OK;
}

int pcq_file_set_scan_base(pcq_file* f, uint64_t scan_base) {
  if (!f) return fail(PCQ_ERR_ARG, "null file");
  f->has_scan_base = true;
  f->scan_base = scan_base;
  return PCQ_OK;
}

int pcq_file_desc_get(const pcq_file* f, pcq_file_desc* out) {
  if (!f || !out) return fail(PCQ_ERR_ARG, "null argument");
  *out = f->desc;
  return PCQ_OK;
}

void pcq_file_release(pcq_file* f) {
  if (!f) return;
  if (f->owned) {
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    cudaFree(f->owned);
  }
  ctx_unref(f->ctx);
  delete f;
}

// ---- chunk index ---------------------------------------------------------------------------------

int pcq_file_build_index(pcq_file* f) {
  if (!f) return fail(PCQ_ERR_ARG, "null file");
  if (!f->index.empty() || f->n_points == 0) return PCQ_OK;
  pcq_ctx* ctx = f->ctx;
  RC(use_device(ctx));
  ChunkIndexArgs a{};
  a.rec = f->rec;
  a.cls = f->cls;
  a.n_points = f->n_points;
  a.n_chunks = (f->n_points + PCQ_INDEX_CHUNK_POINTS - 1) / PCQ_INDEX_CHUNK_POINTS;
  a.layout = f->desc.layout;
  a.record_len = f->desc.layout == PCQ_LAYOUT_LAS ? f->desc.record_len : 12u;
  a.cls_off = cls_offset_in_record(f->desc.format);
  a.align = field_alignment(f->rec, a.record_len);
  a.parts = kIndexPartBox | kIndexPartCls;
  // device scratch and pinned landing buffer live in the context and only grow: cudaMalloc / cudaFree per build
  // cost several times the kernel (274 us for 64 M format-1 records)
  const size_t bytes = (size_t)a.n_chunks * sizeof(pcq_chunk_header);
  if (ctx->index_scratch_cap < bytes) {
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->index_scratch) cudaFree(ctx->index_scratch);
    if (ctx->index_bounce) cudaFreeHost(ctx->index_bounce);
    ctx->index_scratch = ctx->index_bounce = nullptr;
    ctx->index_scratch_cap = 0;
    const size_t cap = round_up(bytes + bytes / 4, 1u << 16);
    if (cudaMalloc(&ctx->index_scratch, cap) != cudaSuccess || cudaMallocHost(&ctx->index_bounce, cap) != cudaSuccess) {
      cudaGetLastError();
      if (ctx->index_scratch) cudaFree(ctx->index_scratch);
      ctx->index_scratch = nullptr;
      return fail(PCQ_ERR_NOMEM, "cannot allocate %zu bytes for chunk headers", cap);
    }
    ctx->index_scratch_cap = cap;
  }
  std::vector<pcq_chunk_header> headers;
  try {
    headers.resize(a.n_chunks);
  } catch (const std::bad_alloc&) {
    return fail(PCQ_ERR_NOMEM, "out of host memory for %llu chunk headers", (unsigned long long)a.n_chunks);
  }
  if (launch_chunk_index(a, static_cast<pcq_chunk_header*>(ctx->index_scratch), ctx->sm_count, ctx->stream) != 0)
    return fail(PCQ_ERR_CUDA, "chunk index launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  ctx->launches++;
  CU(cudaMemcpyAsync(ctx->index_bounce, ctx->index_scratch, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  std::memcpy(headers.data(), ctx->index_bounce, bytes);
  f->index.swap(headers);
  return PCQ_OK;
}

void pcq_file_drop_index(pcq_file* f) {
  if (f) std::vector<pcq_chunk_header>().swap(f->index);
}

int pcq_file_index(const pcq_file* f, const pcq_chunk_header** out_headers, uint64_t* out_n) {
  if (!f || !out_headers || !out_n) return fail(PCQ_ERR_ARG, "null argument");
  *out_headers = f->index.empty() ? nullptr : f->index.data();
  *out_n = f->index.size();
  return PCQ_OK;
}

int pcq_ctx_set_auto_index(pcq_ctx* ctx, uint32_t after_n_scans) {
  if (!ctx) return fail(PCQ_ERR_ARG, "null context");
  ctx->auto_index_after = after_n_scans;
  return PCQ_OK;
}

int pcq_ctx_last_scan_stats(const pcq_ctx* ctx, pcq_scan_stats* out) {
  if (!ctx || !out) return fail(PCQ_ERR_ARG, "null argument");
  *out = ctx->stats;
  return PCQ_OK;
}

// ---- collectors ----------------------------------------------------------------------------------

int pcq_collector_create(pcq_ctx* ctx, int kind, const double gmin[3], const double gmax[3], double cell_size,
                         pcq_collector** out) {
  if (!ctx || !out) return fail(PCQ_ERR_ARG, "pcq_collector_create: null argument");
  if (kind < PCQ_COLLECT_COUNT || kind > PCQ_COLLECT_GRID) return fail(PCQ_ERR_ARG, "bad collector kind %d", kind);
  RC(use_device(ctx));
  pcq_collector* c = new (std::nothrow) pcq_collector();
  if (!c) return fail(PCQ_ERR_NOMEM, "out of host memory");
  c->ctx = ctx;
  c->kind = kind;
  if (cudaMalloc(&c->dev, 256) != cudaSuccess) {
    delete c;
    return fail(PCQ_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemsetAsync(c->dev, 0, 256, ctx->stream);
  ctx->refs++;  // released by pcq_collector_destroy (also on the error paths below)
  if (kind == PCQ_COLLECT_GRID) {
    if (!gmin || !gmax) {
      pcq_collector_destroy(c);
      return fail(PCQ_ERR_ARG, "grid collector needs bounds");
    }
    for (int i = 0; i < 3; ++i) {
      if (gmin[i] > gmax[i]) {
        pcq_collector_destroy(c);
        return fail(PCQ_ERR_PANIC, "AABB::from_min_max: grid bounds have min > max on axis %d", i);
      }
      c->gmin[i] = gmin[i];
      c->gmax[i] = gmax[i];
    }
    c->cell = cell_size;
    int rc = grid_params(gmin, gmax, cell_size, c->dims, c->bits);
    if (rc != PCQ_OK) {
      pcq_collector_destroy(c);
      return rc;
    }
    uint64_t total_bits = c->bits[0] + c->bits[1] + c->bits[2];
    if (c->bits[0] > 62 || c->bits[1] > 62 || c->bits[2] > 62 || total_bits > 64) {
      pcq_collector_destroy(c);
      return fail(PCQ_ERR_GRID, "SparseGrid with %llu+%llu+%llu key bits is not representable",
                  (unsigned long long)c->bits[0], (unsigned long long)c->bits[1], (unsigned long long)c->bits[2]);
    }
    GridDev& g = c->grid;
    for (int i = 0; i < 3; ++i) {
      g.bmin[i] = gmin[i];
      g.bmax[i] = gmax[i];
      g.dims_f[i] = (double)c->dims[i];  // `self.dimensions.x as f64`
      g.mask[i] = c->bits[i] >= 64 ? ~0ull : ((1ull << c->bits[i]) - 1ull);
      g.ext[i] = gmax[i] - gmin[i];
      g.inv_ext[i] = 1.0 / g.ext[i];
    }
    g.fast_div = 1u;
    for (int i = 0; i < 3; ++i)
      if (!(g.ext[i] >= 0x1p-500 && g.ext[i] <= 0x1p500)) g.fast_div = 0u;  // (also NaN, zero extent)
    if (const char* e = std::getenv("PCQ_GRID_FAST_DIV")) g.fast_div = std::atoi(e) ? g.fast_div : 0u;
    g.cell_size = cell_size;
    g.shift_y = (uint32_t)c->bits[0];
    g.shift_z = (uint32_t)(c->bits[0] + c->bits[1]);
    if (g.shift_z >= 64 || g.shift_y >= 64) {
      pcq_collector_destroy(c);
      return fail(PCQ_ERR_GRID, "SparseGrid key shifts of 64 bits are not representable");
    }
    // the cell table is allocated when the collector is first used (ensure_grid_tables): only then is it known how
    // many grids share the HBM — `query --parallel --density` makes one per file, and 64 dense ca13-XL tables are 512 GB
    c->total_bits = total_bits;
  }
  *out = c;
  return PCQ_OK;
}

void pcq_collector_destroy(pcq_collector* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->dev) cudaFree(c->dev);
  if (c->d_out) cudaFree(c->d_out);
  if (c->grid.table) cudaFree(c->grid.table);
  if (c->grid.hkeys) cudaFree(c->grid.hkeys);
  if (c->grid.cands) cudaFree(c->grid.cands);
  if (c->d_final) cudaFree(c->d_final);
  if (c->d_export) cudaFree(c->d_export);
  if (c->d_finlist) cudaFree(c->d_finlist);
  if (c->d_akeys) cudaFree(c->d_akeys);
  if (c->d_aord) cudaFree(c->d_aord);
  if (c->d_astates) cudaFree(c->d_astates);
  if (c->d_log) cudaFree(c->d_log);
  if (c->d_rawlog) cudaFree(c->d_rawlog);
  if (c->h_pts) cudaFreeHost(c->h_pts);
  ctx_unref(c->ctx);
  delete c;
}

}  // extern "C"

namespace pcq {

__global__ void k_zero_blocks(unsigned long long* const* blocks, uint32_t n, uint32_t words) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * words) blocks[i / words][i % words] = 0ull;
}

// pcq_collector_reset for many COUNT / BUFFER collectors of one context with one upload and one launch (a group
// search resets one collector per file and query: 192 memsets per step of the benchmark were 0.5 ms of launch overhead)
int reset_collectors_batched(pcq_ctx* ctx, pcq_collector* const* cols, size_t n) {
  if (n == 0) return PCQ_OK;
  RC(use_device(ctx));
  std::vector<unsigned long long*> ptrs(n);
  for (size_t i = 0; i < n; ++i) {
    pcq_collector* c = cols[i];
    if (c->kind == PCQ_COLLECT_GRID || c->ctx != ctx) return fail(PCQ_ERR_ARG, "batched reset serves count / buffer collectors of one context");
    ptrs[i] = reinterpret_cast<unsigned long long*>(c->dev);
    c->scan_total = 0;
    c->out_len = 0;
    c->cand_len = 0;
    c->final_valid = false;
    c->final_n = 0;
    c->scan_hi = 0;
  }
  void* d_ptrs = nullptr;
  RC(upload(ctx, ptrs.data(), n * sizeof(void*), &d_ptrs));
  const uint32_t words = 256 / 8;
  k_zero_blocks<<<(unsigned)((n * words + 255) / 256), 256, 0, ctx->stream>>>(static_cast<unsigned long long* const*>(d_ptrs), (uint32_t)n, words);
  if (cudaGetLastError() != cudaSuccess) return fail(PCQ_ERR_CUDA, "k_zero_blocks launch failed");
  ctx->launches++;
  return PCQ_OK;
}

}  // namespace pcq

extern "C" {

int pcq_collector_reset(pcq_collector* c) {
  if (!c) return fail(PCQ_ERR_ARG, "null collector");
  pcq_ctx* ctx = c->ctx;
  RC(use_device(ctx));
  CU(cudaMemsetAsync(c->dev, 0, 256, ctx->stream));
  c->scan_total = 0;
  c->out_len = 0;
  c->cand_len = 0;
  c->final_valid = false;
  c->table_holds_winners = false;  // (the table is cleared below)
  c->final_n = 0;
  c->akeys.clear();
  c->astates.clear();
  c->a_slots = 0;
  c->scan_hi = 0;
  c->pass_mode = 0;
  c->rawlog_len = 0;
  c->own_parts = c->own_me = 0;
  if (c->kind == PCQ_COLLECT_GRID && c->grid.table) {  // (a sub-box table is kept: ensure_grid_tables checks the next box)
    CU(cudaMemsetAsync(c->grid.table, 0xFF, c->grid.table_slots * 8ull, ctx->stream));
    if (c->grid.hkeys) CU(cudaMemsetAsync(c->grid.hkeys, 0xFF, c->grid.table_slots * 8ull, ctx->stream));
  }
  return PCQ_OK;
}

int pcq_grid_cells_under_box(const double gmin[3], const double gmax[3], double cell_size, const double box_min[3],
                             const double box_max[3], uint64_t lo[3], uint64_t n[3]) {
  if (!gmin || !gmax || !box_min || !box_max || !lo || !n) return fail(PCQ_ERR_ARG, "null argument");
  uint64_t dims[3], bits[3];
  RC(grid_params(gmin, gmax, cell_size, dims, bits));
  GridDev g{};
  for (int a = 0; a < 3; ++a) {
    if (bits[a] > 62) return fail(PCQ_ERR_GRID, "SparseGrid axis with %llu key bits is not representable", (unsigned long long)bits[a]);
    g.bmin[a] = gmin[a];
    g.bmax[a] = gmax[a];
    g.dims_f[a] = (double)dims[a];
    g.ext[a] = gmax[a] - gmin[a];
    g.mask[a] = (1ull << bits[a]) - 1ull;
  }
  for (int a = 0; a < 3; ++a)
    if (!cell_range(g, a, box_min[a], box_max[a], &lo[a], &n[a])) {  // no usable range on this axis: every cell
      lo[a] = 0;
      n[a] = g.mask[a] + 1ull;
    }
  return PCQ_OK;
}

int pcq_collectors_reset(pcq_collector* const* collectors, uint32_t n) {
  if (!collectors && n) return fail(PCQ_ERR_ARG, "null collectors");
  std::vector<pcq_collector*> rest;
  for (uint32_t i = 0; i < n; ++i) {
    if (!collectors[i]) return fail(PCQ_ERR_ARG, "null collector");
    if (collectors[i]->kind == PCQ_COLLECT_GRID)
      RC(pcq_collector_reset(collectors[i]));
    else
      rest.push_back(collectors[i]);
  }
  // per context, in one launch each
  while (!rest.empty()) {
    pcq_ctx* ctx = rest.front()->ctx;
    std::vector<pcq_collector*> mine, other;
    for (pcq_collector* c : rest) (c->ctx == ctx ? mine : other).push_back(c);
    RC(reset_collectors_batched(ctx, mine.data(), mine.size()));
    rest.swap(other);
  }
  return PCQ_OK;
}

int pcq_collector_point_count(pcq_collector* c, uint64_t* out) {
  if (!c || !out) return fail(PCQ_ERR_ARG, "null argument");
  RC(use_device(c->ctx));
  if (c->kind == PCQ_COLLECT_GRID) {
    RC(grid_finalize(c));
    *out = c->final_n;  // self.grid.points().count(), collect_points.rs:124-126
    return PCQ_OK;
  }
  DevBlock b;
  RC(read_devblock(c, &b));
  *out = b.count;
  return PCQ_OK;
}

int pcq_collector_points_device(pcq_collector* c, const void** out_dev_points, uint64_t* out_n) {
  if (!c || !out_dev_points || !out_n) return fail(PCQ_ERR_ARG, "null argument");
  RC(use_device(c->ctx));
  *out_dev_points = nullptr;
  *out_n = 0;
  if (c->kind == PCQ_COLLECT_COUNT) return PCQ_OK;  // points() == None
  if (c->kind == PCQ_COLLECT_BUFFER) {
    CU(cudaStreamSynchronize(c->ctx->stream));
    *out_dev_points = c->d_out;
    *out_n = c->out_len;
    return PCQ_OK;
  }
  RC(grid_finalize(c));
  *out_dev_points = c->d_final;
  *out_n = c->final_n;
  return PCQ_OK;
}

int pcq_collector_points(pcq_collector* c, const pcq_point** out_points, uint64_t* out_n) {
  if (!c || !out_points || !out_n) return fail(PCQ_ERR_ARG, "null argument");
  const void* dptr = nullptr;
  uint64_t n = 0;
  RC(pcq_collector_points_device(c, &dptr, &n));
  *out_points = nullptr;
  *out_n = 0;
  if (n == 0) return PCQ_OK;
  if (c->h_cap < n) {
    if (c->h_pts) cudaFreeHost(c->h_pts);
    c->h_pts = nullptr;
    c->h_cap = 0;
    uint64_t cap = std::max<uint64_t>(n, 4096);
    if (cudaMallocHost(&c->h_pts, cap * 31ull) != cudaSuccess) {
      cudaGetLastError();
      return fail(PCQ_ERR_NOMEM, "cannot pin %llu bytes of host memory", (unsigned long long)(cap * 31ull));
    }
    c->h_cap = cap;
  }
  CU(cudaMemcpyAsync(c->h_pts, dptr, n * 31ull, cudaMemcpyDeviceToHost, c->ctx->stream));
  CU(cudaStreamSynchronize(c->ctx->stream));
  *out_points = static_cast<const pcq_point*>(c->h_pts);
  *out_n = n;
  return PCQ_OK;
}

// ---- the scan ------------------------------------------------------------------------------------

namespace {

bool chunk_may_match(const pcq_chunk_header& h, const pcq_query* q, const SegmentPlan& plan) {
  if (q->kind == PCQ_QUERY_BOUNDS) {
    // a record matches iff lo <= v <= hi on every axis (las.rs:106-119): the chunk's extent must overlap that box
    for (int a = 0; a < 3; ++a)
      if (h.hi[a] < plan.lo[a] || h.lo[a] > plan.hi[a]) return false;
    return true;
  }
  return ((h.cls_bits[q->cls >> 5] >> (q->cls & 31u)) & 1u) != 0;
}

struct ChunkRun {
  uint64_t first, end;  // chunks [first, end)
};

// Runs of chunks of an indexed file that can hold a match.  Runs less than `gap` chunks apart are joined (scanning a
// few chunks that cannot match is cheaper than another point range in the launch); returns the chunks that may match.
uint64_t surviving_runs(const pcq_chunk_header* headers, uint64_t n, const pcq_query* q, const SegmentPlan& plan, uint64_t gap,
                        std::vector<ChunkRun>& runs) {
  runs.clear();
  uint64_t may = 0;
  for (uint64_t c = 0; c < n; ++c) {
    if (!chunk_may_match(headers[c], q, plan)) continue;
    ++may;
    if (!runs.empty() && c - runs.back().end < gap)
      runs.back().end = c + 1;
    else
      runs.push_back({c, c + 1});
  }
  return may;
}

constexpr uint64_t kIndexJoinGap = 4;        // chunks
constexpr uint64_t kHostIndexJoinGap = 16;   // host-staged: every run is a PCIe copy of its own, keep them megabytes long
constexpr size_t kIndexMaxRunsPerFile = 4096;

}  // namespace

int pcq_index_filter(const pcq_chunk_header* headers, uint64_t n_chunks, const pcq_file_desc* desc, const pcq_query* query,
                     uint64_t join_gap, uint64_t* runs, uint64_t cap_runs, uint64_t* n_runs, uint64_t* n_may) {
  if ((!headers && n_chunks) || !desc || !query || !n_runs || (!runs && cap_runs)) return fail(PCQ_ERR_ARG, "null argument");
  if (query->kind == PCQ_QUERY_BOUNDS)
    for (int i = 0; i < 3; ++i)
      if (query->qmin[i] > query->qmax[i]) return fail(PCQ_ERR_PANIC, "AABB::from_min_max: query bounds have min > max on axis %d", i);
  *n_runs = 0;
  if (n_may) *n_may = 0;
  SegmentPlan plan;
  RC(plan_file(*desc, desc->format, query, &plan));
  if (plan.skip) return PCQ_OK;  // the file's header already excludes the query (las.rs:82-84)
  std::vector<ChunkRun> found;
  const uint64_t may = surviving_runs(headers, n_chunks, query, plan, join_gap ? join_gap : 1, found);
  if (n_may) *n_may = may;
  *n_runs = found.size();
  for (size_t i = 0; i < found.size() && i < cap_runs; ++i) {
    runs[2 * i] = found[i].first;
    runs[2 * i + 1] = found[i].end;
  }
  return PCQ_OK;
}

static int check_search_args(pcq_ctx* ctx, uint32_t n_files, const pcq_query* q, pcq_collector* const* collectors,
                             uint32_t n_collectors) {
  if (!ctx || !q || !collectors) return fail(PCQ_ERR_ARG, "pcq_search: null argument");
  if (n_collectors != 1 && n_collectors != n_files)
    return fail(PCQ_ERR_ARG, "n_collectors must be 1 (sequential) or n_files (parallel), got %u for %u files", n_collectors, n_files);
  for (uint32_t l = 0; l < n_collectors; ++l) {
    if (!collectors[l]) return fail(PCQ_ERR_ARG, "null collector %u", l);
    if (collectors[l]->kind != collectors[0]->kind) return fail(PCQ_ERR_ARG, "collectors of one search must be of one kind");
    if (collectors[l]->ctx != ctx) return fail(PCQ_ERR_ARG, "collector %u belongs to another context", l);
  }
  if (q->kind == PCQ_QUERY_BOUNDS)
    for (int i = 0; i < 3; ++i)
      if (q->qmin[i] > q->qmax[i]) return fail(PCQ_ERR_PANIC, "AABB::from_min_max: query bounds have min > max on axis %d", i);
  return PCQ_OK;
}

int pcq_search_files(pcq_ctx* ctx, pcq_file* const* files, uint32_t n_files, const pcq_query* query,
                     pcq_collector* const* collectors, uint32_t n_collectors) {
  RC(check_search_args(ctx, n_files, query, collectors, n_collectors));
  if (n_files == 0) return PCQ_OK;
  if (!files) return fail(PCQ_ERR_ARG, "null files");
  RC(use_device(ctx));

  std::vector<Segment> segs;
  segs.reserve(n_files);
  std::vector<uint64_t> lane_points(n_collectors, 0);
  LaneBoxes lane_box(n_collectors);
  bool frac_known = true;
  double frac_pts = 0.0, all_pts = 0.0;
  pcq_scan_stats st{};
  std::vector<ChunkRun> runs;
  for (uint32_t i = 0; i < n_files; ++i) {
    pcq_file* f = files[i];
    if (!f) return fail(PCQ_ERR_ARG, "null file %u", i);
    if (f->ctx != ctx) return fail(PCQ_ERR_ARG, "file %u belongs to another context", i);
    const uint32_t lane = n_collectors == 1 ? 0 : i;
    lane_box.add(lane, f->desc);
    pcq_collector* c = collectors[lane];
    const uint64_t base = f->has_scan_base ? f->scan_base : c->scan_total;
    if (!f->has_scan_base) c->scan_total += f->n_points;
    SegmentPlan plan;
    RC(plan_file(f->desc, f->raw_format, query, &plan));
    if (plan.skip || f->n_points == 0 || c->pass_mode == 2) continue;
    if (ctx->auto_index_after != 0 && f->index.empty() && f->scans >= ctx->auto_index_after) RC(pcq_file_build_index(f));
    f->scans++;
    st.points_total += f->n_points;

    // chunk index: launch over the runs of chunks that can hold a match.  Every run keeps its place in the scan
    // order (scan_base) and on its lane, so counts, record streams and density ties are those of the full scan.
    bool whole = true;
    if (!f->index.empty()) {
      const uint64_t n_chunks = f->index.size();
      uint64_t gap = kIndexJoinGap;
      const uint64_t may = surviving_runs(f->index.data(), n_chunks, query, plan, gap, runs);
      while (runs.size() > kIndexMaxRunsPerFile) surviving_runs(f->index.data(), n_chunks, query, plan, gap *= 4, runs);
      uint64_t kept = 0;
      for (const ChunkRun& r : runs) kept += r.end - r.first;
      st.chunks_total += n_chunks;
      if (may == 0) {
        st.chunks_skipped += n_chunks;
        continue;
      }
      // below 90 % a fragmented launch pays; counting a class in a LAST column (1 byte per point, 41 us for 128 M
      // points) is over before the tile-scheduled launch over the runs has started unless most of it can go
      const bool byte_count = query->kind == PCQ_QUERY_CLASS && f->desc.layout == PCQ_LAYOUT_LAST && c->kind == PCQ_COLLECT_COUNT;
      if (byte_count ? kept * 8 < n_chunks : kept * 10 < n_chunks * 9) {
        whole = false;
        st.chunks_skipped += n_chunks - kept;
        const uint64_t R = f->desc.layout == PCQ_LAYOUT_LAS ? f->desc.record_len : 12u;
        for (const ChunkRun& r : runs) {
          const uint64_t p0 = r.first * PCQ_INDEX_CHUNK_POINTS;
          const uint64_t n = std::min<uint64_t>(r.end * PCQ_INDEX_CHUNK_POINTS, f->n_points) - p0;
          Segment s;
          fill_segment(&s, f->desc, f->rec + p0 * R, f->cls ? f->cls + p0 : nullptr, f->rgb ? f->rgb + p0 * 6 : nullptr, n, plan,
                       lane, base + p0, query->kind);
          lane_points[lane] += n;
          all_pts += (double)n;
          segs.push_back(s);
        }
      }
    }
    // the matches the header box promises are all in the ranges that are left
    const double fr = expected_match_fraction(f->desc, query);
    if (fr < 0.0) frac_known = false;
    frac_pts += (fr < 0.0 ? 0.0 : fr) * (double)f->n_points;
    if (whole) {
      Segment s;
      fill_segment(&s, f->desc, f->rec, f->cls, f->rgb, f->n_points, plan, lane, base, query->kind);
      lane_points[lane] += f->n_points;
      all_pts += (double)f->n_points;
      segs.push_back(s);
    }
  }
  st.points_scanned = (uint64_t)all_pts;
  st.segments = (uint32_t)segs.size();
  ctx->stats = st;
  lane_box.cut(query);
  return run_batch(ctx, segs, query, collectors, n_collectors, lane_points,
                   frac_known && all_pts > 0.0 ? std::min(1.0, frac_pts / all_pts) : -1.0, &lane_box.v);
}

int pcq_host_alloc(size_t n_bytes, void** out) {
  if (!out) return fail(PCQ_ERR_ARG, "null out");
  if (cudaMallocHost(out, n_bytes ? n_bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return fail(PCQ_ERR_NOMEM, "cannot pin %zu bytes of host memory", n_bytes);
  }
  return PCQ_OK;
}

void pcq_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int pcq_host_register(void* p, size_t n_bytes) {
  if (!p || !n_bytes) return fail(PCQ_ERR_ARG, "pcq_host_register: null argument");
  const cudaError_t e = cudaHostRegister(p, n_bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(PCQ_ERR_NOMEM, "cannot pin %zu bytes of host memory: %s", n_bytes, cudaGetErrorString(e));
  }
  return PCQ_OK;
}

int pcq_host_unregister(void* p) {
  if (!p) return PCQ_OK;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return fail(PCQ_ERR_ARG, "pcq_host_unregister: not a registered range");
  }
  return PCQ_OK;
}

int pcq_ctx_bind_host_thread(pcq_ctx* ctx, int* out_node) {
  if (!ctx) return fail(PCQ_ERR_ARG, "null context");
  if (out_node) *out_node = -1;
#ifdef __linux__
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), ctx->device) != cudaSuccess) {
    cudaGetLastError();
    return PCQ_OK;
  }
  for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
  char path[128];
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  int node = -1;
  if (FILE* f = std::fopen(path, "r")) {
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
  }
  if (node < 0) return PCQ_OK;  // single-node box or a platform that does not say
  std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  cpu_set_t set;
  CPU_ZERO(&set);
  int n_cpus = 0;
  if (FILE* f = std::fopen(path, "r")) {
    int a = 0, b = 0;
    for (;;) {
      if (std::fscanf(f, "%d", &a) != 1) break;
      b = a;
      int ch = std::fgetc(f);
      if (ch == '-') {
        if (std::fscanf(f, "%d", &b) != 1) b = a;
        ch = std::fgetc(f);
      }
      for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
        CPU_SET(c, &set);
        ++n_cpus;
      }
      if (ch != ',') break;
    }
    std::fclose(f);
  }
  // only CPUs this process may use anyway (a container's cpuset)
  cpu_set_t allowed;
  if (n_cpus && sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
    cpu_set_t both;
    CPU_AND(&both, &set, &allowed);
    if (CPU_COUNT(&both) > 0) sched_setaffinity(0, sizeof(both), &both);
  }
  // first-touch placement follows the thread; ask for the node explicitly as well (MPOL_PREFERRED = 1)
  if (node < 64) {
    unsigned long mask = 1ul << node;
    syscall(SYS_set_mempolicy, 1, &mask, (unsigned long)(8 * sizeof(mask)));
  }
  if (out_node) *out_node = node;
#endif
  return PCQ_OK;
}

// Host-staged scan: file images stream through a ring of HBM chunk buffers; the copy of chunk k+2
// overlaps the scan of chunk k.  Replaces mmap + page-fault driven reads (las.rs:24-31).  Several
// queries can share one pass: every chunk is scanned by each query that needs its file while it is
// resident, so the bytes cross PCIe once per batch instead of once per query.
}  // extern "C"
namespace pcq {
int search_host_multi(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes, const char* const* exts,
                      uint32_t n_files, const pcq_query* queries, uint32_t n_queries, pcq_collector* const* collectors,
                      uint32_t n_collectors, pcq_host_index* hix, const HostRange* ranges) {
  if (n_queries == 0) return PCQ_OK;
  for (uint32_t q = 0; q < n_queries; ++q)
    RC(check_search_args(ctx, n_files, queries + q, collectors + (size_t)q * n_collectors, n_collectors));
  if (n_files == 0) return PCQ_OK;
  if (!file_bytes || !n_bytes || !exts) return fail(PCQ_ERR_ARG, "null argument");
  RC(use_device(ctx));

  struct Run {
    uint64_t first, n;  // points
  };
  struct Piece {  // what one ring buffer holds: runs [run0, run0 + n_runs) of one file, packed back to back
    uint32_t file;
    uint32_t run0, n_runs;
  };
  struct FilePlan {
    pcq_file_desc d;
    uint8_t raw;
    std::vector<SegmentPlan> plan;  // per query
    std::vector<uint64_t> base;     // per query: scan index of the file's point 0 in its collector
    uint32_t lane;
    int layout;
    bool any, need_pos, need_cls, need_rgb;
    bool build_box, build_cls;      // this pass computes that part of the file's chunk headers
    double kept_fraction;           // points that cross PCIe / points of the file
    uint64_t per_point;             // bytes of a point that cross PCIe
  };
  // Pass 1: everything that can fail on a file (header, length, per-query planning) before any collector or
  // index state is touched, so that a failed call leaves the collectors' scan bases where they were.
  std::vector<FilePlan> fps(n_files);
  uint64_t max_per_point = 1;
  for (uint32_t i = 0; i < n_files; ++i) {
    FilePlan& fp = fps[i];
    fp.layout = layout_of_ext(exts[i]);
    if (fp.layout < 0) return fail(PCQ_ERR_FORMAT, "Unsupported file extension \"%s\"", exts[i] ? exts[i] : "");
    if (!file_bytes[i]) return fail(PCQ_ERR_ARG, "null file image %u", i);
    RC(parse_header(file_bytes[i], n_bytes[i], fp.layout, 1, &fp.d, &fp.raw));
    if ((uint64_t)fp.d.point_data_off > (uint64_t)n_bytes[i] || fp.d.record_len == 0 ||
        fp.d.n_points > ((uint64_t)n_bytes[i] - fp.d.point_data_off) / fp.d.record_len)
      return fail(PCQ_ERR_IO, "file image %u is shorter than its header promises", i);
    fp.lane = n_collectors == 1 ? 0 : i;
    fp.plan.resize(n_queries);
    fp.base.assign(n_queries, 0);
    fp.any = fp.need_pos = fp.need_cls = fp.need_rgb = false;
    fp.build_box = fp.build_cls = false;
    fp.kept_fraction = 1.0;
    for (uint32_t q = 0; q < n_queries; ++q) {
      const pcq_collector* c = collectors[(size_t)q * n_collectors + fp.lane];
      RC(plan_file(fp.d, fp.raw, queries + q, &fp.plan[q]));
      if (fp.plan[q].skip || fp.d.n_points == 0 || c->pass_mode == 2 || (ranges && ranges[i].n_points == 0)) {
        fp.plan[q].skip = true;
        continue;
      }
      const bool emit = c->kind != PCQ_COLLECT_COUNT;
      fp.any = true;
      fp.need_pos |= queries[q].kind == PCQ_QUERY_BOUNDS || emit;
      fp.need_cls |= queries[q].kind == PCQ_QUERY_CLASS || emit;
      fp.need_rgb |= emit && rgb_offset_in_record(fp.d.format) >= 0;
    }
    fp.per_point = fp.layout == PCQ_LAYOUT_LAS
                       ? fp.d.record_len
                       : (uint64_t)(fp.need_pos ? 12 : 0) + (fp.need_cls ? 1 : 0) + (fp.need_rgb ? 6 : 0);
    if (fp.any) max_per_point = std::max(max_per_point, fp.per_point);
  }

  // 256 MB pieces keep a PCIe 5 x16 link at ~54 GB/s (64 MB: 53, 16 MB: 49); small inputs get small buffers.  A piece
  // never holds less than one index chunk (PCQ_INDEX_CHUNK_POINTS points, each column rounded up to 256 bytes), so a
  // long record length (extra bytes) raises the floor.
  size_t chunk_bytes = 256u << 20;
  const size_t chunk_floor = round_up((size_t)PCQ_INDEX_CHUNK_POINTS * max_per_point + 768 + 1024, 1u << 20);
  const bool chunk_forced = std::getenv("PCQ_CHUNK_MB") != nullptr;
  if (chunk_forced) {
    chunk_bytes = (size_t)std::max(1, std::atoi(std::getenv("PCQ_CHUNK_MB"))) << 20;
  } else {
    size_t largest = 0;
    for (uint32_t i = 0; i < n_files; ++i) largest = std::max(largest, n_bytes[i]);
    chunk_bytes = std::min(chunk_bytes, std::max<size_t>(round_up(largest + 4096, 1u << 20), 4u << 20));
  }
  chunk_bytes = std::max(chunk_bytes, chunk_floor);
  if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  if (chunk_forced ? ctx->chunk_cap != chunk_bytes : ctx->chunk_cap < chunk_bytes) {
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    for (int b = 0; b < kChunkBuffers; ++b) {
      if (ctx->chunk[b]) cudaFree(ctx->chunk[b]);
      ctx->chunk[b] = nullptr;
    }
    ctx->chunk_cap = 0;
    for (int b = 0; b < kChunkBuffers; ++b)
      if (cudaMalloc(&ctx->chunk[b], chunk_bytes + 1024) != cudaSuccess) {
        cudaGetLastError();
        return fail(PCQ_ERR_NOMEM, "cannot allocate %zu bytes of HBM for the staging ring", chunk_bytes);
      }
    ctx->chunk_cap = chunk_bytes;
  }
  chunk_bytes = ctx->chunk_cap;  // (a larger ring left by an earlier call is used whole)
  for (int b = 0; b < kChunkBuffers; ++b) {
    if (!ctx->chunk_copied[b]) CU(cudaEventCreateWithFlags(&ctx->chunk_copied[b], cudaEventDisableTiming));
    if (!ctx->chunk_free[b]) CU(cudaEventCreateWithFlags(&ctx->chunk_free[b], cudaEventDisableTiming));
  }
  // file images in pageable memory (a memory-mapped file, las.rs:24-31) go through a pinned bounce ring
  std::vector<char> pageable(n_files, 0);
  bool any_pageable = false;
  for (uint32_t i = 0; i < n_files; ++i) {
    // tested where this call's copies start: a rank of a sharded scan may have pinned only its own range of the image
    const uint8_t* probe = static_cast<const uint8_t*>(file_bytes[i]);
    if (ranges && fps[i].any && ranges[i].n_points != 0 && fps[i].layout == PCQ_LAYOUT_LAS)
      probe += fps[i].d.point_data_off + std::min<uint64_t>(ranges[i].first_point, fps[i].d.n_points) * fps[i].d.record_len;
    else if (ranges && (!fps[i].any || ranges[i].n_points == 0))
      continue;
    pageable[i] = is_pinned_host(probe) ? 0 : 1;
    any_pageable |= pageable[i] != 0;
  }
  if (std::getenv("PCQ_NO_BOUNCE")) any_pageable = false, std::fill(pageable.begin(), pageable.end(), 0);
  if (any_pageable) {
    const size_t want = chunk_bytes + 1024;
    if (ctx->bounce_cap < want) {
      CU(cudaStreamSynchronize(ctx->copy_stream));
      for (int b = 0; b < kChunkBuffers; ++b) {
        if (ctx->bounce[b]) cudaFreeHost(ctx->bounce[b]);
        ctx->bounce[b] = nullptr;
        ctx->bounce_busy[b] = false;
      }
      ctx->bounce_cap = 0;
      for (int b = 0; b < kChunkBuffers; ++b)
        if (cudaMallocHost(&ctx->bounce[b], want) != cudaSuccess) {
          cudaGetLastError();
          return fail(PCQ_ERR_NOMEM, "cannot pin %zu bytes of host memory for the staging ring", want);
        }
      ctx->bounce_cap = want;
    }
    for (int b = 0; b < kChunkBuffers; ++b)
      if (!ctx->bounce_done[b]) CU(cudaEventCreateWithFlags(&ctx->bounce_done[b], cudaEventDisableTiming));
  }

  if (hix) {
    if (hix->ctx != ctx) return fail(PCQ_ERR_ARG, "host index belongs to another context");
    if (hix->pending) {  // headers the last building pass left on the device
      CU(cudaStreamSynchronize(ctx->stream));
      std::vector<pcq_chunk_header> tmp;
      for (pcq_host_index::File& xf : hix->files) {
        if (!xf.unfetched) continue;
        tmp.resize(xf.n_chunks);
        CU(cudaMemcpy(tmp.data(), xf.d_headers, (size_t)xf.n_chunks * sizeof(pcq_chunk_header), cudaMemcpyDeviceToHost));
        if (xf.unfetched & kIndexPartBox) xf.box = tmp;
        if (xf.unfetched & kIndexPartCls) xf.cls = tmp;
        xf.unfetched = 0;
      }
      hix->pending = false;
    }
    if (hix->files.empty()) hix->files.resize(n_files);
    if (hix->files.size() != n_files)
      return fail(PCQ_ERR_ARG, "host index covers %zu files, this search names %u", hix->files.size(), n_files);
    // chunk headers are only good for the image they were made from: an entry is keyed on the image's address,
    // length and the header fields the filter depends on, and dropped when any of them changed
    for (uint32_t i = 0; i < n_files; ++i) {
      pcq_host_index::File& xf = hix->files[i];
      const FilePlan& fp = fps[i];
      uint64_t key = 1469598103934665603ull;
      auto mixin = [&key](const void* p, size_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(p);
        for (size_t k = 0; k < n; ++k) key = (key ^ b[k]) * 1099511628211ull;
      };
      const uintptr_t addr = reinterpret_cast<uintptr_t>(file_bytes[i]);
      mixin(&addr, sizeof(addr));
      mixin(&n_bytes[i], sizeof(size_t));
      mixin(&fp.d, sizeof(fp.d));
      if (xf.n_chunks != 0 && xf.key != key) {  // another image sits at this position now
        xf.has_box = xf.has_cls = false;
        xf.unfetched = 0;
        std::vector<pcq_chunk_header>().swap(xf.box);
        std::vector<pcq_chunk_header>().swap(xf.cls);
        if (xf.d_headers) {
          CU(cudaStreamSynchronize(ctx->stream));
          cudaFree(xf.d_headers);
        }
        xf.d_headers = nullptr;
        xf.n_chunks = 0;
      }
      xf.key = key;
    }
  }

  // Pass 2: scan bases, chunk-header use / build decisions, and the pieces that cross PCIe.
  std::vector<Run> runs;
  std::vector<Piece> pieces;
  std::vector<ChunkRun> cruns;
  pcq_scan_stats stats{};
  for (uint32_t i = 0; i < n_files; ++i) {
    FilePlan& fp = fps[i];
    const int layout = fp.layout;
    for (uint32_t q = 0; q < n_queries; ++q) {
      pcq_collector* c = collectors[(size_t)q * n_collectors + fp.lane];
      if (ranges) {
        fp.base[q] = ranges[i].scan_base;  // (a sharded scan: the caller knows where the file sits in the scan order)
      } else {
        fp.base[q] = c->scan_total;
        c->scan_total += fp.d.n_points;
      }
    }
    if (!fp.any) continue;
    const uint64_t N = fp.d.n_points;
    // a sharded scan covers the point range [r_first, r_end) of the file only (group.cu)
    uint64_t r_first = 0, r_end = N;
    if (ranges) {
      r_first = std::min<uint64_t>(ranges[i].first_point, N);
      r_end = r_first + std::min<uint64_t>(ranges[i].n_points, N - r_first);
      if (r_first % PCQ_INDEX_CHUNK_POINTS != 0)
        return fail(PCQ_ERR_ARG, "point range of file %u starts at %llu: ranges start on multiples of %u points", i,
                    (unsigned long long)r_first, PCQ_INDEX_CHUNK_POINTS);
      if (hix) return fail(PCQ_ERR_ARG, "a host index cannot be combined with point ranges");
      if (r_end == r_first) continue;
    }
    stats.points_total += r_end - r_first;

    // chunk headers of this file: use them when every query of the batch can be filtered, else build what the
    // columns of this pass allow ("while scanning first (without an index) ...", improvements.md:6)
    pcq_host_index::File* xf = hix ? &hix->files[i] : nullptr;
    bool filter = false;
    if (xf) {
      xf->n_points = N;
      xf->n_chunks = (N + PCQ_INDEX_CHUNK_POINTS - 1) / PCQ_INDEX_CHUNK_POINTS;
      filter = true;
      for (uint32_t q = 0; q < n_queries; ++q)
        if (!fp.plan[q].skip && !(queries[q].kind == PCQ_QUERY_BOUNDS ? xf->has_box : xf->has_cls)) filter = false;
      if (!filter) {
        const bool las = layout == PCQ_LAYOUT_LAS;
        fp.build_box = !xf->has_box && (las || fp.need_pos);
        fp.build_cls = !xf->has_cls && (las || fp.need_cls);
        const size_t bytes = (size_t)xf->n_chunks * sizeof(pcq_chunk_header);
        if ((fp.build_box || fp.build_cls) && !xf->d_headers && cudaMalloc(reinterpret_cast<void**>(&xf->d_headers), bytes) != cudaSuccess) {
          cudaGetLastError();
          return fail(PCQ_ERR_NOMEM, "cannot allocate %zu bytes of HBM for chunk headers", bytes);
        }
      }
    }

    const uint64_t per_point = fp.per_point;
    // a run's columns are rounded up to 256 bytes each: three roundings per run at most
    uint64_t pts = (chunk_bytes - 1024 - 768) / per_point;
    pts = std::max<uint64_t>(PCQ_INDEX_CHUNK_POINTS, pts / PCQ_INDEX_CHUNK_POINTS * PCQ_INDEX_CHUNK_POINTS);
    auto run_bytes = [&](uint64_t n) -> uint64_t {
      if (layout == PCQ_LAYOUT_LAS) return round_up(n * fp.d.record_len, 256);
      return (fp.need_pos ? round_up(n * 12, 256) : 0) + (fp.need_cls ? round_up(n, 256) : 0) + (fp.need_rgb ? round_up(n * 6, 256) : 0);
    };
    cruns.clear();
    if (filter) {
      // a chunk crosses PCIe when some query of the batch may find a match in it; short gaps are copied along
      stats.chunks_total += xf->n_chunks;
      for (uint64_t c = 0; c < xf->n_chunks; ++c) {
        bool may = false;
        for (uint32_t q = 0; q < n_queries && !may; ++q)
          if (!fp.plan[q].skip)
            may = chunk_may_match(queries[q].kind == PCQ_QUERY_BOUNDS ? xf->box[c] : xf->cls[c], queries + q, fp.plan[q]);
        if (!may) continue;
        if (!cruns.empty() && c - cruns.back().end < kHostIndexJoinGap)
          cruns.back().end = c + 1;
        else
          cruns.push_back({c, c + 1});
      }
      uint64_t kept = 0;
      for (const ChunkRun& r : cruns) kept += r.end - r.first;
      stats.chunks_skipped += xf->n_chunks - kept;
    } else {
      cruns.push_back({r_first / PCQ_INDEX_CHUNK_POINTS, (r_end + PCQ_INDEX_CHUNK_POINTS - 1) / PCQ_INDEX_CHUNK_POINTS});
    }
    uint64_t kept_points = 0;
    uint64_t room = 0;
    for (const ChunkRun& r : cruns) {
      const uint64_t p0 = r.first * PCQ_INDEX_CHUNK_POINTS, p1 = std::min<uint64_t>(r.end * PCQ_INDEX_CHUNK_POINTS, r_end);
      for (uint64_t first = p0; first < p1; first += pts) {
        const uint64_t n = std::min<uint64_t>(pts, p1 - first);
        const uint64_t need = run_bytes(n);
        if (pieces.empty() || pieces.back().file != i || need > room) {
          pieces.push_back({i, (uint32_t)runs.size(), 0});
          room = chunk_bytes - 1024;
        }
        if (need > room) return fail(PCQ_ERR_ARG, "internal: a run of %llu bytes does not fit a %zu-byte ring buffer", (unsigned long long)need, chunk_bytes);
        runs.push_back({first, n});
        pieces.back().n_runs++;
        room -= need;
        kept_points += n;
      }
    }
    stats.points_scanned += kept_points;
    fp.kept_fraction = (double)kept_points / (double)(r_end - r_first);
  }
  struct Staged {
    const uint8_t *rec, *cls, *rgb;
  };
  std::vector<Staged> staged(runs.size());
  auto issue_copy = [&](size_t j) -> int {
    const Piece& pc = pieces[j];
    const FilePlan& fp = fps[pc.file];
    const int b = (int)(j % kChunkBuffers);
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_free[b], 0));
    uint8_t* dst = static_cast<uint8_t*>(ctx->chunk[b]);
    const uint8_t* src = static_cast<const uint8_t*>(file_bytes[pc.file]) + fp.d.point_data_off;
    const uint64_t N = fp.d.n_points;
    const bool via_bounce = pageable[pc.file] != 0;
    uint8_t* hb = via_bounce ? static_cast<uint8_t*>(ctx->bounce[b]) : nullptr;
    // the slot's last H2D (possibly issued by an EARLIER call on this context: COUNT searches return without
    // synchronising) must have left it before host threads overwrite it
    if (via_bounce && ctx->bounce_busy[b]) {
      CU(cudaEventSynchronize(ctx->bounce_done[b]));
      ctx->bounce_busy[b] = false;
    }
    // one column (or the record block): host bytes -> [pinned bounce slot ->] device chunk at offset o
    auto put = [&](size_t o, const uint8_t* from, size_t n) -> int {
      if (via_bounce) {
        parallel_memcpy(hb + o, from, n);
        CU(cudaMemcpyAsync(dst + o, hb + o, n, cudaMemcpyHostToDevice, ctx->copy_stream));
      } else {
        CU(cudaMemcpyAsync(dst + o, from, n, cudaMemcpyHostToDevice, ctx->copy_stream));
      }
      return PCQ_OK;
    };
    size_t o = 0;
    for (uint32_t r = pc.run0; r < pc.run0 + pc.n_runs; ++r) {
      const Run& rn = runs[r];
      Staged st{nullptr, nullptr, nullptr};
      if (fp.d.layout == PCQ_LAYOUT_LAS) {
        RC(put(o, src + rn.first * fp.d.record_len, rn.n * fp.d.record_len));
        st.rec = dst + o;
        o += round_up(rn.n * fp.d.record_len, 256);
      } else {
        if (fp.need_pos) {
          RC(put(o, src + rn.first * 12, rn.n * 12));
          st.rec = dst + o;
          o += round_up(rn.n * 12, 256);
        } else {
          st.rec = dst;  // never dereferenced by a class count
        }
        if (fp.need_cls) {
          RC(put(o, src + (uint64_t)cls_offset_in_record(fp.d.format) * N + rn.first, rn.n));
          st.cls = dst + o;
          o += round_up(rn.n, 256);
        }
        if (fp.need_rgb) {
          RC(put(o, src + (uint64_t)rgb_offset_in_record(fp.d.format) * N + rn.first * 6, rn.n * 6));
          st.rgb = dst + o;
          o += round_up(rn.n * 6, 256);
        }
      }
      staged[r] = st;
    }
    if (via_bounce) {
      CU(cudaEventRecord(ctx->bounce_done[b], ctx->copy_stream));
      ctx->bounce_busy[b] = true;
    }
    CU(cudaEventRecord(ctx->chunk_copied[b], ctx->copy_stream));
    return PCQ_OK;
  };

  // An error inside the streaming loop must not leave H2D copies queued against the caller's buffers and the bounce
  // slots: the loop runs in a lambda and every exit drains both streams first.
  auto stream_all = [&]() -> int {
  // buffers start out free
  for (int b = 0; b < kChunkBuffers; ++b) CU(cudaEventRecord(ctx->chunk_free[b], ctx->stream));
  const size_t prefetch = kChunkBuffers - 1;
  for (size_t j = 0; j < std::min(prefetch, pieces.size()); ++j) RC(issue_copy(j));
  std::vector<Segment> segs;
  std::vector<uint64_t> lane_points(n_collectors, 0);
  std::vector<LaneBoxes> lane_boxes;  // per query: every file of the call on its lane
  for (uint32_t q = 0; q < n_queries; ++q) {
    lane_boxes.emplace_back(n_collectors);
    for (uint32_t i = 0; i < n_files; ++i) lane_boxes.back().add(fps[i].lane, fps[i].d);
    lane_boxes.back().cut(queries + q);
  }
  for (size_t j = 0; j < pieces.size(); ++j) {
    if (j + prefetch < pieces.size()) RC(issue_copy(j + prefetch));
    const Piece& pc = pieces[j];
    const FilePlan& fp = fps[pc.file];
    const int b = (int)(j % kChunkBuffers);
    CU(cudaStreamWaitEvent(ctx->stream, ctx->chunk_copied[b], 0));
    for (uint32_t q = 0; q < n_queries; ++q) {
      if (fp.plan[q].skip) continue;
      segs.resize(pc.n_runs);
      std::fill(lane_points.begin(), lane_points.end(), 0);
      for (uint32_t k = 0; k < pc.n_runs; ++k) {
        const Run& rn = runs[pc.run0 + k];
        const Staged& st = staged[pc.run0 + k];
        fill_segment(&segs[k], fp.d, st.rec, st.cls, st.rgb, rn.n, fp.plan[q], fp.lane, fp.base[q] + rn.first, queries[q].kind);
        lane_points[fp.lane] += rn.n;
      }
      // run_batch indexes lanes by Segment::lane, so hand it the query's full collector array
      const double fr = expected_match_fraction(fp.d, queries + q);
      RC(run_batch(ctx, segs, queries + q, collectors + (size_t)q * n_collectors, n_collectors, lane_points,
                   fr < 0.0 ? fr : std::min(1.0, fr / fp.kept_fraction), &lane_boxes[q].v));
    }
    if (fp.build_box || fp.build_cls) {
      // chunk headers as a by-product: the piece is resident anyway (unfiltered files travel as one run per piece,
      // starting on a chunk boundary)
      pcq_host_index::File& xf = hix->files[pc.file];
      for (uint32_t k = 0; k < pc.n_runs; ++k) {
        const Run& rn = runs[pc.run0 + k];
        const Staged& st = staged[pc.run0 + k];
        ChunkIndexArgs a{};
        a.rec = st.rec;
        a.cls = st.cls;
        a.n_points = rn.n;
        a.n_chunks = (rn.n + PCQ_INDEX_CHUNK_POINTS - 1) / PCQ_INDEX_CHUNK_POINTS;
        a.layout = fp.d.layout;
        a.record_len = fp.d.layout == PCQ_LAYOUT_LAS ? fp.d.record_len : 12u;
        a.cls_off = cls_offset_in_record(fp.d.format);
        a.align = field_alignment(st.rec, a.record_len);
        a.parts = (uint8_t)((fp.build_box ? kIndexPartBox : 0) | (fp.build_cls ? kIndexPartCls : 0));
        if (launch_chunk_index(a, xf.d_headers + rn.first / PCQ_INDEX_CHUNK_POINTS, ctx->sm_count, ctx->stream) != 0)
          return fail(PCQ_ERR_CUDA, "chunk index launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        ctx->launches++;
      }
    }
    CU(cudaEventRecord(ctx->chunk_free[b], ctx->stream));
  }
  return PCQ_OK;
  };
  const int stream_rc = stream_all();
  if (stream_rc != PCQ_OK) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (int b = 0; b < kChunkBuffers; ++b) ctx->bounce_busy[b] = false;
    return stream_rc;
  }
  if (hix) {
    for (uint32_t i = 0; i < n_files; ++i) {
      pcq_host_index::File& xf = hix->files[i];
      if (fps[i].build_box) xf.has_box = true, xf.unfetched |= kIndexPartBox, hix->pending = true;
      if (fps[i].build_cls) xf.has_cls = true, xf.unfetched |= kIndexPartCls, hix->pending = true;
    }
  }
  stats.segments = (uint32_t)runs.size();
  ctx->stats = stats;
  return PCQ_OK;
}
}  // namespace pcq
extern "C" {

int pcq_search_host_files(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes, const char* const* exts,
                          uint32_t n_files, const pcq_query* query, pcq_collector* const* collectors,
                          uint32_t n_collectors) {
  if (!query) return fail(PCQ_ERR_ARG, "pcq_search: null argument");
  return search_host_multi(ctx, file_bytes, n_bytes, exts, n_files, query, 1, collectors, n_collectors, nullptr, nullptr);
}

int pcq_search_host_files_multi(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes,
                                const char* const* exts, uint32_t n_files, const pcq_query* queries, uint32_t n_queries,
                                pcq_collector* const* collectors, uint32_t n_collectors_per_query) {
  if (!queries && n_queries) return fail(PCQ_ERR_ARG, "pcq_search: null argument");
  return search_host_multi(ctx, file_bytes, n_bytes, exts, n_files, queries, n_queries, collectors, n_collectors_per_query, nullptr, nullptr);
}

int pcq_host_index_create(pcq_ctx* ctx, pcq_host_index** out) {
  if (!ctx || !out) return fail(PCQ_ERR_ARG, "null argument");
  pcq_host_index* ix = new (std::nothrow) pcq_host_index();
  if (!ix) return fail(PCQ_ERR_NOMEM, "out of host memory");
  ix->ctx = ctx;
  ctx->refs++;
  *out = ix;
  return PCQ_OK;
}

void pcq_host_index_destroy(pcq_host_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->ctx->device);
  cudaStreamSynchronize(ix->ctx->stream);
  for (pcq_host_index::File& f : ix->files)
    if (f.d_headers) cudaFree(f.d_headers);
  ctx_unref(ix->ctx);
  delete ix;
}

int pcq_host_index_info(pcq_host_index* ix, uint32_t file, uint64_t* n_chunks, int* has_box, int* has_cls) {
  if (!ix) return fail(PCQ_ERR_ARG, "null index");
  const bool known = file < ix->files.size();
  if (n_chunks) *n_chunks = known ? ix->files[file].n_chunks : 0;
  if (has_box) *has_box = known && ix->files[file].has_box;
  if (has_cls) *has_cls = known && ix->files[file].has_cls;
  return PCQ_OK;
}

int pcq_search_host_files_indexed(pcq_ctx* ctx, const void* const* file_bytes, const size_t* n_bytes,
                                  const char* const* exts, uint32_t n_files, const pcq_query* queries, uint32_t n_queries,
                                  pcq_collector* const* collectors, uint32_t n_collectors_per_query, pcq_host_index* index) {
  if (!queries && n_queries) return fail(PCQ_ERR_ARG, "pcq_search: null argument");
  if (!index) return fail(PCQ_ERR_ARG, "null index");
  return search_host_multi(ctx, file_bytes, n_bytes, exts, n_files, queries, n_queries, collectors, n_collectors_per_query, index, nullptr);
}

// ---- multi-GPU density exchange --------------------------------------------------------------------

int pcq_grid_export_candidates(pcq_collector* c, uint32_t n_parts, const void** out_dev_candidates, uint64_t* counts) {
  if (!c || !out_dev_candidates || !counts || n_parts == 0) return fail(PCQ_ERR_ARG, "null argument");
  if (c->kind != PCQ_COLLECT_GRID) return fail(PCQ_ERR_ARG, "not a grid collector");
  pcq_ctx* ctx = c->ctx;
  RC(use_device(ctx));
  *out_dev_candidates = nullptr;
  for (uint32_t p = 0; p < n_parts; ++p) counts[p] = 0;
  if (!c->akeys.empty())
    return fail(PCQ_ERR_ALIASED,
                "density grid: %zu cell keys suffer key aliasing (grid_sampling.rs:62-70 vs 78-82); their result is a "
                "sequential fold in scan order, which cannot be merged across GPUs",
                c->akeys.size());
  const uint64_t n = c->cand_len;
  if (n == 0) return PCQ_OK;
  RC(grid_restore(c));
  if (ctx->part_scratch_cap < n_parts) {
    if (ctx->part_scratch) cudaFree(ctx->part_scratch);
    ctx->part_scratch = nullptr;
    CU(cudaMalloc(&ctx->part_scratch, 2ull * n_parts * sizeof(unsigned long long)));
    ctx->part_scratch_cap = n_parts;
  }
  CU(cudaMemsetAsync(ctx->part_scratch, 0, 2ull * n_parts * sizeof(unsigned long long), ctx->stream));
  GridDev g = grid_view(c);
  g.own_parts = 0;                // every locally occupied cell is exported
  RC(grid_pick_winners(c, g, n));
  if (launch_grid_emit(g, n, c->d_finlist, &c->dev->fin_count, 0, n_parts, ctx->part_scratch, nullptr, nullptr, nullptr, nullptr,
                       ctx->sm_count, ctx->stream) != 0)
    return fail(PCQ_ERR_CUDA, "launch failed");
  ctx->launches += 1;
  std::vector<unsigned long long> h(n_parts);
  CU(cudaMemcpyAsync(h.data(), ctx->part_scratch, n_parts * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  std::vector<unsigned long long> cursor(n_parts);
  uint64_t total = 0;
  for (uint32_t p = 0; p < n_parts; ++p) {
    cursor[p] = total;
    counts[p] = h[p];
    total += h[p];
  }
  if (c->export_cap < total) {
    if (c->d_export) cudaFree(c->d_export);
    c->d_export = nullptr;
    c->export_cap = 0;
    const uint64_t cap = total + total / 4 + 1024;
    CU(cudaMalloc(&c->d_export, cap * sizeof(Candidate)));
    c->export_cap = cap;
  }
  CU(cudaMemcpyAsync(ctx->part_scratch + n_parts, cursor.data(), n_parts * sizeof(unsigned long long),
                     cudaMemcpyHostToDevice, ctx->stream));
  if (launch_grid_emit(g, n, c->d_finlist, &c->dev->fin_count, 1, n_parts, ctx->part_scratch, ctx->part_scratch + n_parts, c->d_export,
                       nullptr, nullptr, ctx->sm_count, ctx->stream) != 0)
    return fail(PCQ_ERR_CUDA, "launch failed");
  ctx->launches += 1;
  RC(grid_restore(c));  // distances back
  CU(cudaStreamSynchronize(ctx->stream));
  *out_dev_candidates = c->d_export;
  return PCQ_OK;
}

int pcq_grid_import_candidates(pcq_collector* c, const void* dev_candidates, uint64_t n) {
  if (!c) return fail(PCQ_ERR_ARG, "null collector");
  if (c->kind != PCQ_COLLECT_GRID) return fail(PCQ_ERR_ARG, "not a grid collector");
  if (n == 0) return PCQ_OK;
  if (!dev_candidates) return fail(PCQ_ERR_ARG, "null candidates");
  pcq_ctx* ctx = c->ctx;
  RC(use_device(ctx));
  RC(ensure_grid_tables(&c, 1, nullptr));
  RC(grid_restore(c));
  c->final_valid = false;
  for (int attempt = 0; attempt < 8; ++attempt) {
    RC(grow_cands(c, c->cand_len + n));
    GridDev g = grid_view(c);
    if (launch_grid_import(g, static_cast<const Candidate*>(dev_candidates), n, ctx->sm_count, ctx->stream) != 0)
      return fail(PCQ_ERR_CUDA, "k_grid_import launch failed");
    ctx->launches++;
    DevBlock b;
    RC(read_devblock(c, &b));
    if (b.flags & kFlagHashFull) {
      RC(rehash_grid(c));
      continue;
    }
    c->cand_len = std::min<uint64_t>(b.cand_count, c->grid.cand_cap);
    return PCQ_OK;
  }
  return fail(PCQ_ERR_NOMEM, "density table did not converge");
}

}  // extern "C"
