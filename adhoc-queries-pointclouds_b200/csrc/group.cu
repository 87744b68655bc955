// group.cu — a group of GPUs of one box behind the C ABI (include/pcq.h, "multi-GPU").
//
// The reference has one parallelism strategy: one rayon task per file, no communication
// (run_search_parallel, query/src/main.rs:146-183).  Here files AND point ranges of files shard across the GPUs of a
// box:
//   * count queries    every member scans its ranges; the per-file counts are summed on the host (one process) or
//                      with one ncclAllReduce of n_files integers (one process per GPU) — no data-path collective;
//   * select queries   every member compacts its ranges in scan order; the per-(file, member) record streams are
//                      concatenated in member order, which is the order ONE BufferCollector would have seen;
//   * density queries  the one real exchange: every member builds a local cell table over its ranges, exports one
//                      candidate per locally occupied cell partitioned by owner = mix64(key) % holders, the parts
//                      travel with one grouped ncclSend/ncclRecv all-to-all over NVLink, and every owner folds what
//                      it receives into its own table; ties break on the GLOBAL scan index, so the union of the
//                      owners' winners is the sequential fold of grid_sampling.rs:97-102 over the whole dataset.
//                      Keys that suffer key aliasing (alias.cu) have an order-dependent result: every point of such
//                      a key is routed to the key's owner and folded there in global scan order.
//
// A group is either ONE process driving n GPUs (ncclCommInitAll; the `query --gpus N` CLI) or one process per GPU
// (ncclCommInitRank with an id the launcher distributes; bench.py under torchrun).  Both run the same code: a group
// is a list of local members plus the world size.  NCCL is resolved at run time (dlopen of libnccl.so.2), so that
// libpcq.so loads — and serves one GPU — on a box without it, and shares the copy a host framework already loaded.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: every function is resolved with dlsym

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
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

namespace {

// owner hash of a cell key: the murmur3 finaliser the kernels use (grid_math.cuh mix64)
inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// ---- NCCL, resolved at run time ----------------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return PCQ_OK;
  const char* names[] = {std::getenv("PCQ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) return fail(PCQ_ERR_CUDA, "multi-GPU groups need NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
  NcclApi a;
  a.handle = h;
  bool ok = true;
  auto sym = [&](const char* name) -> void* {
    void* p = dlsym(h, name);
    if (!p) ok = false;
    return p;
  };
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
  a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
  a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
  a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  if (!ok) {
    dlclose(h);
    return fail(PCQ_ERR_CUDA, "libnccl.so.2 lacks a symbol this library needs");
  }
  g_nccl = a;
  return PCQ_OK;
}

#define NC(call)                                                                                          \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess) return fail(PCQ_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString(r_));         \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes, cudaStream_t st) {
    if (cap >= bytes) return PCQ_OK;
    if (p) {
      cudaStreamSynchronize(st);
      cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    const size_t want = round_up(std::max<size_t>(bytes, 4096), 4096);
    if (cudaMalloc(&p, want) != cudaSuccess) {
      cudaGetLastError();
      return fail(PCQ_ERR_NOMEM, "cannot allocate %zu bytes of HBM for the exchange", want);
    }
    cap = want;
    return PCQ_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct GridParams {
  double gmin[3]{}, gmax[3]{}, cell = 0;
  bool operator==(const GridParams& o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; }
};

struct Member {
  int device = 0;
  uint32_t rank = 0;
  pcq_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  // collectors are pooled: a query takes what it needs, reset
  std::vector<pcq_collector*> pool[3];
  GridParams pool_grid;
  unsigned long long* h_counts = nullptr;  // pinned
  size_t h_counts_cap = 0;
  DevBuf d_counts, d_gather, xsend, xrecv, stage;
};

// one u64 per (pointer, slot): dst[idx[i]] = *src[i]  (per-file counts of a member, gathered without a sync per lane)
__global__ void k_gather_u64(const unsigned long long* const* src, const uint32_t* idx, uint32_t n, unsigned long long* dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = *src[i];
}

}  // namespace

struct pcq_group {
  uint32_t world = 1;
  bool one_process = true;  // every member of the world is local
  // how the parts of the density exchange travel: NCCL send / recv (default), or — one process only — peer copies
  // between the members' streams (cudaMemcpyPeerAsync; PCQ_GROUP_TRANSPORT=p2p, and always when two members share a
  // device, which is how a box with one GPU tests the whole exchange)
  bool use_nccl = true;
  std::vector<Member> local;
  // COUNT searches return before their kernels finish; their counts land in the members' pinned h_counts, which the
  // next search reuses: results still pending when another search starts are completed first
  std::vector<pcq_result*> pending;
  pcq_group_stats stats{};  // phases of the last select / density search
  // Pinned host buffers of the results are recycled: page-locking tens of megabytes costs more than the search that
  // fills them (~1 ms per MB on a virtualised host), so a released result hands its buffer back to the group.
  struct HostBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    bool in_use = false;
  };
  std::vector<HostBuf> host_pool;
  std::vector<pcq_result*> live;  // results that hold a pool buffer (orphaned if the group goes first)
};

struct pcq_dataset {
  pcq_group* g = nullptr;
  uint32_t n_files = 0;
  std::vector<uint64_t> points_per_file, file_start;  // global: scan order of run_search_sequential
  struct Piece {
    pcq_file* f = nullptr;
    uint32_t file = 0;
    bool owned = false;
  };
  std::vector<std::vector<Piece>> pieces;        // per local member, ascending file index
  std::vector<std::vector<uint32_t>> holders;    // per file: ranks that hold a range of it, ascending
};

struct pcq_result {
  int kind = 0;
  uint32_t n_lanes = 0;
  std::vector<uint64_t> counts;       // per lane, group-wide
  std::vector<uint64_t> lane_off;     // records: first record of the lane within `points` (n_lanes + 1 entries)
  uint8_t* points = nullptr;          // pinned host memory (a buffer of the group's pool), 31-byte records
  pcq_group* owner = nullptr;         // whose pool `points` came from
  uint64_t n_points_held = 0;         // records this process holds (rank 0's process: all of them)
  bool has_points = false;
  // a COUNT search returns before its kernels finish: the counts land in h_counts[count_slot ...] of the members
  pcq_group* pending_group = nullptr;
  size_t count_slot = 0;
  bool per_file = false;
  uint32_t n_files = 0;
};

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int sync_members(pcq_group* g) {
  for (Member& M : g->local) {
    RC(use_device(M.ctx));
    CU(cudaStreamSynchronize(M.ctx->stream));
  }
  return PCQ_OK;
}

int member_index_of_rank(const pcq_group* g, uint32_t rank) {
  for (size_t m = 0; m < g->local.size(); ++m)
    if (g->local[m].rank == rank) return (int)m;
  return -1;
}

// run fn(member index) for every local member, each on its own host thread (the calls synchronise their streams);
// the first error wins and its message becomes this thread's pcq_last_error
template <class Fn>
int for_each_member(pcq_group* g, Fn fn) {
  const size_t n = g->local.size();
  if (n == 1) return fn(0);
  std::vector<int> rcs(n, PCQ_OK);
  std::vector<std::string> msgs(n);
  std::vector<std::thread> pool;
  for (size_t m = 0; m < n; ++m)
    pool.emplace_back([&, m] {
      rcs[m] = fn((uint32_t)m);
      if (rcs[m] != PCQ_OK) msgs[m] = last_error();
    });
  for (std::thread& t : pool) t.join();
  for (size_t m = 0; m < n; ++m)
    if (rcs[m] != PCQ_OK) return fail(rcs[m], "GPU %d (rank %u): %s", g->local[m].device, g->local[m].rank, msgs[m].c_str());
  return PCQ_OK;
}

// Every member contributes a vector of L integers; everybody learns all of them: out[rank * L + i].
int allgather_u64(pcq_group* g, const std::vector<std::vector<uint64_t>>& mine, size_t L, std::vector<uint64_t>& out) {
  out.assign((size_t)g->world * L, 0);
  if (L == 0) return PCQ_OK;
  if (g->one_process) {
    for (size_t m = 0; m < g->local.size(); ++m) std::copy(mine[m].begin(), mine[m].end(), out.begin() + g->local[m].rank * L);
    return PCQ_OK;
  }
  for (size_t m = 0; m < g->local.size(); ++m) {
    Member& M = g->local[m];
    RC(use_device(M.ctx));
    RC(M.d_gather.ensure(((size_t)g->world + 1) * L * 8, M.ctx->stream));
    CU(cudaMemcpyAsync(M.d_gather.p, mine[m].data(), L * 8, cudaMemcpyHostToDevice, M.ctx->stream));
  }
  NC(g_nccl.GroupStart());
  for (Member& M : g->local) {
    uint64_t* base = static_cast<uint64_t*>(M.d_gather.p);
    NC(g_nccl.AllGather(base, base + L, L, ncclUint64, M.comm, M.ctx->stream));
  }
  NC(g_nccl.GroupEnd());
  Member& M0 = g->local[0];
  RC(use_device(M0.ctx));
  CU(cudaMemcpyAsync(out.data(), static_cast<uint64_t*>(M0.d_gather.p) + L, (size_t)g->world * L * 8, cudaMemcpyDeviceToHost, M0.ctx->stream));
  for (Member& M : g->local) {
    RC(use_device(M.ctx));
    CU(cudaStreamSynchronize(M.ctx->stream));
  }
  return PCQ_OK;
}

// One part for every (sender, lane, receiver): the all-to-all of the density exchange.
struct Part {
  uint32_t lane;       // global lane id (sends to one peer are matched in ascending lane order on both sides)
  uint32_t peer;       // rank
  const void* src;     // sender side: device pointer
  void* dst;           // receiver side: device pointer
  uint64_t bytes;
};

// sends[m] / recvs[m]: the parts local member m sends / receives, both in (pass, lane, holder) order, so that the k-th
// part a member sends to a peer is the k-th part that peer receives from it
int all_to_all(pcq_group* g, const std::vector<std::vector<Part>>& sends, const std::vector<std::vector<Part>>& recvs) {
  bool any = false;
  for (size_t m = 0; m < g->local.size(); ++m) any |= !sends[m].empty() || !recvs[m].empty();
  if (!any) return PCQ_OK;
  if (g->use_nccl) {
    NC(g_nccl.GroupStart());
    for (size_t m = 0; m < g->local.size(); ++m) {
      Member& M = g->local[m];
      for (const Part& p : sends[m]) NC(g_nccl.Send(p.src, p.bytes, ncclUint8, (int)p.peer, M.comm, M.ctx->stream));
      for (const Part& p : recvs[m]) NC(g_nccl.Recv(p.dst, p.bytes, ncclUint8, (int)p.peer, M.comm, M.ctx->stream));
    }
    NC(g_nccl.GroupEnd());
    return PCQ_OK;
  }
  // peer copies (one process): the receiver's stream waits for the sender's data, copies, and the sender's stream
  // waits for the copies before it may reuse its export buffers
  const size_t nl = g->local.size();
  std::vector<cudaEvent_t> ready(nl, nullptr), done(nl, nullptr);
  auto cleanup = [&] {
    for (cudaEvent_t e : ready)
      if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : done)
      if (e) cudaEventDestroy(e);
  };
  int rc = PCQ_OK;
  for (size_t m = 0; m < nl && rc == PCQ_OK; ++m) {
    Member& M = g->local[m];
    rc = use_device(M.ctx);
    if (rc != PCQ_OK) break;
    if (cudaEventCreateWithFlags(&ready[m], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&done[m], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventRecord(ready[m], M.ctx->stream) != cudaSuccess)
      rc = fail(PCQ_ERR_CUDA, "event setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  for (size_t d = 0; d < nl && rc == PCQ_OK; ++d) {
    Member& D = g->local[d];
    rc = use_device(D.ctx);
    std::vector<size_t> next(nl, 0);  // per sender: how many of its parts for D have been matched
    for (const Part& r : recvs[d]) {
      if (rc != PCQ_OK) break;
      const int sm = member_index_of_rank(g, r.peer);
      if (sm < 0) {
        rc = fail(PCQ_ERR_ARG, "internal: peer copy from a rank that is not local");
        break;
      }
      const Part* sp = nullptr;
      size_t& k = next[sm];
      while (k < sends[sm].size()) {
        const Part& c = sends[sm][k++];
        if (c.peer == D.rank) {
          sp = &c;
          break;
        }
      }
      if (!sp || sp->bytes != r.bytes || sp->lane != r.lane) {
        rc = fail(PCQ_ERR_ARG, "internal: send / receive lists of the exchange do not match");
        break;
      }
      if (cudaStreamWaitEvent(D.ctx->stream, ready[sm], 0) != cudaSuccess ||
          cudaMemcpyPeerAsync(r.dst, D.device, sp->src, g->local[sm].device, r.bytes, D.ctx->stream) != cudaSuccess)
        rc = fail(PCQ_ERR_CUDA, "peer copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == PCQ_OK && cudaEventRecord(done[d], D.ctx->stream) != cudaSuccess) rc = fail(PCQ_ERR_CUDA, "cudaEventRecord failed");
  }
  for (size_t m = 0; m < nl && rc == PCQ_OK; ++m) {
    rc = use_device(g->local[m].ctx);
    for (size_t d = 0; d < nl && rc == PCQ_OK; ++d)
      if (d != m && cudaStreamWaitEvent(g->local[m].ctx->stream, done[d], 0) != cudaSuccess) rc = fail(PCQ_ERR_CUDA, "cudaStreamWaitEvent failed");
  }
  if (rc != PCQ_OK) {
    const std::string msg = last_error();
    for (Member& M : g->local) {
      cudaSetDevice(M.device);
      cudaStreamSynchronize(M.ctx->stream);
    }
    cleanup();
    return fail(rc, "%s", msg.c_str());
  }
  cleanup();  // (destroying a recorded event is legal: its resources go when the work completes)
  return PCQ_OK;
}

int take_collectors(Member& M, int kind, size_t n, const GridParams& gp, std::vector<pcq_collector*>& out) {
  std::vector<pcq_collector*>& pool = M.pool[kind];
  if (kind == PCQ_COLLECT_GRID && !(M.pool_grid == gp)) {
    for (pcq_collector* c : pool) pcq_collector_destroy(c);
    pool.clear();
    M.pool_grid = gp;
  }
  while (pool.size() < n) {
    pcq_collector* c = nullptr;
    RC(pcq_collector_create(M.ctx, kind, gp.gmin, gp.gmax, gp.cell, &c));
    pool.push_back(c);
  }
  out.assign(pool.begin(), pool.begin() + n);
  if (kind == PCQ_COLLECT_GRID) {
    for (pcq_collector* c : out) RC(pcq_collector_reset(c));
    return PCQ_OK;
  }
  return reset_collectors_batched(M.ctx, out.data(), out.size());
}

int ensure_h_counts(Member& M, size_t n) {
  if (M.h_counts_cap >= n) return PCQ_OK;
  if (M.h_counts) cudaFreeHost(M.h_counts);
  M.h_counts = nullptr;
  M.h_counts_cap = 0;
  const size_t cap = std::max<size_t>(n, 256);
  if (cudaMallocHost(reinterpret_cast<void**>(&M.h_counts), cap * 8) != cudaSuccess) {
    cudaGetLastError();
    return fail(PCQ_ERR_NOMEM, "cannot pin %zu bytes of host memory", cap * 8);
  }
  M.h_counts_cap = cap;
  return PCQ_OK;
}

// What one member contributes to one query: its internal lanes (one per piece, or one for a sequential grid)
struct MemberRun {
  std::vector<pcq_collector*> cols;
  std::vector<uint32_t> lane_of;  // global lane id of each collector (file index, or 0)
};

// ---- COUNT: per-file counts of every member -> group totals, without a host sync per lane --------------------------
// runs[q][m]; the counts of query q land in slot q * n_lanes of the members' count arrays
int combine_counts(pcq_group* g, uint32_t n_lanes, const std::vector<std::vector<MemberRun>>& runs, pcq_result* const* res) {
  const size_t nq = runs.size(), total = nq * n_lanes;
  for (size_t m = 0; m < g->local.size(); ++m) {
    Member& M = g->local[m];
    RC(use_device(M.ctx));
    RC(M.d_counts.ensure(total * 8, M.ctx->stream));
    RC(ensure_h_counts(M, total));
    CU(cudaMemsetAsync(M.d_counts.p, 0, total * 8, M.ctx->stream));
    uint32_t n = 0;
    for (size_t q = 0; q < nq; ++q) n += (uint32_t)runs[q][m].cols.size();
    if (n) {
      // pointer table + slot table in one upload
      std::vector<uint64_t> tab(n + (n + 1) / 2);
      uint32_t* idx = reinterpret_cast<uint32_t*>(tab.data() + n);
      uint32_t k = 0;
      for (size_t q = 0; q < nq; ++q)
        for (size_t i = 0; i < runs[q][m].cols.size(); ++i, ++k) {
          tab[k] = reinterpret_cast<uint64_t>(&runs[q][m].cols[i]->dev->count);
          idx[k] = (uint32_t)(q * n_lanes) + runs[q][m].lane_of[i];
        }
      void* d_tab = nullptr;
      RC(upload(M.ctx, tab.data(), tab.size() * 8, &d_tab));
      k_gather_u64<<<(n + 127) / 128, 128, 0, M.ctx->stream>>>(static_cast<const unsigned long long* const*>(d_tab),
                                                                reinterpret_cast<const uint32_t*>(static_cast<uint64_t*>(d_tab) + n), n,
                                                                static_cast<unsigned long long*>(M.d_counts.p));
      if (cudaGetLastError() != cudaSuccess) return fail(PCQ_ERR_CUDA, "k_gather_u64 launch failed");
      M.ctx->launches++;
    }
  }
  if (!g->one_process) {
    NC(g_nccl.GroupStart());
    for (Member& M : g->local) NC(g_nccl.AllReduce(M.d_counts.p, M.d_counts.p, total, ncclUint64, ncclSum, M.comm, M.ctx->stream));
    NC(g_nccl.GroupEnd());
  }
  // one process: every member's counts come to the host and are summed there (pcq_result_counts)
  for (size_t m = 0; m < g->local.size(); ++m) {
    Member& M = g->local[m];
    if (!g->one_process && m != 0) break;
    RC(use_device(M.ctx));
    CU(cudaMemcpyAsync(M.h_counts, M.d_counts.p, total * 8, cudaMemcpyDeviceToHost, M.ctx->stream));
  }
  for (size_t q = 0; q < nq; ++q) {
    res[q]->pending_group = g;
    res[q]->count_slot = q * n_lanes;
    g->pending.push_back(res[q]);
  }
  return PCQ_OK;
}

int finish_counts(pcq_result* res) {
  pcq_group* g = res->pending_group;
  if (!g) return PCQ_OK;
  res->pending_group = nullptr;
  g->pending.erase(std::remove(g->pending.begin(), g->pending.end(), res), g->pending.end());
  const uint32_t n_int = res->n_files;  // internal lanes of a COUNT search are files
  std::vector<uint64_t> tot(n_int, 0);
  for (size_t m = 0; m < g->local.size(); ++m) {
    Member& M = g->local[m];
    if (!g->one_process && m != 0) break;
    RC(use_device(M.ctx));
    CU(cudaStreamSynchronize(M.ctx->stream));
    for (uint32_t i = 0; i < n_int; ++i) tot[i] += M.h_counts[res->count_slot + i];
  }
  if (res->per_file) {
    res->counts = tot;
  } else {
    uint64_t s = 0;
    for (uint64_t v : tot) s += v;
    res->counts.assign(1, s);
  }
  return PCQ_OK;
}

int finish_pending(pcq_group* g) {
  while (!g->pending.empty()) RC(finish_counts(g->pending.back()));
  return PCQ_OK;
}

// pinned host buffer for a result: the smallest free pool buffer that fits, else a new one (a few spares are kept)
int pool_acquire(pcq_group* g, size_t bytes, pcq_result* res) {
  int best = -1;
  for (size_t i = 0; i < g->host_pool.size(); ++i) {
    const pcq_group::HostBuf& b = g->host_pool[i];
    if (!b.in_use && b.cap >= bytes && (best < 0 || b.cap < g->host_pool[best].cap)) best = (int)i;
  }
  if (best < 0) {
    // drop free buffers that are too small before pinning more
    for (pcq_group::HostBuf& b : g->host_pool)
      if (!b.in_use && b.p) {
        cudaFreeHost(b.p);
        b.p = nullptr;
        b.cap = 0;
      }
    pcq_group::HostBuf nb;
    nb.cap = round_up(bytes + bytes / 8, 1u << 20);
    if (cudaMallocHost(reinterpret_cast<void**>(&nb.p), nb.cap) != cudaSuccess) {
      cudaGetLastError();
      return fail(PCQ_ERR_NOMEM, "cannot pin %zu bytes of host memory for the result", nb.cap);
    }
    size_t slot = g->host_pool.size();
    for (size_t i = 0; i < g->host_pool.size(); ++i)
      if (!g->host_pool[i].p) slot = i;
    if (slot == g->host_pool.size()) g->host_pool.push_back(nb);
    else g->host_pool[slot] = nb;
    best = (int)slot;
  }
  g->host_pool[best].in_use = true;
  res->points = g->host_pool[best].p;
  res->owner = g;
  g->live.push_back(res);
  return PCQ_OK;
}

void pool_release(pcq_result* res) {
  pcq_group* g = res->owner;
  if (!g || !res->points) return;
  for (pcq_group::HostBuf& b : g->host_pool)
    if (b.p == res->points) b.in_use = false;
  g->live.erase(std::remove(g->live.begin(), g->live.end(), res), g->live.end());
  res->points = nullptr;
  res->owner = nullptr;
}

// ---- records of every (member, lane) -> host lanes in member order ---------------------------------------------------
struct Stream {  // records one local member holds for one lane, in HBM
  uint32_t lane;
  const uint8_t* dev;
  uint64_t n;
};

// streams[m]: ascending lane.  n_int internal lanes; the result has n_int lanes (per_file) or one (their concatenation).
int gather_records(pcq_group* g, uint32_t n_int, bool per_file, const std::vector<std::vector<Stream>>& streams, pcq_result* res) {
  const uint32_t W = g->world;
  std::vector<std::vector<uint64_t>> mine(g->local.size(), std::vector<uint64_t>(n_int, 0));
  for (size_t m = 0; m < g->local.size(); ++m)
    for (const Stream& s : streams[m]) mine[m][s.lane] += s.n;
  std::vector<uint64_t> all;
  RC(allgather_u64(g, mine, n_int, all));  // all[rank * n_int + lane]
  // layout of the host buffer: lane-major, inside a lane by rank
  std::vector<uint64_t> lane_off(n_int + 1, 0);
  for (uint32_t l = 0; l < n_int; ++l) {
    uint64_t t = 0;
    for (uint32_t r = 0; r < W; ++r) t += all[(size_t)r * n_int + l];
    lane_off[l + 1] = lane_off[l] + t;
  }
  const uint64_t total = lane_off[n_int];
  if (per_file) {
    res->n_lanes = n_int;
    res->counts.resize(n_int);
    for (uint32_t l = 0; l < n_int; ++l) res->counts[l] = lane_off[l + 1] - lane_off[l];
    res->lane_off = lane_off;
  } else {
    res->n_lanes = 1;
    res->counts.assign(1, total);
    res->lane_off = {0, total};
  }
  res->has_points = true;
  const bool root_here = member_index_of_rank(g, 0) >= 0;
  if (!root_here) {  // one process per GPU, not rank 0: the records go to rank 0
    res->lane_off.assign(res->n_lanes + 1, 0);
    res->n_points_held = 0;
  } else {
    res->n_points_held = total;
    if (total) RC(pool_acquire(g, total * 31ull, res));
  }
  auto place = [&](uint32_t lane, uint32_t rank) -> uint64_t {  // first record of (lane, rank) in the host buffer
    uint64_t o = lane_off[lane];
    for (uint32_t r = 0; r < rank; ++r) o += all[(size_t)r * n_int + lane];
    return o;
  };
  // local members copy straight to the host (every GPU has its own PCIe link)
  if (root_here) {
    for (size_t m = 0; m < g->local.size(); ++m) {
      Member& M = g->local[m];
      RC(use_device(M.ctx));
      for (const Stream& s : streams[m])
        if (s.n) CU(cudaMemcpyAsync(res->points + place(s.lane, M.rank) * 31ull, s.dev, s.n * 31ull, cudaMemcpyDeviceToHost, M.ctx->stream));
    }
  }
  // remote members (one process per GPU) send theirs to rank 0 over NVLink, which lands them in HBM and copies down
  if (!g->one_process) {
    std::vector<std::vector<Part>> sends(g->local.size()), recvs(g->local.size());
    const int root = member_index_of_rank(g, 0);
    uint64_t stage_bytes = 0;
    if (root >= 0)
      for (uint32_t l = 0; l < n_int; ++l)
        for (uint32_t r = 0; r < W; ++r)
          if (member_index_of_rank(g, r) < 0) stage_bytes += round_up(all[(size_t)r * n_int + l] * 31ull, 16);
    if (root >= 0) {
      Member& R = g->local[root];
      RC(use_device(R.ctx));
      RC(R.stage.ensure(stage_bytes, R.ctx->stream));
    }
    std::vector<std::pair<uint64_t, std::pair<uint64_t, uint64_t>>> landed;  // (stage offset, (host record, n))
    uint64_t so = 0;
    for (uint32_t l = 0; l < n_int; ++l)
      for (uint32_t r = 0; r < W; ++r) {
        const uint64_t n = all[(size_t)r * n_int + l];
        if (n == 0) continue;
        const int lm = member_index_of_rank(g, r);
        if (root >= 0 && lm < 0) {
          recvs[root].push_back({l, r, nullptr, static_cast<uint8_t*>(g->local[root].stage.p) + so, n * 31ull});
          landed.push_back({so, {place(l, r), n}});
          so += round_up(n * 31ull, 16);
        } else if (root < 0 && lm >= 0) {
          for (const Stream& s : streams[lm])
            if (s.lane == l && s.n) sends[lm].push_back({l, 0u, s.dev, nullptr, s.n * 31ull});
        }
      }
    RC(all_to_all(g, sends, recvs));
    if (root >= 0) {
      Member& R = g->local[root];
      RC(use_device(R.ctx));
      for (const auto& it : landed)
        CU(cudaMemcpyAsync(res->points + it.second.first * 31ull, static_cast<uint8_t*>(R.stage.p) + it.first, it.second.second * 31ull,
                           cudaMemcpyDeviceToHost, R.ctx->stream));
    }
  }
  for (Member& M : g->local) {
    RC(use_device(M.ctx));
    CU(cudaStreamSynchronize(M.ctx->stream));
  }
  return PCQ_OK;
}

// ---- GRID: the exchange ------------------------------------------------------------------------------------------------

// rank index of `rank` among the holders of a lane, or -1
int holder_index(const std::vector<uint32_t>& holders, uint32_t rank) {
  for (size_t i = 0; i < holders.size(); ++i)
    if (holders[i] == rank) return (int)i;
  return -1;
}

// `rerun(member)`: run the member's search of this query again (the collectors are in log-only mode by then)
template <class Rerun>
int combine_grid(pcq_group* g, uint32_t n_int, bool per_file, const std::vector<std::vector<uint32_t>>& holders,
                 std::vector<MemberRun>& runs, Rerun rerun, pcq_result* res) {
  const uint32_t W = g->world;
  const size_t nl = g->local.size();
  std::vector<std::vector<Stream>> streams(nl);
  pcq_group_stats& st = g->stats;
  double t0 = now_ms();
  if (W > 1) {
    if (g->use_nccl) RC(load_nccl());
    // 1. affected keys (key aliasing, alias.cu) of every lane, group-wide
    std::vector<std::vector<uint64_t>> n_keys(nl, std::vector<uint64_t>(1, 0));
    for (size_t m = 0; m < nl; ++m)
      for (pcq_collector* c : runs[m].cols) n_keys[m][0] += c->akeys.size();
    std::vector<uint64_t> all_n;
    RC(allgather_u64(g, n_keys, 1, all_n));
    uint64_t max_keys = 0;
    for (uint64_t v : all_n) max_keys = std::max(max_keys, v);
    std::vector<std::vector<uint64_t>> lane_keys(n_int);  // sorted, unique
    if (max_keys) {
      std::vector<std::vector<uint64_t>> pairs(nl, std::vector<uint64_t>(2 * max_keys, ~0ull));
      for (size_t m = 0; m < nl; ++m) {
        size_t k = 0;
        for (size_t i = 0; i < runs[m].cols.size(); ++i)
          for (uint64_t key : runs[m].cols[i]->akeys) {
            pairs[m][2 * k] = runs[m].lane_of[i];
            pairs[m][2 * k + 1] = key;
            ++k;
          }
      }
      std::vector<uint64_t> all_pairs;
      RC(allgather_u64(g, pairs, 2 * max_keys, all_pairs));
      for (size_t i = 0; i + 1 < all_pairs.size(); i += 2)
        if (all_pairs[i] != ~0ull && all_pairs[i] < n_int) lane_keys[all_pairs[i]].push_back(all_pairs[i + 1]);
      for (std::vector<uint64_t>& v : lane_keys) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
      }
      // 2. every member collects ALL its points of the affected keys of its lanes: a second, log-only pass.  The fold
      //    of such a key depends on the order of all its points, so local winners are not enough.
      for (size_t m = 0; m < nl; ++m)
        for (size_t i = 0; i < runs[m].cols.size(); ++i) {
          pcq_collector* c = runs[m].cols[i];
          const std::vector<uint64_t>& keys = lane_keys[runs[m].lane_of[i]];
          c->akeys = keys;
          c->astates.assign(keys.size(), AliasState{});
          c->rawlog_len = 0;
          c->pass_mode = keys.empty() ? 2 : 1;  // the second pass only logs; lanes without affected keys sit it out
          c->final_valid = false;
          RC(use_device(c->ctx));
          RC(alias_upload(c));
        }
      const int rc = for_each_member(g, [&](uint32_t m) -> int {
        bool any = false;
        for (pcq_collector* c : runs[m].cols) any |= c->pass_mode == 1;
        return any ? rerun(m) : PCQ_OK;
      });
      for (size_t m = 0; m < nl; ++m)
        for (pcq_collector* c : runs[m].cols) c->pass_mode = 0;
      RC(rc);
      for (const std::vector<uint64_t>& v : lane_keys) st.affected_keys += v.size();
    }
    st.rescan_ms += now_ms() - t0;
    t0 = now_ms();

    // 3. export: one candidate per locally occupied cell, partitioned by owner among the holders of the lane
    struct Exp {
      const uint8_t* dev = nullptr;
      std::vector<uint64_t> counts;      // per holder
      std::vector<Candidate> log_parts;  // raw log of affected keys, bucketed by owner (host)
      std::vector<uint64_t> log_counts;  // per holder
    };
    std::vector<std::vector<Exp>> exps(nl);
    for (size_t m = 0; m < nl; ++m) {
      exps[m].resize(runs[m].cols.size());
      for (size_t i = 0; i < runs[m].cols.size(); ++i) {
        pcq_collector* c = runs[m].cols[i];
        const std::vector<uint32_t>& H = holders[runs[m].lane_of[i]];
        Exp& e = exps[m][i];
        e.counts.assign(H.size(), 0);
        e.log_counts.assign(H.size(), 0);
        const void* d = nullptr;
        // (the set of affected keys is uploaded, so the export leaves them out: their points travel as raw log)
        std::vector<uint64_t> saved;
        saved.swap(c->akeys);  // pcq_grid_export_candidates refuses collectors with affected keys; the device-side set stays
        const int rc = pcq_grid_export_candidates(c, (uint32_t)H.size(), &d, e.counts.data());
        saved.swap(c->akeys);
        RC(rc);
        e.dev = static_cast<const uint8_t*>(d);
        if (c->rawlog_len) {
          std::vector<Candidate> raw(c->rawlog_len);
          RC(use_device(c->ctx));
          CU(cudaMemcpy(raw.data(), c->d_rawlog, raw.size() * sizeof(Candidate), cudaMemcpyDeviceToHost));
          for (const Candidate& x : raw) e.log_counts[mix64(x.key) % H.size()]++;
          std::vector<uint64_t> cur(H.size(), 0);
          for (size_t h = 1; h < H.size(); ++h) cur[h] = cur[h - 1] + e.log_counts[h - 1];
          e.log_parts.resize(raw.size());
          for (const Candidate& x : raw) e.log_parts[cur[mix64(x.key) % H.size()]++] = x;
        }
      }
    }
    st.export_ms += now_ms() - t0;
    t0 = now_ms();
    // 4. everybody learns every part size: sizes[(rank * n_int + lane) * 2W + 2 * owner rank + {0: cells, 1: log}]
    const size_t L = (size_t)n_int * 2 * W;
    std::vector<std::vector<uint64_t>> mine(nl, std::vector<uint64_t>(L, 0));
    for (size_t m = 0; m < nl; ++m)
      for (size_t i = 0; i < runs[m].cols.size(); ++i) {
        const uint32_t lane = runs[m].lane_of[i];
        const std::vector<uint32_t>& H = holders[lane];
        for (size_t h = 0; h < H.size(); ++h) {
          mine[m][((size_t)lane * W + H[h]) * 2 + 0] = exps[m][i].counts[h];
          mine[m][((size_t)lane * W + H[h]) * 2 + 1] = exps[m][i].log_counts[h];
        }
      }
    std::vector<uint64_t> sizes;
    RC(allgather_u64(g, mine, L, sizes));
    auto size_of = [&](uint32_t from, uint32_t lane, uint32_t to, int what) -> uint64_t {
      return sizes[(size_t)from * L + ((size_t)lane * W + to) * 2 + what];
    };
    // 5. the all-to-all: candidates, then raw log entries, 64 bytes each, over NCCL send / recv
    std::vector<std::vector<Part>> sends(nl), recvs(nl);
    std::vector<std::vector<uint64_t>> recv_cells(nl), recv_logs(nl);  // per collector: entries received
    std::vector<std::vector<uint64_t>> recv_cell_off(nl), recv_log_off(nl), send_log_off(nl);
    for (size_t m = 0; m < nl; ++m) {
      Member& M = g->local[m];
      RC(use_device(M.ctx));
      const size_t nc = runs[m].cols.size();
      recv_cells[m].assign(nc, 0);
      recv_logs[m].assign(nc, 0);
      recv_cell_off[m].assign(nc, 0);
      recv_log_off[m].assign(nc, 0);
      send_log_off[m].assign(nc, 0);
      uint64_t rbytes = 0, sbytes = 0;
      for (size_t i = 0; i < nc; ++i) {
        const uint32_t lane = runs[m].lane_of[i];
        for (uint32_t r : holders[lane])
          if (r != M.rank) {
            recv_cells[m][i] += size_of(r, lane, M.rank, 0);
            recv_logs[m][i] += size_of(r, lane, M.rank, 1);
          }
        recv_cell_off[m][i] = rbytes;
        rbytes += recv_cells[m][i] * sizeof(Candidate);
        recv_log_off[m][i] = rbytes;
        rbytes += recv_logs[m][i] * sizeof(Candidate);
        send_log_off[m][i] = sbytes;
        sbytes += exps[m][i].log_parts.size() * sizeof(Candidate);
      }
      RC(M.xrecv.ensure(rbytes, M.ctx->stream));
      RC(M.xsend.ensure(sbytes, M.ctx->stream));
      for (size_t i = 0; i < nc; ++i)
        if (!exps[m][i].log_parts.empty())
          CU(cudaMemcpyAsync(static_cast<uint8_t*>(M.xsend.p) + send_log_off[m][i], exps[m][i].log_parts.data(),
                             exps[m][i].log_parts.size() * sizeof(Candidate), cudaMemcpyHostToDevice, M.ctx->stream));
    }
    for (int what = 0; what < 2; ++what)
      for (size_t m = 0; m < nl; ++m) {
        Member& M = g->local[m];
        for (size_t i = 0; i < runs[m].cols.size(); ++i) {
          const uint32_t lane = runs[m].lane_of[i];
          const std::vector<uint32_t>& H = holders[lane];
          uint64_t so = 0, ro = 0;
          for (size_t h = 0; h < H.size(); ++h) {
            const uint32_t r = H[h];
            const uint64_t ns = what == 0 ? exps[m][i].counts[h] : exps[m][i].log_counts[h];
            if (r != M.rank && ns) {
              const uint8_t* base = what == 0 ? exps[m][i].dev : static_cast<const uint8_t*>(M.xsend.p) + send_log_off[m][i];
              sends[m].push_back({lane, r, base + so * sizeof(Candidate), nullptr, ns * sizeof(Candidate)});
            }
            so += ns;
            if (r != M.rank) {
              const uint64_t nr = size_of(r, lane, M.rank, what);
              if (nr) {
                uint8_t* base = static_cast<uint8_t*>(M.xrecv.p) + (what == 0 ? recv_cell_off[m][i] : recv_log_off[m][i]);
                recvs[m].push_back({lane, r, nullptr, base + ro * sizeof(Candidate), nr * sizeof(Candidate)});
                ro += nr;
              }
            }
          }
        }
      }
    RC(all_to_all(g, sends, recvs));
    RC(sync_members(g));
    for (size_t m = 0; m < nl; ++m)
      for (const Part& p : sends[m]) {
        st.bytes_sent += p.bytes;
      }
    for (size_t m = 0; m < nl; ++m)
      for (size_t i = 0; i < runs[m].cols.size(); ++i) {
        const std::vector<uint32_t>& H = holders[runs[m].lane_of[i]];
        for (size_t h = 0; h < H.size(); ++h)
          if (H[h] != g->local[m].rank) {
            st.cells_sent += exps[m][i].counts[h];
            st.log_entries_sent += exps[m][i].log_counts[h];
          }
      }
    st.exchange_ms += now_ms() - t0;
    t0 = now_ms();
    // 6. owners fold what they received: cells into their own table (their own part is already there), raw log
    //    entries — together with their own — through the ordered replay
    RC(for_each_member(g, [&](uint32_t m) -> int {
      Member& M = g->local[m];
      RC(use_device(M.ctx));
      for (size_t i = 0; i < runs[m].cols.size(); ++i) {
        pcq_collector* c = runs[m].cols[i];
        const uint32_t lane = runs[m].lane_of[i];
        const std::vector<uint32_t>& H = holders[lane];
        const int me = holder_index(H, M.rank);
        if (recv_cells[m][i])
          RC(pcq_grid_import_candidates(c, static_cast<uint8_t*>(M.xrecv.p) + recv_cell_off[m][i], recv_cells[m][i]));
        uint64_t own_log = 0, own_off = 0;
        if (me >= 0 && !exps[m][i].log_counts.empty()) {
          own_log = exps[m][i].log_counts[me];
          for (int h = 0; h < me; ++h) own_off += exps[m][i].log_counts[h];
        }
        const uint64_t n_log = own_log + recv_logs[m][i];
        if (n_log) {
          RC(grow_log(c, n_log));
          if (own_log)
            CU(cudaMemcpyAsync(c->d_log, static_cast<uint8_t*>(M.xsend.p) + send_log_off[m][i] + own_off * sizeof(Candidate),
                               own_log * sizeof(Candidate), cudaMemcpyDeviceToDevice, M.ctx->stream));
          if (recv_logs[m][i])
            CU(cudaMemcpyAsync(c->d_log + own_log, static_cast<uint8_t*>(M.xrecv.p) + recv_log_off[m][i],
                               recv_logs[m][i] * sizeof(Candidate), cudaMemcpyDeviceToDevice, M.ctx->stream));
          GridDev gv = grid_view(c);
          const int rc = alias_replay(gv, n_log, c->d_astates, M.ctx->sm_count, M.ctx->stream);
          if (rc != 0) return fail(rc == -2 ? PCQ_ERR_NOMEM : PCQ_ERR_CUDA, "alias replay of %llu points failed", (unsigned long long)n_log);
          M.ctx->launches += 5;
          CU(cudaMemcpy(c->astates.data(), c->d_astates, c->astates.size() * sizeof(AliasState), cudaMemcpyDeviceToHost));
        }
        c->own_parts = (uint32_t)H.size();
        c->own_me = me >= 0 ? (uint32_t)me : 0u;
        c->final_valid = false;
      }
      return PCQ_OK;
    }));
  }
  if (W > 1) st.import_ms += now_ms() - t0;
  t0 = now_ms();
  // 7. every owner's winners -> the result
  RC(for_each_member(g, [&](uint32_t m) -> int {
    for (size_t i = 0; i < runs[m].cols.size(); ++i) {
      const void* d = nullptr;
      uint64_t n = 0;
      RC(pcq_collector_points_device(runs[m].cols[i], &d, &n));
      streams[m].push_back({runs[m].lane_of[i], static_cast<const uint8_t*>(d), n});
    }
    return PCQ_OK;
  }));
  const int grc = gather_records(g, n_int, per_file, streams, res);
  st.finalize_ms += now_ms() - t0;
  return grc;
}

int check_query_args(pcq_group* g, const pcq_query* q, uint32_t n_queries, int kind, const double* gmin, const double* gmax) {
  if (!g || (!q && n_queries)) return fail(PCQ_ERR_ARG, "pcq_group_search: null argument");
  if (kind < PCQ_COLLECT_COUNT || kind > PCQ_COLLECT_GRID) return fail(PCQ_ERR_ARG, "bad collector kind %d", kind);
  if (kind == PCQ_COLLECT_GRID && (!gmin || !gmax)) return fail(PCQ_ERR_ARG, "grid collector needs bounds");
  for (uint32_t k = 0; k < n_queries; ++k)
    if (q[k].kind == PCQ_QUERY_BOUNDS)
      for (int i = 0; i < 3; ++i)
        if (q[k].qmin[i] > q[k].qmax[i]) return fail(PCQ_ERR_PANIC, "AABB::from_min_max: query bounds have min > max on axis %d", i);
  return PCQ_OK;
}

GridParams grid_params_of(int kind, const double* gmin, const double* gmax, double cell) {
  GridParams gp;
  if (kind == PCQ_COLLECT_GRID)
    for (int i = 0; i < 3; ++i) {
      gp.gmin[i] = gmin[i];
      gp.gmax[i] = gmax[i];
    }
  gp.cell = kind == PCQ_COLLECT_GRID ? cell : 0.0;
  return gp;
}

}  // namespace

// =====================================================================================================================
extern "C" {

int pcq_shard_plan(const uint64_t* points_per_file, uint32_t n_files, uint32_t world, int mode, pcq_shard* out, uint64_t cap,
                   uint64_t* n_out) {
  if ((!points_per_file && n_files) || !n_out || (!out && cap) || world == 0) return fail(PCQ_ERR_ARG, "pcq_shard_plan: bad argument");
  if (mode != PCQ_SHARD_RANGES && mode != PCQ_SHARD_FILES) return fail(PCQ_ERR_ARG, "unknown shard mode %d", mode);
  std::vector<pcq_shard> plan;
  if (mode == PCQ_SHARD_RANGES) {
    // every file is cut into `world` contiguous ranges of whole index chunks; rank r takes the r-th range of every file,
    // so that any query — also one that touches a few files only — spreads over all GPUs
    for (uint32_t f = 0; f < n_files; ++f) {
      const uint64_t N = points_per_file[f];
      const uint64_t chunks = (N + PCQ_INDEX_CHUNK_POINTS - 1) / PCQ_INDEX_CHUNK_POINTS;
      for (uint32_t r = 0; r < world; ++r) {
        const uint64_t c0 = chunks * r / world, c1 = chunks * (r + 1) / world;
        const uint64_t p0 = c0 * PCQ_INDEX_CHUNK_POINTS, p1 = std::min<uint64_t>(c1 * PCQ_INDEX_CHUNK_POINTS, N);
        if (p1 > p0) plan.push_back({f, r, p0, p1 - p0});
      }
    }
  } else {
    // whole files, largest first onto the least loaded rank (the reference's unit of parallelism, main.rs:153-161)
    std::vector<uint32_t> order(n_files);
    for (uint32_t f = 0; f < n_files; ++f) order[f] = f;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return points_per_file[a] > points_per_file[b]; });
    std::vector<uint64_t> load(world, 0);
    std::vector<uint32_t> owner(n_files, 0);
    for (uint32_t f : order) {
      uint32_t best = 0;
      for (uint32_t r = 1; r < world; ++r)
        if (load[r] < load[best]) best = r;
      owner[f] = best;
      load[best] += points_per_file[f];
    }
    for (uint32_t f = 0; f < n_files; ++f)
      if (points_per_file[f]) plan.push_back({f, owner[f], 0, points_per_file[f]});
  }
  *n_out = plan.size();
  for (size_t i = 0; i < plan.size() && i < cap; ++i) out[i] = plan[i];
  return PCQ_OK;
}

static int group_finish_create(pcq_group* g) {
  for (Member& M : g->local) {
    RC(pcq_ctx_create(M.device, &M.ctx));
  }
  return PCQ_OK;
}

void pcq_group_destroy(pcq_group* g) {
  if (!g) return;
  finish_pending(g);
  // results that outlive their group keep their counts and lose their records
  for (pcq_result* r : g->live) {
    r->points = nullptr;
    r->owner = nullptr;
    r->has_points = false;
    r->lane_off.assign(r->n_lanes + 1, 0);
  }
  g->live.clear();
  for (pcq_group::HostBuf& b : g->host_pool)
    if (b.p) cudaFreeHost(b.p);
  for (Member& M : g->local) {
    if (M.ctx) {
      cudaSetDevice(M.device);
      cudaStreamSynchronize(M.ctx->stream);
    }
    for (int k = 0; k < 3; ++k)
      for (pcq_collector* c : M.pool[k]) pcq_collector_destroy(c);
    if (M.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(M.comm);
    if (M.h_counts) cudaFreeHost(M.h_counts);
    M.d_counts.release();
    M.d_gather.release();
    M.xsend.release();
    M.xrecv.release();
    M.stage.release();
    if (M.ctx) pcq_ctx_destroy(M.ctx);
  }
  delete g;
}

int pcq_group_create(const int* devices, uint32_t n_devices, pcq_group** out) {
  if (!out || n_devices == 0) return fail(PCQ_ERR_ARG, "pcq_group_create: bad argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(PCQ_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  }
  if ((int)n_devices > n_dev && !devices && !std::getenv("PCQ_GROUP_DEVICES"))
    return fail(PCQ_ERR_ARG, "%u GPUs asked for, %d present", n_devices, n_dev);
  pcq_group* g = new (std::nothrow) pcq_group();
  if (!g) return fail(PCQ_ERR_NOMEM, "out of host memory");
  g->world = n_devices;
  g->one_process = true;
  g->local.resize(n_devices);
  std::vector<int> devs(n_devices);
  // PCQ_GROUP_DEVICES="0,0,1": which device every member of a default group uses (a box with fewer GPUs than members
  // lets members share a device; the exchange then travels as peer copies)
  std::vector<int> env_devs;
  if (!devices)
    if (const char* e = std::getenv("PCQ_GROUP_DEVICES"))
      for (const char* c = e; *c;) {
        env_devs.push_back(std::atoi(c));
        while (*c && *c != ',') ++c;
        if (*c == ',') ++c;
      }
  for (uint32_t i = 0; i < n_devices; ++i) {
    devs[i] = devices ? devices[i] : (i < env_devs.size() ? env_devs[i] : (int)i);
    g->local[i].device = devs[i];
    g->local[i].rank = i;
  }
  bool distinct = true;
  for (uint32_t i = 0; i < n_devices; ++i)
    for (uint32_t j = 0; j < i; ++j) distinct &= devs[i] != devs[j];
  const char* tr = std::getenv("PCQ_GROUP_TRANSPORT");
  g->use_nccl = n_devices > 1 && distinct && !(tr && std::strcmp(tr, "p2p") == 0);
  int rc = group_finish_create(g);
  if (rc == PCQ_OK && !g->use_nccl && n_devices > 1) {
    // peer copies: let the GPUs reach each other's memory directly where the box allows it (NVLink / PCIe P2P);
    // cudaMemcpyPeerAsync stages through the host otherwise
    for (uint32_t i = 0; i < n_devices; ++i)
      for (uint32_t j = 0; j < n_devices; ++j) {
        int can = 0;
        if (devs[i] == devs[j] || cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can) continue;
        cudaSetDevice(devs[i]);
        if (cudaDeviceEnablePeerAccess(devs[j], 0) != cudaSuccess) cudaGetLastError();  // (already enabled: fine)
      }
  }
  if (rc == PCQ_OK && g->use_nccl) {
    rc = load_nccl();
    if (rc == PCQ_OK) {
      std::vector<ncclComm_t> comms(n_devices);
      const ncclResult_t r = g_nccl.CommInitAll(comms.data(), (int)n_devices, devs.data());
      if (r != ncclSuccess)
        rc = fail(PCQ_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(r));
      else
        for (uint32_t i = 0; i < n_devices; ++i) g->local[i].comm = comms[i];
    }
  }
  if (rc != PCQ_OK) {
    const std::string msg = last_error();
    pcq_group_destroy(g);
    return fail(rc, "%s", msg.c_str());
  }
  *out = g;
  return PCQ_OK;
}

int pcq_group_unique_id(void* id_out) {
  if (!id_out) return fail(PCQ_ERR_ARG, "null id");
  static_assert(sizeof(ncclUniqueId) == PCQ_GROUP_ID_BYTES, "ncclUniqueId is 128 bytes");
  RC(load_nccl());
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return PCQ_OK;
}

int pcq_group_create_rank(int device, uint32_t rank, uint32_t world, const void* id, pcq_group** out) {
  if (!out || world == 0 || rank >= world || (world > 1 && !id)) return fail(PCQ_ERR_ARG, "pcq_group_create_rank: bad argument");
  pcq_group* g = new (std::nothrow) pcq_group();
  if (!g) return fail(PCQ_ERR_NOMEM, "out of host memory");
  g->world = world;
  g->one_process = world == 1;
  g->use_nccl = world > 1;
  g->local.resize(1);
  g->local[0].device = device;
  g->local[0].rank = rank;
  int rc = group_finish_create(g);
  if (rc == PCQ_OK && world > 1) {
    rc = load_nccl();
    if (rc == PCQ_OK) {
      ncclUniqueId nid;
      std::memcpy(&nid, id, sizeof(nid));
      cudaSetDevice(device);
      const ncclResult_t r = g_nccl.CommInitRank(&g->local[0].comm, (int)world, nid, (int)rank);
      if (r != ncclSuccess) rc = fail(PCQ_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    }
  }
  if (rc != PCQ_OK) {
    const std::string msg = last_error();
    pcq_group_destroy(g);
    return fail(rc, "%s", msg.c_str());
  }
  *out = g;
  return PCQ_OK;
}

uint32_t pcq_group_world(const pcq_group* g) { return g ? g->world : 0; }
uint32_t pcq_group_local_count(const pcq_group* g) { return g ? (uint32_t)g->local.size() : 0; }
uint32_t pcq_group_local_rank(const pcq_group* g, uint32_t local_index) {
  return g && local_index < g->local.size() ? g->local[local_index].rank : ~0u;
}
pcq_ctx* pcq_group_ctx(pcq_group* g, uint32_t local_index) { return g && local_index < g->local.size() ? g->local[local_index].ctx : nullptr; }

uint64_t pcq_group_launch_count(const pcq_group* g) {
  uint64_t n = 0;
  if (g)
    for (const Member& M : g->local) n += M.ctx ? M.ctx->launches : 0;
  return n;
}

int pcq_group_last_stats(const pcq_group* g, pcq_group_stats* out) {
  if (!g || !out) return fail(PCQ_ERR_ARG, "null argument");
  *out = g->stats;
  return PCQ_OK;
}

int pcq_group_synchronize(pcq_group* g) {
  if (!g) return fail(PCQ_ERR_ARG, "null group");
  for (Member& M : g->local) RC(pcq_ctx_synchronize(M.ctx));
  return PCQ_OK;
}

// ---- datasets ------------------------------------------------------------------------------------------------------

static int dataset_finish(pcq_group* g, pcq_dataset* ds) {
  // who holds a range of which file: the owners of a per-file density grid are chosen among them
  const uint32_t W = g->world, F = ds->n_files;
  std::vector<std::vector<uint64_t>> mine(g->local.size(), std::vector<uint64_t>(F, 0));
  for (size_t m = 0; m < g->local.size(); ++m) {
    uint32_t prev = ~0u;
    for (const pcq_dataset::Piece& p : ds->pieces[m]) {
      if (p.file >= F) return fail(PCQ_ERR_ARG, "file index %u out of range", p.file);
      if (prev != ~0u && p.file <= prev) return fail(PCQ_ERR_ARG, "a member holds at most one range of a file, files in ascending order");
      prev = p.file;
      mine[m][p.file] = 1;
    }
  }
  if (W > 1 && !g->one_process) RC(load_nccl());
  std::vector<uint64_t> all;
  RC(allgather_u64(g, mine, F, all));
  ds->holders.assign(F, {});
  for (uint32_t f = 0; f < F; ++f)
    for (uint32_t r = 0; r < W; ++r)
      if (all[(size_t)r * F + f]) ds->holders[f].push_back(r);
  ds->file_start.assign(F + 1, 0);
  for (uint32_t f = 0; f < F; ++f) ds->file_start[f + 1] = ds->file_start[f] + ds->points_per_file[f];
  return PCQ_OK;
}

void pcq_dataset_release(pcq_dataset* ds) {
  if (!ds) return;
  for (std::vector<pcq_dataset::Piece>& v : ds->pieces)
    for (pcq_dataset::Piece& p : v)
      if (p.owned && p.f) pcq_file_release(p.f);
  delete ds;
}

int pcq_group_stage_host_files(pcq_group* g, const void* const* file_bytes, const size_t* n_bytes, const char* const* exts,
                               uint32_t n_files, int shard_mode, pcq_dataset** out) {
  if (!g || !out || (n_files && (!file_bytes || !n_bytes || !exts))) return fail(PCQ_ERR_ARG, "pcq_group_stage_host_files: null argument");
  std::vector<uint64_t> ppf(n_files);
  for (uint32_t i = 0; i < n_files; ++i) {
    const int layout = exts[i] && std::strcmp(exts[i], "las") == 0 ? PCQ_LAYOUT_LAS : (exts[i] && std::strcmp(exts[i], "last") == 0 ? PCQ_LAYOUT_LAST : -1);
    if (layout < 0) return fail(PCQ_ERR_FORMAT, "Unsupported file extension \"%s\" (this path serves las and last)", exts[i] ? exts[i] : "");
    pcq_file_desc d;
    RC(parse_header(file_bytes[i], n_bytes[i], layout, 1, &d, nullptr));
    ppf[i] = d.n_points;
  }
  uint64_t n_sh = 0;
  RC(pcq_shard_plan(ppf.data(), n_files, g->world, shard_mode, nullptr, 0, &n_sh));
  std::vector<pcq_shard> plan(n_sh);
  RC(pcq_shard_plan(ppf.data(), n_files, g->world, shard_mode, plan.data(), n_sh, &n_sh));
  pcq_dataset* ds = new (std::nothrow) pcq_dataset();
  if (!ds) return fail(PCQ_ERR_NOMEM, "out of host memory");
  ds->g = g;
  ds->n_files = n_files;
  ds->points_per_file = ppf;
  ds->pieces.resize(g->local.size());
  int rc = for_each_member(g, [&](uint32_t m) -> int {
    Member& M = g->local[m];
    for (const pcq_shard& s : plan) {
      if (s.rank != M.rank) continue;
      pcq_file* f = nullptr;
      RC(pcq_file_stage_host(M.ctx, file_bytes[s.file], n_bytes[s.file], exts[s.file], s.first_point, s.n_points, &f));
      ds->pieces[m].push_back({f, s.file, true});
    }
    return PCQ_OK;
  });
  if (rc == PCQ_OK) rc = dataset_finish(g, ds);
  if (rc != PCQ_OK) {
    const std::string msg = last_error();
    pcq_dataset_release(ds);
    return fail(rc, "%s", msg.c_str());
  }
  *out = ds;
  return PCQ_OK;
}

int pcq_group_wrap_files(pcq_group* g, const uint64_t* points_per_file, uint32_t n_files, pcq_file* const* files,
                         const uint32_t* local_member, const uint32_t* file_index, uint32_t n_local, pcq_dataset** out) {
  if (!g || !out || (n_files && !points_per_file) || (n_local && (!files || !local_member || !file_index)))
    return fail(PCQ_ERR_ARG, "pcq_group_wrap_files: null argument");
  pcq_dataset* ds = new (std::nothrow) pcq_dataset();
  if (!ds) return fail(PCQ_ERR_NOMEM, "out of host memory");
  ds->g = g;
  ds->n_files = n_files;
  ds->points_per_file.assign(points_per_file, points_per_file + n_files);
  ds->pieces.resize(g->local.size());
  int rc = PCQ_OK;
  for (uint32_t i = 0; i < n_local && rc == PCQ_OK; ++i) {
    if (local_member[i] >= g->local.size() || !files[i])
      rc = fail(PCQ_ERR_ARG, "piece %u: bad member or null file", i);
    else if (files[i]->ctx != g->local[local_member[i]].ctx)
      rc = fail(PCQ_ERR_ARG, "piece %u was not made on its member's context (pcq_group_ctx)", i);
    else
      ds->pieces[local_member[i]].push_back({files[i], file_index[i], false});
  }
  if (rc == PCQ_OK) rc = dataset_finish(g, ds);
  if (rc != PCQ_OK) {
    const std::string msg = last_error();
    pcq_dataset_release(ds);
    return fail(rc, "%s", msg.c_str());
  }
  *out = ds;
  return PCQ_OK;
}

// ---- results -------------------------------------------------------------------------------------------------------

void pcq_result_release(pcq_result* r) {
  if (!r) return;
  if (r->pending_group) finish_counts(r);
  pool_release(r);
  delete r;
}

int pcq_result_counts(pcq_result* r, const uint64_t** counts, uint32_t* n_lanes) {
  if (!r || !counts || !n_lanes) return fail(PCQ_ERR_ARG, "null argument");
  RC(finish_counts(r));
  *counts = r->counts.data();
  *n_lanes = (uint32_t)r->counts.size();
  return PCQ_OK;
}

int pcq_result_points(pcq_result* r, uint32_t lane, const pcq_point** out_points, uint64_t* out_n) {
  if (!r || !out_points || !out_n) return fail(PCQ_ERR_ARG, "null argument");
  *out_points = nullptr;
  *out_n = 0;
  if (!r->has_points) return PCQ_OK;  // CountCollector: points() == None
  if (lane >= r->n_lanes) return fail(PCQ_ERR_ARG, "lane %u out of range (%u lanes)", lane, r->n_lanes);
  const uint64_t a = r->lane_off[lane], b = r->lane_off[lane + 1];
  if (b > a) {
    *out_points = reinterpret_cast<const pcq_point*>(r->points + a * 31ull);
    *out_n = b - a;
  }
  return PCQ_OK;
}

// ---- the searches --------------------------------------------------------------------------------------------------

static int new_result(int kind, bool per_file, uint32_t n_files, pcq_result** out) {
  pcq_result* r = new (std::nothrow) pcq_result();
  if (!r) return fail(PCQ_ERR_NOMEM, "out of host memory");
  r->kind = kind;
  r->per_file = per_file;
  r->n_files = n_files;
  r->n_lanes = per_file ? n_files : 1;
  *out = r;
  return PCQ_OK;
}

int pcq_group_search(pcq_group* g, pcq_dataset* ds, const pcq_query* queries, uint32_t n_queries, int collector_kind,
                     const double gmin[3], const double gmax[3], double cell_size, int per_file, pcq_result** out) {
  RC(check_query_args(g, queries, n_queries, collector_kind, gmin, gmax));
  if (!ds || !out) return fail(PCQ_ERR_ARG, "pcq_group_search: null argument");
  if (ds->g != g) return fail(PCQ_ERR_ARG, "dataset belongs to another group");
  if (n_queries == 0) return PCQ_OK;
  RC(finish_pending(g));
  const uint32_t F = ds->n_files;
  const size_t nl = g->local.size();
  const bool grid_seq = collector_kind == PCQ_COLLECT_GRID && !per_file;
  const GridParams gp = grid_params_of(collector_kind, gmin, gmax, cell_size);
  std::vector<std::vector<MemberRun>> runs(n_queries, std::vector<MemberRun>(nl));
  std::vector<std::vector<pcq_file*>> files(nl);
  std::vector<std::vector<pcq_collector*>> cols(nl);
  // one launch per member and query over all of the member's ranges
  auto search_member = [&](uint32_t m, uint32_t q) -> int {
    Member& M = g->local[m];
    if (files[m].empty()) return PCQ_OK;
    return pcq_search_files(M.ctx, files[m].data(), (uint32_t)files[m].size(), queries + q, runs[q][m].cols.data(),
                            (uint32_t)runs[q][m].cols.size());
  };
  auto prepare_member = [&](uint32_t m) -> int {
    Member& M = g->local[m];
    RC(use_device(M.ctx));
    const std::vector<pcq_dataset::Piece>& P = ds->pieces[m];
    const size_t per_q = grid_seq ? 1 : P.size();
    RC(take_collectors(M, collector_kind, per_q * n_queries, gp, cols[m]));
    for (size_t i = 0; i < P.size(); ++i) {
      pcq_file* f = P[i].f;
      // the scan index of a point: its index in its file (one collector per file) or in the whole dataset (one grid
      // over all files, run_search_sequential)
      f->has_scan_base = true;
      f->scan_base = (grid_seq ? ds->file_start[P[i].file] : 0) + f->first_point;
      files[m].push_back(f);
    }
    for (uint32_t q = 0; q < n_queries; ++q) {
      MemberRun& R = runs[q][m];
      R.cols.assign(cols[m].begin() + q * per_q, cols[m].begin() + (q + 1) * per_q);
      if (grid_seq)
        R.lane_of.assign(1, 0u);
      else
        for (size_t i = 0; i < P.size(); ++i) R.lane_of.push_back(P[i].file);
    }
    return PCQ_OK;
  };
  std::vector<pcq_result*> res(n_queries, nullptr);
  int rc = PCQ_OK;
  for (uint32_t q = 0; q < n_queries && rc == PCQ_OK; ++q) rc = new_result(collector_kind, per_file != 0, F, &res[q]);
  if (rc == PCQ_OK && g->world > 1 && !g->one_process) rc = load_nccl();
  if (rc == PCQ_OK && collector_kind == PCQ_COLLECT_COUNT) {
    // fully asynchronous: launches, gather, (all-reduce,) copy to the host are queued; pcq_result_counts waits
    for (uint32_t m = 0; m < nl && rc == PCQ_OK; ++m) {
      rc = prepare_member(m);
      for (uint32_t q = 0; q < n_queries && rc == PCQ_OK; ++q) rc = search_member(m, q);
    }
    if (rc == PCQ_OK) rc = combine_counts(g, F, runs, res.data());
  } else if (rc == PCQ_OK) {
    g->stats = pcq_group_stats{};
    const double t_scan = now_ms();
    rc = for_each_member(g, [&](uint32_t m) -> int {
      RC(prepare_member(m));
      for (uint32_t q = 0; q < n_queries; ++q) RC(search_member(m, q));
      return pcq_ctx_synchronize(g->local[m].ctx);
    });
    g->stats.scan_ms = now_ms() - t_scan;
    for (uint32_t q = 0; q < n_queries && rc == PCQ_OK; ++q) {
      if (collector_kind == PCQ_COLLECT_BUFFER) {
        std::vector<std::vector<Stream>> streams(nl);
        for (size_t m = 0; m < nl; ++m)
          for (size_t i = 0; i < runs[q][m].cols.size(); ++i)
            streams[m].push_back({runs[q][m].lane_of[i], runs[q][m].cols[i]->d_out, runs[q][m].cols[i]->out_len});
        rc = gather_records(g, F, per_file != 0, streams, res[q]);
      } else {
        std::vector<std::vector<uint32_t>> holders;
        if (grid_seq) {
          holders.assign(1, {});
          for (uint32_t r = 0; r < g->world; ++r) holders[0].push_back(r);
        } else {
          holders = ds->holders;
        }
        rc = combine_grid(g, grid_seq ? 1u : F, per_file != 0, holders, runs[q], [&](uint32_t m) -> int { return search_member(m, q); }, res[q]);
      }
    }
  }
  if (rc != PCQ_OK) {
    const std::string msg = last_error();
    for (pcq_result* r : res)
      if (r) {
        r->pending_group = nullptr;
        g->pending.erase(std::remove(g->pending.begin(), g->pending.end(), r), g->pending.end());
        pcq_result_release(r);
      }
    return fail(rc, "%s", msg.c_str());
  }
  for (uint32_t q = 0; q < n_queries; ++q) out[q] = res[q];
  return PCQ_OK;
}

int pcq_group_search_host_files(pcq_group* g, const void* const* file_bytes, const size_t* n_bytes, const char* const* exts,
                                uint32_t n_files, const pcq_query* queries, uint32_t n_queries, int collector_kind,
                                const double gmin[3], const double gmax[3], double cell_size, int per_file, int shard_mode,
                                pcq_result** out) {
  RC(check_query_args(g, queries, n_queries, collector_kind, gmin, gmax));
  if (!out || (n_files && (!file_bytes || !n_bytes || !exts))) return fail(PCQ_ERR_ARG, "pcq_group_search_host_files: null argument");
  if (n_queries == 0) return PCQ_OK;
  RC(finish_pending(g));
  const uint32_t F = n_files, W = g->world;
  const bool grid_seq = collector_kind == PCQ_COLLECT_GRID && !per_file;
  const GridParams gp = grid_params_of(collector_kind, gmin, gmax, cell_size);
  // the plan: headers are read by everybody, point data only by the member that holds the range
  std::vector<uint64_t> ppf(F), file_start(F + 1, 0);
  for (uint32_t i = 0; i < F; ++i) {
    const int layout = exts[i] && std::strcmp(exts[i], "las") == 0 ? PCQ_LAYOUT_LAS : (exts[i] && std::strcmp(exts[i], "last") == 0 ? PCQ_LAYOUT_LAST : -1);
    if (layout < 0) return fail(PCQ_ERR_FORMAT, "Unsupported file extension \"%s\" (this path serves las and last)", exts[i] ? exts[i] : "");
    pcq_file_desc d;
    RC(parse_header(file_bytes[i], n_bytes[i], layout, 1, &d, nullptr));
    ppf[i] = d.n_points;
    file_start[i + 1] = file_start[i] + d.n_points;
  }
  uint64_t n_sh = 0;
  RC(pcq_shard_plan(ppf.data(), F, W, shard_mode, nullptr, 0, &n_sh));
  std::vector<pcq_shard> plan(n_sh);
  RC(pcq_shard_plan(ppf.data(), F, W, shard_mode, plan.data(), n_sh, &n_sh));
  std::vector<std::vector<uint32_t>> holders(grid_seq ? 1 : F);
  if (grid_seq)
    for (uint32_t r = 0; r < W; ++r) holders[0].push_back(r);
  else
    for (const pcq_shard& s : plan) holders[s.file].push_back(s.rank);
  for (std::vector<uint32_t>& h : holders) std::sort(h.begin(), h.end());

  const size_t nl = g->local.size();
  // per member: ranges of all files (n_points 0 = nothing held), collectors of all queries
  std::vector<std::vector<HostRange>> ranges(nl, std::vector<HostRange>(F, HostRange{0, 0, 0}));
  const uint32_t cols_per_query = grid_seq ? 1 : F;
  std::vector<std::vector<pcq_collector*>> cols(nl);
  std::vector<std::vector<MemberRun>> runs(n_queries, std::vector<MemberRun>(nl));
  auto search_member = [&](uint32_t m, uint32_t q0, uint32_t nq) -> int {
    Member& M = g->local[m];
    return search_host_multi(M.ctx, file_bytes, n_bytes, exts, F, queries + q0, nq, cols[m].data() + (size_t)q0 * cols_per_query,
                             cols_per_query, nullptr, ranges[m].data());
  };
  g->stats = pcq_group_stats{};
  const double t_scan = now_ms();
  int rc = for_each_member(g, [&](uint32_t m) -> int {
    Member& M = g->local[m];
    RC(use_device(M.ctx));
    for (const pcq_shard& s : plan)
      if (s.rank == M.rank) ranges[m][s.file] = HostRange{s.first_point, s.n_points, grid_seq ? file_start[s.file] : 0};
    // (every file gets a lane on every member: the host-staged scan addresses collectors by file index)
    RC(take_collectors(M, collector_kind, (size_t)n_queries * cols_per_query, gp, cols[m]));
    for (uint32_t q = 0; q < n_queries; ++q) {
      MemberRun& R = runs[q][m];
      if (grid_seq) {
        R.cols.assign(1, cols[m][q]);
        R.lane_of.assign(1, 0u);
      } else {
        for (uint32_t f = 0; f < F; ++f)
          if (ranges[m][f].n_points || collector_kind == PCQ_COLLECT_COUNT) {
            R.cols.push_back(cols[m][(size_t)q * F + f]);
            R.lane_of.push_back(f);
          }
      }
    }
    RC(search_member(m, 0, n_queries));
    return collector_kind == PCQ_COLLECT_COUNT ? PCQ_OK : pcq_ctx_synchronize(M.ctx);
  });
  g->stats.scan_ms = now_ms() - t_scan;
  if (rc == PCQ_OK && W > 1 && !g->one_process) rc = load_nccl();
  for (uint32_t q = 0; q < n_queries; ++q) out[q] = nullptr;
  for (uint32_t q = 0; q < n_queries && rc == PCQ_OK; ++q) rc = new_result(collector_kind, per_file != 0, F, &out[q]);
  if (rc == PCQ_OK && collector_kind == PCQ_COLLECT_COUNT) {
    rc = combine_counts(g, F, runs, out);
    if (rc == PCQ_OK) rc = finish_pending(g);  // host-staged searches are synchronous: the file images may go away
  }
  for (uint32_t q = 0; q < n_queries && rc == PCQ_OK && collector_kind != PCQ_COLLECT_COUNT; ++q) {
    pcq_result* res = out[q];
    if (collector_kind == PCQ_COLLECT_BUFFER) {
      std::vector<std::vector<Stream>> streams(nl);
      for (size_t m = 0; m < nl; ++m)
        for (size_t i = 0; i < runs[q][m].cols.size(); ++i)
          streams[m].push_back({runs[q][m].lane_of[i], runs[q][m].cols[i]->d_out, runs[q][m].cols[i]->out_len});
      rc = gather_records(g, F, per_file != 0, streams, res);
    } else {
      rc = combine_grid(g, grid_seq ? 1u : F, per_file != 0, holders, runs[q], [&](uint32_t m) -> int { return search_member(m, q, 1); }, res);
    }
  }
  if (rc != PCQ_OK) {
    const std::string msg = last_error();
    for (uint32_t q = 0; q < n_queries; ++q) {
      if (out[q]) {
        out[q]->pending_group = nullptr;
        g->pending.erase(std::remove(g->pending.begin(), g->pending.end(), out[q]), g->pending.end());
        pcq_result_release(out[q]);
      }
      out[q] = nullptr;
    }
    return fail(rc, "%s", msg.c_str());
  }
  return PCQ_OK;
}

}  // extern "C"
