// pcq_host.hpp — C++ host mirror of the reference's operator interface for the scan path, layered on
// the C ABI (include/pcq.h).  The reference is compiled Rust; with no Rust toolchain in this image
// the host side is C++ with the same names, argument meaning and error behaviour:
//
//   query/src/search/searcher.rs   SearchImplementation, Searcher, BoundsSearcher, ClassSearcher
//   query/src/collect_points.rs    ResultCollector, CountCollector, BufferCollector, GridSampledCollector
//   query/src/dump_points.rs       PointDumper, IgnoreDumper, FileDumper
//   query/src/main.rs:59-92, 122-183  parse_aabb, run_search_sequential, run_search_parallel
//
// anyhow::Error / panics become pcq_host::Error (a std::runtime_error carrying the pcq_status).
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pcq.h"

namespace pcq_host {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// pasture_core::math::AABB<f64>
struct AABB {
  double min[3];
  double max[3];
  // panics (throws PCQ_ERR_PANIC) when min > max on any axis
  static AABB from_min_max(const double mn[3], const double mx[3]);
  static AABB from_min_max_unchecked(const double mn[3], const double mx[3]);
  static AABB union_of(const AABB& a, const AABB& b);
};

// "minX;minY;minZ;maxX;maxY;maxZ" (main.rs:59-92)
AABB parse_aabb(const std::string& s);

// RAII wrapper of pcq_ctx
class Context {
 public:
  explicit Context(int device = 0);
  ~Context();
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  pcq_ctx* get() const { return ctx_; }

 private:
  pcq_ctx* ctx_ = nullptr;
};

// trait ResultCollector (collect_points.rs:7-12); filled in bulk by Searcher::search_file(s)
class ResultCollector {
 public:
  virtual ~ResultCollector();
  ResultCollector(const ResultCollector&) = delete;
  ResultCollector& operator=(const ResultCollector&) = delete;
  // Option<Vec<Point>>
  virtual std::optional<std::vector<pcq_point>> points();
  // Option<&[Point]>: pointer stays valid until the next call on this collector
  virtual bool points_ref(const pcq_point** out, uint64_t* n);
  size_t point_count();
  pcq_collector* handle() const { return h_; }
  int kind() const { return kind_; }

 protected:
  ResultCollector(Context& ctx, int kind, const AABB* bounds, double cell_size);
  pcq_collector* h_ = nullptr;
  int kind_;
};

class CountCollector : public ResultCollector {
 public:
  explicit CountCollector(Context& ctx) : ResultCollector(ctx, PCQ_COLLECT_COUNT, nullptr, 0.0) {}
};
class BufferCollector : public ResultCollector {
 public:
  explicit BufferCollector(Context& ctx) : ResultCollector(ctx, PCQ_COLLECT_BUFFER, nullptr, 0.0) {}
};
class GridSampledCollector : public ResultCollector {
 public:
  // GridSampledCollector::new(bounds, cell_size) (collect_points.rs:104-108)
  GridSampledCollector(Context& ctx, const AABB& bounds, double cell_size)
      : ResultCollector(ctx, PCQ_COLLECT_GRID, &bounds, cell_size) {}
};

enum class SearchImplementation { Regular, Optimized };

// a memory-mapped input file (open_file_reader, las.rs:24-31)
class MappedFile {
 public:
  explicit MappedFile(const std::string& path);
  ~MappedFile();
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
  const void* data() const { return data_; }
  size_t size() const { return size_; }
  const std::string& path() const { return path_; }
  // "las" / "last" / ... ; empty when the path has no extension
  std::string extension() const;

 private:
  std::string path_;
  void* data_ = nullptr;
  size_t size_ = 0;
};

// trait Searcher (searcher.rs:24-31)
class Searcher {
 public:
  explicit Searcher(Context& ctx) : ctx_(&ctx) {}
  Searcher() = default;  // predicate only: searched through a Group (search_files_on_group)
  virtual ~Searcher() = default;
  void search_file(const std::string& path, SearchImplementation impl, ResultCollector& collector);
  // one batch: collectors.size() == 1 (sequential) or == paths.size() (parallel)
  void search_files(const std::vector<std::string>& paths, SearchImplementation impl,
                    const std::vector<ResultCollector*>& collectors);
  // the predicate as the C ABI takes it
  pcq_query to_query() const { return query(); }

 protected:
  virtual pcq_query query() const = 0;
  Context* ctx_ = nullptr;
};

class BoundsSearcher : public Searcher {
 public:
  BoundsSearcher(Context& ctx, const AABB& bounds) : Searcher(ctx), bounds_(bounds) {}
  explicit BoundsSearcher(const AABB& bounds) : bounds_(bounds) {}

 protected:
  pcq_query query() const override;
  AABB bounds_;
};

class ClassSearcher : public Searcher {
 public:
  ClassSearcher(Context& ctx, uint8_t cls) : Searcher(ctx), class_(cls) {}
  explicit ClassSearcher(uint8_t cls) : class_(cls) {}

 protected:
  pcq_query query() const override;
  uint8_t class_;
};

// A group of GPUs of one box (pcq_group): files and point ranges of files shard across them, the results are those of
// run_search_sequential / run_search_parallel over the whole file list (main.rs:122-183).
class Group {
 public:
  explicit Group(uint32_t n_gpus);
  ~Group();
  Group(const Group&) = delete;
  Group& operator=(const Group&) = delete;
  pcq_group* get() const { return g_; }

 private:
  pcq_group* g_ = nullptr;
};

// what the collectors of one group search hold: one lane (sequential) or one per file (parallel)
class GroupResult {
 public:
  explicit GroupResult(pcq_result* r) : r_(r) {}
  ~GroupResult();
  GroupResult(GroupResult&& o) noexcept : r_(o.r_) { o.r_ = nullptr; }
  GroupResult(const GroupResult&) = delete;
  GroupResult& operator=(const GroupResult&) = delete;
  std::vector<uint64_t> counts() const;
  // false for a CountCollector (`points()` == None)
  bool points(uint32_t lane, const pcq_point** out, uint64_t* n) const;

 private:
  pcq_result* r_;
};

// kind: pcq_collector_kind; grid_bounds / cell_size as GridSampledCollector::new; per_file = one collector per file
GroupResult search_files_on_group(Group& group, const std::vector<std::string>& paths, SearchImplementation impl,
                                  const Searcher& searcher, int kind, const AABB* grid_bounds, double cell_size, bool per_file);

// dump_points.rs
class PointDumper {
 public:
  virtual ~PointDumper() = default;
  virtual void dump_points(const pcq_point* points, size_t n) = 0;
  virtual size_t num_dumped_points() const = 0;
  // what run_search_* do with a finished collector (main.rs:134-143, 165-169): hand its points to dump_points.
  // Returns false for a CountCollector (`points()` == None).
  virtual bool dump_collector(ResultCollector& collector);
};

class IgnoreDumper : public PointDumper {
 public:
  void dump_points(const pcq_point*, size_t n) override { dumped_ += n; }
  // (counts what it is handed, dump_points.rs:27-30: no point has to leave the device for that)
  bool dump_collector(ResultCollector& collector) override {
    if (collector.kind() == PCQ_COLLECT_COUNT) return false;
    dumped_ += collector.point_count();
    return true;
  }
  size_t num_dumped_points() const override { return dumped_; }

 private:
  size_t dumped_ = 0;
};

// FileDumper (dump_points.rs:39-121): matching_points_{k}.las, LAS 1.2 point format 2,
// offset = min position, scale = max(10^ceil(log10(max_extent / i32::MAX)), 0.001)
class FileDumper : public PointDumper {
 public:
  explicit FileDumper(const std::string& root_dir);
  void dump_points(const pcq_point* points, size_t n) override;
  // the same file, with the min / max reduction and the quantisation done on the device (pcq_collector_las_records)
  bool dump_collector(ResultCollector& collector) override;
  size_t num_dumped_points() const override { return dumped_; }

 private:
  void write_file(const double mn[3], const double mx[3], double scale, const uint8_t* records, size_t n);
  std::string root_;
  size_t file_index_ = 0;
  size_t dumped_ = 0;
};

// bounds of all input files (get_total_bounds, main.rs:94-120): union of the header bounds
AABB get_total_bounds(const std::vector<std::string>& files);

}  // namespace pcq_host
