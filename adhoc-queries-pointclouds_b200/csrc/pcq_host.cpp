// pcq_host.cpp — implementation of the C++ host mirror (pcq_host.hpp) on top of the C ABI.
#include "pcq_host.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>

namespace pcq_host {

static void check(int rc) {
  if (rc != PCQ_OK) throw Error(rc, pcq_last_error());
}

AABB AABB::from_min_max(const double mn[3], const double mx[3]) {
  for (int i = 0; i < 3; ++i)
    if (mn[i] > mx[i]) throw Error(PCQ_ERR_PANIC, "AABB::from_min_max: Minimum position must be <= maximum position!");
  return from_min_max_unchecked(mn, mx);
}
AABB AABB::from_min_max_unchecked(const double mn[3], const double mx[3]) {
  AABB b;
  for (int i = 0; i < 3; ++i) {
    b.min[i] = mn[i];
    b.max[i] = mx[i];
  }
  return b;
}
AABB AABB::union_of(const AABB& a, const AABB& b) {
  AABB r;
  for (int i = 0; i < 3; ++i) {
    r.min[i] = std::min(a.min[i], b.min[i]);
    r.max[i] = std::max(a.max[i], b.max[i]);
  }
  return r;
}

AABB parse_aabb(const std::string& s) {
  // main.rs:59-92
  std::vector<std::string> parts;
  size_t start = 0;
  for (;;) {
    size_t p = s.find(';', start);
    parts.push_back(s.substr(start, p == std::string::npos ? std::string::npos : p - start));
    if (p == std::string::npos) break;
    start = p + 1;
  }
  if (parts.size() != 6) throw Error(PCQ_ERR_ARG, "Could not parse AABB from string \"" + s + "\"");
  double v[6];
  for (int i = 0; i < 6; ++i) {
    const std::string& t = parts[i];
    char* end = nullptr;
    // Rust's f64::from_str accepts no surrounding whitespace and no empty strings
    if (t.empty() || std::isspace((unsigned char)t.front()) || std::isspace((unsigned char)t.back()))
      throw Error(PCQ_ERR_ARG, "Could not parse AABB from string \"" + s + "\": invalid float literal");
    v[i] = std::strtod(t.c_str(), &end);
    if (end == t.c_str() || *end != '\0')
      throw Error(PCQ_ERR_ARG, "Could not parse AABB from string \"" + s + "\": invalid float literal");
  }
  return AABB::from_min_max(v, v + 3);
}

Context::Context(int device) { check(pcq_ctx_create(device, &ctx_)); }
Context::~Context() { pcq_ctx_destroy(ctx_); }

ResultCollector::ResultCollector(Context& ctx, int kind, const AABB* bounds, double cell_size) : kind_(kind) {
  check(pcq_collector_create(ctx.get(), kind, bounds ? bounds->min : nullptr, bounds ? bounds->max : nullptr, cell_size, &h_));
}
ResultCollector::~ResultCollector() { pcq_collector_destroy(h_); }

size_t ResultCollector::point_count() {
  uint64_t n = 0;
  check(pcq_collector_point_count(h_, &n));
  return (size_t)n;
}

std::optional<std::vector<pcq_point>> ResultCollector::points() {
  if (kind_ == PCQ_COLLECT_COUNT) return std::nullopt;  // collect_points.rs:88-90
  const pcq_point* p = nullptr;
  uint64_t n = 0;
  check(pcq_collector_points(h_, &p, &n));
  return std::vector<pcq_point>(p, p + n);
}

bool ResultCollector::points_ref(const pcq_point** out, uint64_t* n) {
  if (kind_ != PCQ_COLLECT_BUFFER) return false;  // only BufferCollector has points_ref (collect_points.rs:37-39)
  check(pcq_collector_points(h_, out, n));
  return true;
}

MappedFile::MappedFile(const std::string& path) : path_(path) {
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) throw Error(PCQ_ERR_IO, "cannot open " + path + ": " + std::strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0) {
    ::close(fd);
    throw Error(PCQ_ERR_IO, "cannot stat " + path);
  }
  size_ = (size_t)st.st_size;
  if (size_ > 0) {
    data_ = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    if (data_ == MAP_FAILED) {
      data_ = nullptr;
      ::close(fd);
      throw Error(PCQ_ERR_IO, "cannot mmap " + path);
    }
    madvise(data_, size_, MADV_SEQUENTIAL);
  }
  ::close(fd);
}
MappedFile::~MappedFile() {
  if (data_) munmap(data_, size_);
}
std::string MappedFile::extension() const {
  size_t slash = path_.find_last_of('/');
  size_t dot = path_.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash) || dot + 1 == path_.size()) return "";
  return path_.substr(dot + 1);
}

void Searcher::search_file(const std::string& path, SearchImplementation impl, ResultCollector& collector) {
  search_files({path}, impl, {&collector});
}

namespace {

// the inputs of one search: mapped like the reference maps them (las.rs:24-31), extensions checked as searcher.rs:50-90
struct MappedInputs {
  std::vector<std::unique_ptr<MappedFile>> maps;
  std::vector<const void*> ptrs;
  std::vector<size_t> sizes;
  std::vector<std::string> exts;
  std::vector<const char*> ext_c;
  explicit MappedInputs(const std::vector<std::string>& paths) {
    for (const std::string& p : paths) {
      maps.emplace_back(new MappedFile(p));
      std::string ext = maps.back()->extension();
      if (ext.empty()) throw Error(PCQ_ERR_FORMAT, "Invalid extension on file " + p);
      if (ext == "laz" || ext == "lazer")
        throw Error(PCQ_ERR_FORMAT, "file " + p + ": LAZ / LAZER decoding is out of scope of the accelerated path and stays on the reference");
      if (ext != "las" && ext != "last") throw Error(PCQ_ERR_FORMAT, "Unsupported file extension in file " + p);
      ptrs.push_back(maps.back()->data());
      sizes.push_back(maps.back()->size());
      exts.push_back(ext);
    }
    for (const std::string& e : exts) ext_c.push_back(e.c_str());
  }
};

void require_optimized(SearchImplementation impl) {
  if (impl != SearchImplementation::Optimized)
    throw Error(PCQ_ERR_ARG,
                "SearchImplementation::Regular (pasture readers + f64 AABB::contains) is not part of the accelerated path; "
                "pass --optimized or use the reference");
}

}  // namespace

void Searcher::search_files(const std::vector<std::string>& paths, SearchImplementation impl,
                            const std::vector<ResultCollector*>& collectors) {
  require_optimized(impl);
  if (!ctx_) throw Error(PCQ_ERR_ARG, "this searcher has no device context (it was made for a Group search)");
  MappedInputs in(paths);
  std::vector<pcq_collector*> ch;
  for (ResultCollector* c : collectors) ch.push_back(c->handle());
  pcq_query q = query();
  check(pcq_search_host_files(ctx_->get(), in.ptrs.data(), in.sizes.data(), in.ext_c.data(), (uint32_t)paths.size(), &q, ch.data(),
                              (uint32_t)ch.size()));
  check(pcq_ctx_synchronize(ctx_->get()));  // the mappings go away when we return
}

Group::Group(uint32_t n_gpus) { check(pcq_group_create(nullptr, n_gpus, &g_)); }
Group::~Group() { pcq_group_destroy(g_); }

GroupResult::~GroupResult() { pcq_result_release(r_); }
std::vector<uint64_t> GroupResult::counts() const {
  const uint64_t* c = nullptr;
  uint32_t n = 0;
  check(pcq_result_counts(r_, &c, &n));
  return std::vector<uint64_t>(c, c + n);
}
bool GroupResult::points(uint32_t lane, const pcq_point** out, uint64_t* n) const {
  check(pcq_result_points(r_, lane, out, n));
  return *out != nullptr || *n == 0;
}

GroupResult search_files_on_group(Group& group, const std::vector<std::string>& paths, SearchImplementation impl,
                                  const Searcher& searcher, int kind, const AABB* grid_bounds, double cell_size, bool per_file) {
  require_optimized(impl);
  MappedInputs in(paths);
  const pcq_query q = searcher.to_query();
  pcq_result* r = nullptr;
  check(pcq_group_search_host_files(group.get(), in.ptrs.data(), in.sizes.data(), in.ext_c.data(), (uint32_t)paths.size(), &q, 1, kind,
                                    grid_bounds ? grid_bounds->min : nullptr, grid_bounds ? grid_bounds->max : nullptr, cell_size,
                                    per_file ? 1 : 0, PCQ_SHARD_RANGES, &r));
  return GroupResult(r);
}

pcq_query BoundsSearcher::query() const {
  pcq_query q;
  std::memset(&q, 0, sizeof(q));
  q.kind = PCQ_QUERY_BOUNDS;
  for (int i = 0; i < 3; ++i) {
    q.qmin[i] = bounds_.min[i];
    q.qmax[i] = bounds_.max[i];
  }
  return q;
}

pcq_query ClassSearcher::query() const {
  pcq_query q;
  std::memset(&q, 0, sizeof(q));
  q.kind = PCQ_QUERY_CLASS;
  q.cls = class_;
  return q;
}

AABB get_total_bounds(const std::vector<std::string>& files) {
  // main.rs:94-120
  const double hi[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, lo[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  AABB total = AABB::from_min_max_unchecked(hi, lo);
  for (const std::string& f : files) {
    MappedFile m(f);
    const std::string ext = m.extension();
    pcq_file_desc d;
    check(pcq_parse_header(m.data(), m.size(), ext == "last" ? PCQ_LAYOUT_LAST : PCQ_LAYOUT_LAS, ext == "last" ? 1 : 0, &d));
    total = AABB::union_of(total, AABB::from_min_max_unchecked(d.hdr_min, d.hdr_max));
  }
  return total;
}

// ---- FileDumper: LAS 1.2, point format 2 (dump_points.rs:63-116) ------------------------------------
FileDumper::FileDumper(const std::string& root_dir) : root_(root_dir) {
  struct stat st;
  if (stat(root_dir.c_str(), &st) != 0) throw Error(PCQ_ERR_IO, "Path " + root_dir + " does not exist!");
  if (!S_ISDIR(st.st_mode)) throw Error(PCQ_ERR_IO, "Path " + root_dir + " is no directory!");
}

template <typename T>
static void put(std::vector<uint8_t>& b, size_t off, T v) {
  std::memcpy(b.data() + off, &v, sizeof(T));
}

bool PointDumper::dump_collector(ResultCollector& collector) {
  const pcq_point* ref = nullptr;
  uint64_t n = 0;
  if (collector.points_ref(&ref, &n)) {
    dump_points(ref, n);
    return true;
  }
  if (auto pts = collector.points()) {
    dump_points(pts->data(), pts->size());
    return true;
  }
  return false;
}

static double writer_scale(double max_extent) {
  // :81-88
  const double min_scale = max_extent / (double)INT32_MAX;
  double scale = std::pow(10.0, std::ceil(std::log10(min_scale)));
  if (!(scale >= 0.001)) scale = 0.001;  // `if scale < 0.001` plus the NaN/0 cases of a zero extent
  return scale;
}

void FileDumper::write_file(const double mn[3], const double mx[3], double scale, const uint8_t* records, size_t n) {
  const std::string path = root_ + "/matching_points_" + std::to_string(file_index_) + ".las";
  file_index_ += 1;
  std::printf("Writing %zu points\n", n);  // :108
  const size_t rec = 26;
  std::vector<uint8_t> out(227, 0);
  std::memcpy(out.data(), "LASF", 4);
  out[24] = 1;
  out[25] = 2;
  std::snprintf(reinterpret_cast<char*>(out.data() + 26), 32, "pcq-b200");
  std::snprintf(reinterpret_cast<char*>(out.data() + 58), 32, "pcq query");
  put<uint16_t>(out, 90, 1);
  put<uint16_t>(out, 92, 2026);
  put<uint16_t>(out, 94, 227);
  put<uint32_t>(out, 96, 227);
  put<uint32_t>(out, 100, 0);
  out[104] = 2;
  put<uint16_t>(out, 105, (uint16_t)rec);
  put<uint32_t>(out, 107, (uint32_t)std::min<size_t>(n, 0xFFFFFFFFu));
  put<uint32_t>(out, 111, (uint32_t)std::min<size_t>(n, 0xFFFFFFFFu));
  for (int a = 0; a < 3; ++a) {
    put<double>(out, 131 + 8 * a, scale);
    put<double>(out, 155 + 8 * a, mn[a]);
    put<double>(out, 179 + 16 * a, mx[a]);
    put<double>(out, 187 + 16 * a, mn[a]);
  }
  std::ofstream f(path, std::ios::binary);
  if (!f) throw Error(PCQ_ERR_IO, "cannot create " + path);
  f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
  f.write(reinterpret_cast<const char*>(records), (std::streamsize)(n * rec));
  if (!f) throw Error(PCQ_ERR_IO, "cannot write " + path);
  dumped_ += n;
}

// points that are already on the host (a group's result): the per-point work happens here
void FileDumper::dump_points(const pcq_point* points, size_t n) {
  if (n == 0) return;  // :65-67
  // :74-88 — offset = min position, one scale for all axes
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t i = 0; i < n; ++i) {
    pcq_point p;
    std::memcpy(&p, points + i, sizeof(p));
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], p.pos[a]);
      mx[a] = std::max(mx[a], p.pos[a]);
    }
  }
  const double scale = writer_scale(std::max({mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]}));
  // Byte-level parity with pasture-io's LASWriter is unpinned (un-vendored crate): raw coordinates use
  // las-rs' Transform::inverse rule round((p - offset) / scale).
  const size_t rec = 26;
  std::vector<uint8_t> out(n * rec, 0);
  for (size_t i = 0; i < n; ++i) {
    pcq_point p;
    std::memcpy(&p, points + i, sizeof(p));
    uint8_t* r = out.data() + i * rec;
    for (int a = 0; a < 3; ++a) {
      const double q = std::round((p.pos[a] - mn[a]) / scale);
      const int32_t v = q >= 2147483647.0 ? INT32_MAX : (q <= -2147483648.0 ? INT32_MIN : (int32_t)q);
      std::memcpy(r + 4 * a, &v, 4);
    }
    r[14] = 0x09;  // return 1 of 1
    r[15] = p.cls;
    std::memcpy(r + 20, p.rgb, 6);
  }
  write_file(mn, mx, scale, out.data(), n);
}

// a collector that lives in HBM: the library reduces and quantises on the device
bool FileDumper::dump_collector(ResultCollector& collector) {
  if (collector.kind() == PCQ_COLLECT_COUNT) return false;  // `points()` == None
  double mn[3], mx[3], scale = 0.0;
  const uint8_t* records = nullptr;
  uint64_t n = 0;
  check(pcq_collector_las_records(collector.handle(), mn, mx, &scale, &records, &n));
  if (n != 0) write_file(mn, mx, scale, records, (size_t)n);
  return true;
}

}  // namespace pcq_host
