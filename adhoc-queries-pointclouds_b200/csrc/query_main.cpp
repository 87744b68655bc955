// query_main.cpp — the `query` command line of the reference (query/src/main.rs:191-319) in front of
// the B200 scan path: same flags (-i/--input, --bounds, --class, -o/--output, --density, --parallel,
// --optimized), same stdout lines, same collector / dumper choice and ordering rules.
//
// Differences, all deliberate: only the `--optimized` implementation exists here (the Regular path and
// LAZ / LAZER inputs stay on the reference); `--gpu N` picks the device, `--gpus N` shards the files and point
// ranges of the input over N GPUs of the box (pcq_group: same results, main.rs:122-183 semantics over the whole list).
#include <dirent.h>
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "pcq_host.hpp"

using namespace pcq_host;

namespace {

// get_all_input_files (main.rs:29-57): a file, or the entries of a directory (not recursive, OS order)
std::vector<std::string> get_all_input_files(const std::string& input) {
  struct stat st;
  if (stat(input.c_str(), &st) != 0) throw Error(PCQ_ERR_IO, "Input path " + input + " does not exist!");
  if (S_ISREG(st.st_mode)) return {input};
  if (S_ISDIR(st.st_mode)) {
    std::vector<std::string> out;
    DIR* d = opendir(input.c_str());
    if (!d) throw Error(PCQ_ERR_IO, std::string(std::strerror(errno)));
    while (dirent* e = readdir(d)) {
      if (std::strcmp(e->d_name, ".") == 0 || std::strcmp(e->d_name, "..") == 0) continue;
      out.push_back(input + (input.back() == '/' ? "" : "/") + e->d_name);
    }
    closedir(d);
    return out;
  }
  throw Error(PCQ_ERR_IO, "Input path " + input + " is neither file nor directory!");
}

// is_valid_file (main.rs:185-189)
bool is_valid_file(const std::string& f) {
  size_t slash = f.find_last_of('/'), dot = f.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return false;
  const std::string ex = f.substr(dot + 1);
  return ex == "las" || ex == "laz" || ex == "last" || ex == "lazer";
}

void usage() {
  std::fputs(
      "I/O experiments 0.1 (B200 scan path)\nLAS I/O experiments\n\nUSAGE:\n    query [FLAGS] [OPTIONS] --input <FILE>\n\n"
      "FLAGS:\n        --optimized    Run search with optimized implementation (required: the only one on the GPU path)\n"
      "        --parallel     Run search in parallel: one collector per file\n\nOPTIONS:\n"
      "        --bounds <BOUNDS>      \"minX;minY;minZ;maxX;maxY;maxZ\"\n        --class <CLASS>        8-bit unsigned class\n"
      "        --density <DENSITY>    Maximum density (minimum spacing) of the result\n    -i, --input <FILE>         file or directory\n"
      "    -o, --output <OUTPUT>      output directory for matching_points_{k}.las\n        --gpu <N>              CUDA device (default 0)\n"
      "        --gpus <N>             shard files and point ranges over N GPUs of this box (default 1)\n",
      stderr);
}

// run_search_sequential (main.rs:122-144)
void run_search_sequential(const std::vector<std::string>& files, Searcher& searcher, SearchImplementation impl,
                           ResultCollector& collector, PointDumper& dumper) {
  if (!files.empty()) searcher.search_files(files, impl, {&collector});
  if (!dumper.dump_collector(collector)) std::printf("Found %zu matching points\n", collector.point_count());
}

// run_search_parallel (main.rs:146-183): one collector per file, results consumed in `files` order
void run_search_parallel(const std::vector<std::string>& files, Searcher& searcher, SearchImplementation impl,
                         std::vector<std::unique_ptr<ResultCollector>>& collectors, PointDumper& dumper) {
  std::vector<ResultCollector*> raw;
  for (auto& c : collectors) raw.push_back(c.get());
  if (!files.empty()) searcher.search_files(files, impl, raw);
  bool have_matches = false;
  size_t num_matches = 0;
  for (auto& c : collectors) {
    if (!dumper.dump_collector(*c)) {
      have_matches = true;
      num_matches += c->point_count();
    }
  }
  if (have_matches) std::printf("Found %zu matching points\n", num_matches);
}

}  // namespace

int main(int argc, char** argv) {
  const auto t_start = std::chrono::steady_clock::now();
  std::string input, bounds_s, class_s, output, density_s;
  bool have_input = false, have_bounds = false, have_class = false, have_output = false, have_density = false;
  bool parallel = false, optimized = false;
  int gpu = 0, gpus = 1;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto value = [&](const char* name) -> std::string {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "error: The argument '%s' requires a value but none was supplied\n", name);
        std::exit(1);
      }
      return argv[++i];
    };
    if (a == "-i" || a == "--input") input = value("--input <FILE>"), have_input = true;
    else if (a == "--bounds") bounds_s = value("--bounds <BOUNDS>"), have_bounds = true;  // allow_hyphen_values
    else if (a == "--class") class_s = value("--class <CLASS>"), have_class = true;
    else if (a == "-o" || a == "--output") output = value("--output <OUTPUT>"), have_output = true;
    else if (a == "--density") density_s = value("--density <DENSITY>"), have_density = true;
    else if (a == "--parallel") parallel = true;
    else if (a == "--optimized") optimized = true;
    else if (a == "--gpu") gpu = std::atoi(value("--gpu <N>").c_str());
    else if (a == "--gpus") gpus = std::atoi(value("--gpus <N>").c_str());
    else if (a == "-h" || a == "--help") { usage(); return 0; }
    else {
      std::fprintf(stderr, "error: Found argument '%s' which wasn't expected, or isn't valid in this context\n", a.c_str());
      usage();
      return 1;
    }
  }
  if (!have_input) {
    std::fputs("error: The following required arguments were not provided:\n    --input <FILE>\n", stderr);
    usage();
    return 1;
  }
  try {
    std::vector<std::string> input_files;
    for (const std::string& f : get_all_input_files(input))
      if (is_valid_file(f)) input_files.push_back(f);

    uint64_t total_file_size = 0;
    for (const std::string& f : input_files) {
      struct stat st;
      if (stat(f.c_str(), &st) == 0) total_file_size += (uint64_t)st.st_size;
    }
    const double total_file_size_mib = (double)total_file_size / 1048576.0;

    // main.rs:235-244
    std::unique_ptr<AABB> bounds;
    if (have_bounds) {
      try {
        bounds.reset(new AABB(parse_aabb(bounds_s)));
      } catch (const Error& e) {
        std::fprintf(stderr, "Could not prase argument BOUNDS: %s\n", e.what());
        return 101;  // Rust panic exit code
      }
    }
    int klass = -1;
    if (have_class) {
      char* end = nullptr;
      long v = std::strtol(class_s.c_str(), &end, 10);
      if (class_s.empty() || *end != '\0' || v < 0 || v > 255 || class_s[0] == '-') {
        std::fprintf(stderr, "Could not prase argument CLASS: invalid u8 \"%s\"\n", class_s.c_str());
        return 101;
      }
      klass = (int)v;
    }
    double density = 0.0;
    if (have_density) {
      char* end = nullptr;
      density = std::strtod(density_s.c_str(), &end);
      if (density_s.empty() || *end != '\0') {
        std::fprintf(stderr, "Could not prase argument DENSITY: invalid float literal\n");
        return 101;
      }
    }
    if (bounds && klass >= 0)
      throw Error(PCQ_ERR_ARG, "Specifying BOUNDS and CLASS at the same time is invalid! Specify either BOUNDS or CLASS argument!");
    if (!bounds && klass < 0)
      throw Error(PCQ_ERR_ARG, "Found neither BOUNDS nor CLASS argument but exactly one of these arguments is required!");

    const bool timing = std::getenv("PCQ_CLI_TIMING") != nullptr;  // phase times on stderr (not part of the reference's output)
    auto since_start = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    const double t_args = since_start();
    const SearchImplementation impl = optimized ? SearchImplementation::Optimized : SearchImplementation::Regular;
    double t_ctx = 0.0;
    // A bounds query whose box touches no file's header box finds nothing in any file — every search returns before
    // its per-point loop (`if !file_bounds.intersects(&bounds) { return Ok(()) }`, las.rs:82-84, last.rs:92-94).  That
    // is header work: no device context (seconds on these boxes) is created for it.  Anything that could make the
    // normal path fail (another extension, an unreadable header) goes the normal path.
    bool nothing_to_scan = optimized && bounds != nullptr;
    for (size_t i = 0; i < input_files.size() && nothing_to_scan; ++i) {
      MappedFile m(input_files[i]);
      const std::string ext = m.extension();
      pcq_file_desc d;
      int hit = 1;
      if ((ext != "las" && ext != "last") ||
          pcq_parse_header(m.data(), m.size(), ext == "last" ? PCQ_LAYOUT_LAST : PCQ_LAYOUT_LAS, 0, &d) != PCQ_OK ||
          pcq_file_intersects(&d, bounds->min, bounds->max, &hit) != PCQ_OK || hit)
        nothing_to_scan = false;
    }
    if (nothing_to_scan) {
      t_ctx = since_start();
      std::printf("Searching %zu files...\n", input_files.size());
      // CountCollector: "Found 0 matching points"; Buffer / GridSampled collectors hand the dumper empty slices, which
      // writes nothing and prints nothing (main.rs:134-143, dump_points.rs:65-67)
      if (!have_density && !have_output) std::printf("Found 0 matching points\n");
    } else if (gpus > 1) {
      // ---- a group of GPUs: every file is cut into point ranges, one per GPU; each GPU streams its ranges over its
      // own PCIe link; counts are summed on the host, selected records concatenated in scan order, density cells
      // exchanged by owner (group.cu).  Same collectors, same output rules.
      Group group((uint32_t)gpus);
      t_ctx = since_start();
      std::unique_ptr<AABB> grid_bounds;
      if (have_density) grid_bounds.reset(new AABB(bounds ? *bounds : get_total_bounds(input_files)));
      std::unique_ptr<Searcher> searcher;  // (carries the predicate only: the group owns the device contexts)
      if (bounds) searcher.reset(new BoundsSearcher(*bounds));
      else searcher.reset(new ClassSearcher((uint8_t)klass));
      std::unique_ptr<PointDumper> dumper;
      if (have_output) dumper.reset(new FileDumper(output));
      else dumper.reset(new IgnoreDumper());
      const int kind = have_density ? PCQ_COLLECT_GRID : (have_output ? PCQ_COLLECT_BUFFER : PCQ_COLLECT_COUNT);
      std::printf("Searching %zu files...\n", input_files.size());
      std::fflush(stdout);
      if (!input_files.empty()) {
        GroupResult res = search_files_on_group(group, input_files, impl, *searcher, kind, grid_bounds.get(), density, parallel);
        const std::vector<uint64_t> counts = res.counts();
        if (kind == PCQ_COLLECT_COUNT) {
          uint64_t total = 0;
          for (uint64_t c : counts) total += c;
          std::printf("Found %zu matching points\n", (size_t)total);
        } else {
          for (uint32_t lane = 0; lane < counts.size(); ++lane) {
            const pcq_point* pts = nullptr;
            uint64_t n = 0;
            res.points(lane, &pts, &n);
            dumper->dump_points(pts, n);
          }
        }
      } else if (kind == PCQ_COLLECT_COUNT) {
        std::printf("Found 0 matching points\n");
      }
    } else {
    // The process uses ONE GPU: hiding the others before the first CUDA call keeps the driver from initialising every
    // device of the box (seconds on an 8-GPU node; the reference's throughput line includes start-up, main.rs:192, 309-316)
    if (std::getenv("CUDA_VISIBLE_DEVICES") == nullptr) {
      setenv("CUDA_VISIBLE_DEVICES", std::to_string(gpu).c_str(), 1);
      gpu = 0;
    }
    Context ctx(gpu);
    t_ctx = since_start();
    std::unique_ptr<Searcher> searcher;
    if (bounds) searcher.reset(new BoundsSearcher(ctx, *bounds));
    else searcher.reset(new ClassSearcher(ctx, (uint8_t)klass));

    // collector factory (main.rs:253-273)
    std::unique_ptr<AABB> grid_bounds;
    if (have_density) grid_bounds.reset(new AABB(bounds ? *bounds : get_total_bounds(input_files)));
    auto make_collector = [&]() -> std::unique_ptr<ResultCollector> {
      if (have_density) return std::unique_ptr<ResultCollector>(new GridSampledCollector(ctx, *grid_bounds, density));
      if (have_output) return std::unique_ptr<ResultCollector>(new BufferCollector(ctx));
      return std::unique_ptr<ResultCollector>(new CountCollector(ctx));
    };

    std::unique_ptr<PointDumper> dumper;
    if (have_output) dumper.reset(new FileDumper(output));
    else dumper.reset(new IgnoreDumper());

    std::printf("Searching %zu files...\n", input_files.size());
    std::fflush(stdout);

    if (parallel) {
      std::vector<std::unique_ptr<ResultCollector>> collectors;
      for (size_t i = 0; i < input_files.size(); ++i) collectors.push_back(make_collector());
      run_search_parallel(input_files, *searcher, impl, collectors, *dumper);
    } else {
      std::unique_ptr<ResultCollector> collector = make_collector();
      run_search_sequential(input_files, *searcher, impl, *collector, *dumper);
    }
    }

    if (timing)
      std::fprintf(stderr, "[pcq timing] listing + arguments %.3f s, device context %.3f s, search + output %.3f s\n", t_args,
                   t_ctx - t_args, since_start() - t_ctx);
    const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    const double throughput_mibs = (double)total_file_size / elapsed / 1048576.0;
    std::printf("Searched %.2f MiB in %.2fs (throughput: %.2fMiB/s)\n", total_file_size_mib, elapsed, throughput_mibs);
    return 0;
  } catch (const Error& e) {
    std::fprintf(stderr, "Error: %s\n", e.what());
    return e.code == PCQ_ERR_PANIC ? 101 : 1;
  }
}
