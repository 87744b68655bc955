/*
 * query_ref.c — the reference's `query` command line on the CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * The reference binary cannot be built here (Rust), so the experiment harness (tools/run_query_experiments.py) times
 * this next to the GPU `query`: the same flags (query/src/main.rs:191-250), files memory-mapped like open_file_reader
 * (las.rs:24-31), run_search_sequential / run_search_parallel (main.rs:122-183) with one task per file on
 * min(files, cores) threads as rayon's par_iter schedules them, the same stdout lines.  `-o` is accepted and counted
 * (IgnoreDumper semantics: nothing is written) — the harness never passes it (run_query_experiments.rs:46-56).
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "pcq_oracle.h"

typedef struct {
  char path[4096];
  const char* ext;
  const uint8_t* data;
  size_t size;
  orc_collector* col;
  int rc;
} job_t;

static job_t* jobs;
static size_t n_jobs, next_job;
static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static int q_kind;
static double qmin[3], qmax[3], gmin[3], gmax[3], cell;
static uint8_t q_cls;
static int col_kind;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* worker(void* arg) {
  (void)arg;
  for (;;) {
    pthread_mutex_lock(&mu);
    size_t k = next_job++;
    pthread_mutex_unlock(&mu);
    if (k >= n_jobs) return NULL;
    job_t* j = &jobs[k];
    j->rc = orc_collector_new(col_kind, gmin, gmax, cell, &j->col);
    if (j->rc == ORC_OK) j->rc = orc_search_file(j->data, j->size, j->ext, q_kind, qmin, qmax, q_cls, j->col);
  }
}

static const char* ext_of(const char* p) {
  const char* dot = strrchr(p, '.');
  const char* slash = strrchr(p, '/');
  if (!dot || (slash && dot < slash)) return "";
  return dot + 1;
}

int main(int argc, char** argv) {
  const double t_start = now_s();
  const char *input = NULL, *bounds = NULL, *klass = NULL, *density = NULL, *output = NULL;
  int parallel = 0, optimized = 0;
  for (int i = 1; i < argc; ++i) {
    if ((!strcmp(argv[i], "-i") || !strcmp(argv[i], "--input")) && i + 1 < argc) input = argv[++i];
    else if (!strcmp(argv[i], "--bounds") && i + 1 < argc) bounds = argv[++i];
    else if (!strcmp(argv[i], "--class") && i + 1 < argc) klass = argv[++i];
    else if (!strcmp(argv[i], "--density") && i + 1 < argc) density = argv[++i];
    else if ((!strcmp(argv[i], "-o") || !strcmp(argv[i], "--output")) && i + 1 < argc) output = argv[++i];
    else if (!strcmp(argv[i], "--parallel")) parallel = 1;
    else if (!strcmp(argv[i], "--optimized")) optimized = 1;
    else if (!strcmp(argv[i], "--gpu") && i + 1 < argc) ++i; /* accepted and ignored: same command line as `query` */
    else {
      fprintf(stderr, "error: Found argument '%s' which wasn't expected\n", argv[i]);
      return 1;
    }
  }
  if (!input || (!bounds == !klass) || !optimized) {
    fprintf(stderr, "usage: query_ref -i <file|dir> (--bounds \"x;y;z;X;Y;Z\" | --class N) --optimized [--parallel] [--density D]\n");
    return 1;
  }
  if (bounds) {
    if (sscanf(bounds, "%lf;%lf;%lf;%lf;%lf;%lf", &qmin[0], &qmin[1], &qmin[2], &qmax[0], &qmax[1], &qmax[2]) != 6) return 101;
    for (int a = 0; a < 3; ++a)
      if (qmin[a] > qmax[a]) return 101; /* AABB::from_min_max panics */
    q_kind = 0;
  } else {
    q_kind = 1;
    q_cls = (uint8_t)atoi(klass);
  }
  /* get_all_input_files + is_valid_file (main.rs:29-57, 185-189); laz / lazer are not on this path */
  struct stat st;
  if (stat(input, &st) != 0) {
    fprintf(stderr, "Error: Input path %s does not exist!\n", input);
    return 1;
  }
  size_t cap = 1024;
  jobs = calloc(cap, sizeof(job_t));
  if (S_ISDIR(st.st_mode)) {
    DIR* d = opendir(input);
    struct dirent* e;
    while (d && (e = readdir(d))) {
      const char* x = ext_of(e->d_name);
      if (strcmp(x, "las") && strcmp(x, "last")) continue;
      if (n_jobs == cap) jobs = realloc(jobs, (cap *= 2) * sizeof(job_t));
      memset(&jobs[n_jobs], 0, sizeof(job_t));
      snprintf(jobs[n_jobs].path, sizeof(jobs[n_jobs].path), "%s/%s", input, e->d_name);
      ++n_jobs;
    }
    if (d) closedir(d);
  } else {
    snprintf(jobs[0].path, sizeof(jobs[0].path), "%s", input);
    n_jobs = 1;
  }
  uint64_t total_size = 0;
  for (size_t k = 0; k < n_jobs; ++k) {
    job_t* j = &jobs[k];
    j->ext = ext_of(j->path);
    int fd = open(j->path, O_RDONLY);
    if (fd < 0 || fstat(fd, &st) != 0) {
      fprintf(stderr, "Error: cannot open %s\n", j->path);
      return 1;
    }
    j->size = (size_t)st.st_size;
    total_size += (uint64_t)st.st_size;
    j->data = j->size ? mmap(NULL, j->size, PROT_READ, MAP_PRIVATE, fd, 0) : NULL;
    close(fd);
  }
  /* collector factory (main.rs:253-273) */
  col_kind = density ? ORC_COLLECT_GRID : (output ? ORC_COLLECT_BUFFER : ORC_COLLECT_COUNT);
  if (density) {
    cell = atof(density);
    if (bounds) {
      memcpy(gmin, qmin, sizeof(gmin));
      memcpy(gmax, qmax, sizeof(gmax));
    } else { /* get_total_bounds (main.rs:94-120) */
      for (int a = 0; a < 3; ++a) gmin[a] = 1.7976931348623157e308, gmax[a] = -1.7976931348623157e308;
      for (size_t k = 0; k < n_jobs; ++k) {
        orc_header h;
        if (orc_parse_header(jobs[k].data, jobs[k].size, !strcmp(jobs[k].ext, "last"), &h) != ORC_OK) return 1;
        for (int a = 0; a < 3; ++a) {
          if (h.min[a] < gmin[a]) gmin[a] = h.min[a];
          if (h.max[a] > gmax[a]) gmax[a] = h.max[a];
        }
      }
    }
  }
  printf("Searching %zu files...\n", n_jobs);
  fflush(stdout);
  size_t matches = 0, dumped = 0;
  int failed = 0;
  if (parallel) {
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    size_t n_thr = n_jobs < (size_t)cores ? n_jobs : (size_t)cores;
    if (n_thr == 0) n_thr = 1;
    pthread_t* th = calloc(n_thr, sizeof(pthread_t));
    for (size_t t = 0; t < n_thr; ++t) pthread_create(&th[t], NULL, worker, NULL);
    for (size_t t = 0; t < n_thr; ++t) pthread_join(th[t], NULL);
    for (size_t k = 0; k < n_jobs; ++k) {
      if (jobs[k].rc != ORC_OK) failed = jobs[k].rc;
      else if (col_kind == ORC_COLLECT_COUNT) matches += orc_collector_point_count(jobs[k].col);
      else dumped += orc_collector_point_count(jobs[k].col);
    }
  } else {
    orc_collector* c = NULL;
    failed = orc_collector_new(col_kind, gmin, gmax, cell, &c);
    for (size_t k = 0; k < n_jobs && failed == ORC_OK; ++k)
      failed = orc_search_file(jobs[k].data, jobs[k].size, jobs[k].ext, q_kind, qmin, qmax, q_cls, c);
    if (failed == ORC_OK) {
      if (col_kind == ORC_COLLECT_COUNT) matches = orc_collector_point_count(c);
      else dumped = orc_collector_point_count(c);
    }
  }
  if (failed != ORC_OK) {
    fprintf(stderr, "Error: search failed (%d)\n", failed);
    return failed == ORC_ERR_PANIC ? 101 : 1;
  }
  if (col_kind == ORC_COLLECT_COUNT) printf("Found %zu matching points\n", matches);
  (void)dumped;
  const double el = now_s() - t_start;
  printf("Searched %.2f MiB in %.2fs (throughput: %.2fMiB/s)\n", (double)total_size / 1048576.0, el, (double)total_size / el / 1048576.0);
  return 0;
}
