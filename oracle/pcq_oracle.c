/*
 * pcq_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see pcq_oracle.h for the rules and the
 * parity status).  Plain C restatement of:
 *   query/src/search/las.rs:52-148, 192-261      LAS  bounds / class, Optimized
 *   query/src/search/last.rs:46-166, 213-293     LAST bounds / class, Optimized
 *   query/src/grid_sampling.rs:18-105            SparseGrid::new / insert_point
 *   query/src/collect_points.rs:14-44, 72-127    Buffer / Count / GridSampled collectors
 *   query/src/search/searcher.rs:43-91, 104-152  dispatch on extension
 *   query/src/main.rs:146-183                    run_search_parallel (count collectors)
 * Compile with -ffp-contract=off: Rust never contracts a*b+c into an FMA.
 */
#include "pcq_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Rust `as` casts: truncate toward zero, saturate, NaN -> 0.
 * ------------------------------------------------------------------------------------------- */
int64_t orc_f64_as_i64(double v) {
  if (v != v) return 0;
  if (v >= 9223372036854775808.0) return INT64_MAX;
  if (v <= -9223372036854775808.0) return INT64_MIN;
  return (int64_t)v;
}

uint64_t orc_f64_as_u64(double v) {
  if (v != v) return 0;
  if (v <= 0.0) return 0;
  if (v >= 18446744073709551616.0) return UINT64_MAX;
  return (uint64_t)v;
}

/* ---------------------------------------------------------------------------------------------
 * Little-endian field readers on a byte cursor (Cursor<Mmap> + byteorder in the reference).
 * ------------------------------------------------------------------------------------------- */
static uint16_t rd_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t rd_u32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static int32_t rd_i32(const uint8_t* p) { return (int32_t)rd_u32(p); }
static uint64_t rd_u64(const uint8_t* p) { return (uint64_t)rd_u32(p) | ((uint64_t)rd_u32(p + 4) << 32); }
static double rd_f64(const uint8_t* p) {
  uint64_t u = rd_u64(p);
  double d;
  memcpy(&d, &u, 8);
  return d;
}

/* point record length of LAS point formats 0..10 (ASPRS LAS 1.4 R15, tables 7-17) */
static const uint16_t k_format_len[11] = {20, 28, 26, 34, 57, 63, 30, 36, 38, 59, 67};

/* ---------------------------------------------------------------------------------------------
 * parse_las_header (las.rs:33-36) = las::raw::Header::read_from, then Header::from_raw
 * (las.rs:60).  The `las` crate (0.7.4) is not vendored: the byte layout is the ASPRS LAS 1.x public
 * header block (field order corroborated by query/src/las.rs:6-40).  Assumed validations: "LASF"
 * signature, format <= 10, record length >= format length (extra bytes allowed), number_of_points =
 * legacy u32 count unless it is 0 and a LAS 1.4 64-bit count exists.
 * ------------------------------------------------------------------------------------------- */
int orc_parse_header(const uint8_t* b, size_t n, int mask_format, orc_header* h) {
  memset(h, 0, sizeof(*h));
  if (n < 227) return ORC_ERR_IO;
  if (memcmp(b, "LASF", 4) != 0) return ORC_ERR_FORMAT;
  h->version_major = b[24];
  h->version_minor = b[25];
  h->header_size = rd_u16(b + 94);
  h->offset_to_point_data = rd_u32(b + 96);
  h->n_vlrs = rd_u32(b + 100);
  h->format = b[104];
  h->record_len = rd_u16(b + 105);
  h->legacy_count = rd_u32(b + 107);
  for (int i = 0; i < 3; ++i) h->scale[i] = rd_f64(b + 131 + 8 * i);
  for (int i = 0; i < 3; ++i) h->offset[i] = rd_f64(b + 155 + 8 * i);
  for (int i = 0; i < 3; ++i) {
    h->max[i] = rd_f64(b + 179 + 16 * i);
    h->min[i] = rd_f64(b + 187 + 16 * i);
  }
  size_t need = 227;
  int v13 = h->version_major > 1 || (h->version_major == 1 && h->version_minor >= 3);
  int v14 = h->version_major > 1 || (h->version_major == 1 && h->version_minor >= 4);
  if (v13) need += 8;
  if (v14) need += 140;
  if (n < need) return ORC_ERR_IO;
  if (v14) {
    h->has_large = 1;
    h->large_count = rd_u64(b + 247);
  }
  if (h->header_size > need && n < h->header_size) return ORC_ERR_IO; /* padding read */

  /* last.rs:222 / last_reader.rs:76-79 */
  if (mask_format) h->format &= 0x0F;

  /* Header::from_raw */
  if (h->format > 10) return ORC_ERR_FORMAT;
  if (h->record_len < k_format_len[h->format]) return ORC_ERR_FORMAT;
  if (h->format >= 6 && !v14) return ORC_ERR_FORMAT;
  h->n_points = h->legacy_count > 0 ? (uint64_t)h->legacy_count : (h->has_large ? h->large_count : 0);
  return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * las.rs:88-99 / last.rs:98-109 — NOTE min.y and min.z divide by x_scale_factor (reference quirk).
 * AABB::<i64>::from_min_max panics when min > max on any axis.
 * ------------------------------------------------------------------------------------------- */
int orc_local_bounds(const orc_header* h, const double qmin[3], const double qmax[3], int64_t lo[3],
                     int64_t hi[3]) {
  lo[0] = orc_f64_as_i64((qmin[0] - h->offset[0]) / h->scale[0]);
  lo[1] = orc_f64_as_i64((qmin[1] - h->offset[1]) / h->scale[0]);
  lo[2] = orc_f64_as_i64((qmin[2] - h->offset[2]) / h->scale[0]);
  hi[0] = orc_f64_as_i64((qmax[0] - h->offset[0]) / h->scale[0]);
  hi[1] = orc_f64_as_i64((qmax[1] - h->offset[1]) / h->scale[1]);
  hi[2] = orc_f64_as_i64((qmax[2] - h->offset[2]) / h->scale[2]);
  for (int i = 0; i < 3; ++i)
    if (lo[i] > hi[i]) return ORC_ERR_PANIC;
  return ORC_OK;
}

/* pasture_core::math::AABB::from_min_max (panics on min > max) + intersects (closed intervals) */
static int aabb_check(const double mn[3], const double mx[3]) {
  for (int i = 0; i < 3; ++i)
    if (mn[i] > mx[i]) return ORC_ERR_PANIC;
  return ORC_OK;
}
static int aabb_intersects(const double amin[3], const double amax[3], const double bmin[3],
                           const double bmax[3]) {
  for (int i = 0; i < 3; ++i)
    if (!(amin[i] <= bmax[i] && amax[i] >= bmin[i])) return 0;
  return 1;
}

/* ---------------------------------------------------------------------------------------------
 * SparseGrid (grid_sampling.rs)
 * ------------------------------------------------------------------------------------------- */
static uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

static int grid_alloc(orc_grid* g, size_t cap) {
  g->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
  g->vals = (orc_point*)malloc(cap * sizeof(orc_point));
  g->used = (uint8_t*)calloc(cap, 1);
  g->cap = cap;
  g->len = 0;
  return (g->keys && g->vals && g->used) ? ORC_OK : ORC_ERR_IO;
}

/* grid_sampling.rs:18-47 */
int orc_grid_new(const double bmin[3], const double bmax[3], double cell_size, orc_grid** out) {
  orc_grid* g = (orc_grid*)calloc(1, sizeof(orc_grid));
  if (!g) return ORC_ERR_IO;
  double ncells[3];
  uint64_t bitsum = 0;
  for (int i = 0; i < 3; ++i) {
    g->bmin[i] = bmin[i];
    g->bmax[i] = bmax[i];
    double extent = bmax[i] - bmin[i];           /* :19-23 */
    ncells[i] = ceil(extent / cell_size);        /* :24-28 */
    g->bits[i] = orc_f64_as_u64(ceil(log2(ncells[i]))); /* :29-31 */
    bitsum += g->bits[i];                        /* release build: wrapping add */
  }
  if (bitsum > 64) { /* :32-34 */
    free(g);
    return ORC_ERR_GRID;
  }
  g->cell_size = cell_size;
  for (int i = 0; i < 3; ++i) g->dims[i] = orc_f64_as_u64(ncells[i]); /* :39-43 */
  if (grid_alloc(g, 1024) != ORC_OK) {
    orc_grid_free(g);
    return ORC_ERR_IO;
  }
  *out = g;
  return ORC_OK;
}

void orc_grid_free(orc_grid* g) {
  if (!g) return;
  free(g->keys);
  free(g->vals);
  free(g->used);
  free(g);
}

static size_t grid_find(const orc_grid* g, uint64_t key, int* found) {
  size_t mask = g->cap - 1;
  size_t i = (size_t)mix64(key) & mask;
  while (g->used[i]) {
    if (g->keys[i] == key) {
      *found = 1;
      return i;
    }
    i = (i + 1) & mask;
  }
  *found = 0;
  return i;
}

static void grid_grow(orc_grid* g) {
  orc_grid old = *g;
  grid_alloc(g, old.cap * 2);
  for (size_t i = 0; i < old.cap; ++i) {
    if (!old.used[i]) continue;
    int found;
    size_t j = grid_find(g, old.keys[i], &found);
    g->used[j] = 1;
    g->keys[j] = old.keys[i];
    g->vals[j] = old.vals[i];
    g->len++;
  }
  free(old.keys);
  free(old.vals);
  free(old.used);
}

static void grid_cell(const orc_grid* g, const double pos[3], uint64_t cell[3]) {
  /* :51-60 */
  for (int i = 0; i < 3; ++i) {
    double r = (pos[i] - g->bmin[i]) * (double)g->dims[i] / (g->bmax[i] - g->bmin[i]);
    cell[i] = orc_f64_as_u64(r);
  }
}

static uint64_t bit_mask(uint64_t bits) {
  /* ((1 as u64) << bits) - 1; bits <= 64 here; 1<<64 would overflow the shift in Rust (panic in
   * debug, masked shift in release) — unreachable for grids this path builds (bits per axis < 64
   * unless the extent/cell ratio is >= 2^63). */
  return bits >= 64 ? UINT64_MAX : (((uint64_t)1 << bits) - 1);
}

uint64_t orc_grid_key(const orc_grid* g, const double pos[3], int* aliased) {
  uint64_t cell[3];
  grid_cell(g, pos, cell);
  uint64_t mx = bit_mask(g->bits[0]), my = bit_mask(g->bits[1]), mz = bit_mask(g->bits[2]);
  if (aliased) *aliased = (cell[0] > mx) || (cell[1] > my) || (cell[2] > mz);
  uint64_t ys = g->bits[0], zs = g->bits[0] + g->bits[1];
  uint64_t ky = ys >= 64 ? 0 : ((cell[1] & my) << ys);
  uint64_t kz = zs >= 64 ? 0 : ((cell[2] & mz) << zs);
  return (cell[0] & mx) | ky | kz; /* :62-70 */
}

static double dist2(const double a[3], const double b[3]) {
  /* nalgebra 0.23 distance_squared = (a-b).norm_squared() = (dx*dx + dy*dy) + dz*dz, each op rounded */
  double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  double xx = dx * dx, yy = dy * dy, zz = dz * dz;
  double s = xx + yy;
  return s + zz;
}

/* grid_sampling.rs:49-105 */
int orc_grid_insert_point(orc_grid* g, const orc_point* p) {
  double pos[3] = {p->pos[0], p->pos[1], p->pos[2]};
  uint64_t cell[3];
  grid_cell(g, pos, cell);
  uint64_t key = orc_grid_key(g, pos, NULL);

  if ((g->len + 1) * 2 > g->cap) grid_grow(g);
  int found;
  size_t slot = grid_find(g, key, &found);
  if (!found) { /* :73-76 */
    g->used[slot] = 1;
    g->keys[slot] = key;
    g->vals[slot] = *p;
    g->len++;
    return 1;
  }
  /* :77-103 — centre from the UNMASKED cell of the incoming point */
  double centre[3];
  for (int i = 0; i < 3; ++i) {
    double c = (double)cell[i] + 0.5;
    c = c * g->cell_size;
    centre[i] = c + g->bmin[i];
  }
  double cur[3] = {g->vals[slot].pos[0], g->vals[slot].pos[1], g->vals[slot].pos[2]};
  double cur_d = dist2(centre, cur);
  double new_d = dist2(centre, pos);
  if (new_d < cur_d) {
    g->vals[slot] = *p;
    return 1;
  }
  return 0;
}

size_t orc_grid_len(const orc_grid* g) { return g->len; }

size_t orc_grid_cells(const orc_grid* g, uint64_t* keys_out, size_t cap) {
  size_t k = 0;
  for (size_t i = 0; i < g->cap && k < cap; ++i)
    if (g->used[i]) keys_out[k++] = g->keys[i];
  return k;
}

size_t orc_grid_points(const orc_grid* g, orc_point* out, size_t cap) {
  size_t k = 0;
  for (size_t i = 0; i < g->cap && k < cap; ++i)
    if (g->used[i]) out[k++] = g->vals[i];
  return k;
}

/* ---------------------------------------------------------------------------------------------
 * Collectors (collect_points.rs)
 * ------------------------------------------------------------------------------------------- */
int orc_collector_new(int kind, const double gmin[3], const double gmax[3], double cell, orc_collector** out) {
  orc_collector* c = (orc_collector*)calloc(1, sizeof(orc_collector));
  if (!c) return ORC_ERR_IO;
  c->kind = kind;
  if (kind == ORC_COLLECT_GRID) {
    int rc = orc_grid_new(gmin, gmax, cell, &c->grid); /* collect_points.rs:104-108 */
    if (rc != ORC_OK) {
      free(c);
      return rc;
    }
  }
  *out = c;
  return ORC_OK;
}

void orc_collector_free(orc_collector* c) {
  if (!c) return;
  free(c->buf);
  orc_grid_free(c->grid);
  free(c);
}

void orc_collect_one(orc_collector* c, const orc_point* p) {
  switch (c->kind) {
    case ORC_COLLECT_COUNT: /* collect_points.rs:84-86 */
      c->count += 1;
      break;
    case ORC_COLLECT_BUFFER: /* :29-31 */
      if (c->buf_len == c->buf_cap) {
        size_t ncap = c->buf_cap ? c->buf_cap * 2 : 1024;
        c->buf = (orc_point*)realloc(c->buf, ncap * sizeof(orc_point));
        c->buf_cap = ncap;
      }
      c->buf[c->buf_len++] = *p;
      break;
    case ORC_COLLECT_GRID: /* :112-114 */
      orc_grid_insert_point(c->grid, p);
      break;
  }
}

size_t orc_collector_point_count(const orc_collector* c) {
  switch (c->kind) {
    case ORC_COLLECT_COUNT: return c->count;
    case ORC_COLLECT_BUFFER: return c->buf_len;
    case ORC_COLLECT_GRID: return orc_grid_len(c->grid);
  }
  return 0;
}

size_t orc_collector_points(const orc_collector* c, orc_point* out, size_t cap) {
  switch (c->kind) {
    case ORC_COLLECT_COUNT: return 0; /* None */
    case ORC_COLLECT_BUFFER: {
      size_t k = c->buf_len < cap ? c->buf_len : cap;
      memcpy(out, c->buf, k * sizeof(orc_point));
      return k;
    }
    case ORC_COLLECT_GRID: return orc_grid_points(c->grid, out, cap);
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * The four Optimized searches.  `file`/`n` is the mmap; every read is bounds-checked like
 * Cursor::read_exact (an out-of-range read is an io error that aborts the search).
 * ------------------------------------------------------------------------------------------- */
#define NEED(off, len)                                   \
  do {                                                   \
    if ((uint64_t)(off) + (uint64_t)(len) > (uint64_t)n) \
      return ORC_ERR_IO;                                 \
  } while (0)

static int color_offset_of(uint8_t format) { /* las.rs:38-45, 214-219; last.rs:83-88, 238-243 */
  switch (format) {
    case 2: return 20;
    case 3: return 28;
    case 5: return 28;
    default: return -1;
  }
}

/* las.rs:52-148 */
int orc_search_las_file_by_bounds_optimized(const uint8_t* file, size_t n, const double qmin[3],
                                            const double qmax[3], orc_collector* c) {
  orc_header h;
  int rc = orc_parse_header(file, n, 0, &h); /* :59-60 */
  if (rc != ORC_OK) return rc;
  if ((rc = aabb_check(h.min, h.max)) != ORC_OK) return rc; /* :61-72 */
  /* :73 println!("Point record size: {}") — stdout side effect, not part of the result */
  int color_off = color_offset_of(h.format); /* :74-80 */
  if (!aabb_intersects(h.min, h.max, qmin, qmax)) return ORC_OK; /* :82-84 */

  int64_t lo[3], hi[3];
  if ((rc = orc_local_bounds(&h, qmin, qmax, lo, hi)) != ORC_OK) return rc; /* :88-99 */

  for (uint64_t idx = 0; idx < h.n_points; ++idx) { /* :101 */
    uint64_t off = idx * (uint64_t)h.record_len + (uint64_t)h.offset_to_point_data; /* :102-104 */
    NEED(off, 4);
    int64_t px = (int64_t)rd_i32(file + off); /* :106-109 */
    if (px < lo[0] || px > hi[0]) continue;
    NEED(off + 4, 4);
    int64_t py = (int64_t)rd_i32(file + off + 4); /* :111-114 */
    if (py < lo[1] || py > hi[1]) continue;
    NEED(off + 8, 4);
    int64_t pz = (int64_t)rd_i32(file + off + 8); /* :116-119 */
    if (pz < lo[2] || pz > hi[2]) continue;

    NEED(off + 15, 1);
    uint8_t cls = file[off + 15]; /* :121-124: seek +3 from byte 12 */
    orc_point p;
    if (color_off >= 0) { /* :127-135: seek (color_off-16) from byte 16 */
      NEED(off + (uint64_t)color_off, 6);
      p.rgb[0] = rd_u16(file + off + color_off);
      p.rgb[1] = rd_u16(file + off + color_off + 2);
      p.rgb[2] = rd_u16(file + off + color_off + 4);
    } else {
      p.rgb[0] = p.rgb[1] = p.rgb[2] = 0;
    }
    /* :137-145 — multiply, then add; never fused */
    double mx = (double)px * h.scale[0];
    double my = (double)py * h.scale[1];
    double mz = (double)pz * h.scale[2];
    p.pos[0] = mx + h.offset[0];
    p.pos[1] = my + h.offset[1];
    p.pos[2] = mz + h.offset[2];
    p.cls = cls;
    orc_collect_one(c, &p);
  }
  return ORC_OK;
}

/* las.rs:192-261 */
int orc_search_las_file_by_classification_optimized(const uint8_t* file, size_t n, uint8_t cls,
                                                    orc_collector* c) {
  orc_header h;
  int rc = orc_parse_header(file, n, 0, &h); /* :199-200 */
  if (rc != ORC_OK) return rc;
  uint64_t cls_off = h.format <= 5 ? 15 : 16; /* :202-212 (format > 10 already rejected) */
  int color_off = color_offset_of(h.format);  /* :214-219 */

  for (uint64_t idx = 0; idx < h.n_points; ++idx) { /* :221 */
    uint64_t off = idx * (uint64_t)h.record_len + (uint64_t)h.offset_to_point_data;
    NEED(off + cls_off, 1);
    uint8_t classification = file[off + cls_off]; /* :224-231 */
    if (classification != cls) continue;
    NEED(off, 12);
    int32_t px = rd_i32(file + off), py = rd_i32(file + off + 4), pz = rd_i32(file + off + 8); /* :233-237 */
    orc_point p;
    if (color_off >= 0) { /* :240-248 */
      NEED(off + (uint64_t)color_off, 6);
      p.rgb[0] = rd_u16(file + off + color_off);
      p.rgb[1] = rd_u16(file + off + color_off + 2);
      p.rgb[2] = rd_u16(file + off + color_off + 4);
    } else {
      p.rgb[0] = p.rgb[1] = p.rgb[2] = 0;
    }
    double mx = (double)px * h.scale[0]; /* :250-258 */
    double my = (double)py * h.scale[1];
    double mz = (double)pz * h.scale[2];
    p.pos[0] = mx + h.offset[0];
    p.pos[1] = my + h.offset[1];
    p.pos[2] = mz + h.offset[2];
    p.cls = classification;
    orc_collect_one(c, &p);
  }
  return ORC_OK;
}

/* last.rs:46-166 */
int orc_search_last_file_by_bounds_optimized(const uint8_t* file, size_t n, const double qmin[3],
                                             const double qmax[3], orc_collector* c) {
  orc_header h;
  int rc = orc_parse_header(file, n, 0, &h); /* :53-54 — format byte NOT masked here */
  if (rc != ORC_OK) return rc;
  if ((rc = aabb_check(h.min, h.max)) != ORC_OK) return rc; /* :55-66 */
  uint64_t cls_in_point = h.format <= 5 ? 15 : 16;              /* :68-79 */
  uint64_t cls_block = (uint64_t)h.offset_to_point_data + h.n_points * cls_in_point; /* :80-81 */
  int color_in_point = color_offset_of(h.format);               /* :83-88 */
  uint64_t color_block = color_in_point >= 0
                             ? (uint64_t)h.offset_to_point_data + h.n_points * (uint64_t)color_in_point
                             : 0; /* :89-90 */
  if (!aabb_intersects(h.min, h.max, qmin, qmax)) return ORC_OK; /* :92-94 */

  int64_t lo[3], hi[3];
  if ((rc = orc_local_bounds(&h, qmin, qmax, lo, hi)) != ORC_OK) return rc; /* :98-109 */

  uint64_t pos_block = (uint64_t)h.offset_to_point_data; /* :114 */
  for (uint64_t idx = 0; idx < h.n_points; ++idx) {      /* :117 */
    uint64_t off = pos_block + idx * 12;                 /* :118-121 */
    NEED(off, 4);
    int64_t px = (int64_t)rd_i32(file + off); /* :122-125 */
    if (px < lo[0] || px > hi[0]) continue;
    NEED(off + 4, 4);
    int64_t py = (int64_t)rd_i32(file + off + 4); /* :127-130 */
    if (py < lo[1] || py > hi[1]) continue;
    NEED(off + 8, 4);
    int64_t pz = (int64_t)rd_i32(file + off + 8); /* :132-135 */
    if (pz < lo[2] || pz > hi[2]) continue;

    NEED(cls_block + idx, 1);
    uint8_t cls = file[cls_block + idx]; /* :137-142 */
    orc_point p;
    if (color_in_point >= 0) { /* :145-153 */
      uint64_t co = idx * 6 + color_block;
      NEED(co, 6);
      p.rgb[0] = rd_u16(file + co);
      p.rgb[1] = rd_u16(file + co + 2);
      p.rgb[2] = rd_u16(file + co + 4);
    } else {
      p.rgb[0] = p.rgb[1] = p.rgb[2] = 0;
    }
    double mx = (double)px * h.scale[0]; /* :155-163 */
    double my = (double)py * h.scale[1];
    double mz = (double)pz * h.scale[2];
    p.pos[0] = mx + h.offset[0];
    p.pos[1] = my + h.offset[1];
    p.pos[2] = mz + h.offset[2];
    p.cls = cls;
    orc_collect_one(c, &p);
  }
  return ORC_OK;
}

/* last.rs:213-293 */
int orc_search_last_file_by_classification_optimized(const uint8_t* file, size_t n, uint8_t cls,
                                                     orc_collector* c) {
  orc_header h;
  int rc = orc_parse_header(file, n, 1, &h); /* :220-223 — format &= 0b1111 */
  if (rc != ORC_OK) return rc;
  uint64_t cls_in_point = h.format <= 5 ? 15 : 16; /* :225-236 */
  int color_in_point = color_offset_of(h.format);  /* :238-243 */
  uint64_t cls_block = cls_in_point * h.n_points;  /* :245-246 */
  uint64_t color_block = color_in_point >= 0
                             ? (uint64_t)h.offset_to_point_data + h.n_points * (uint64_t)color_in_point
                             : 0; /* :249-250 */

  for (uint64_t idx = 0; idx < h.n_points; ++idx) { /* :253 */
    uint64_t co = idx + cls_block + (uint64_t)h.offset_to_point_data; /* :254-257 */
    NEED(co, 1);
    uint8_t classification = file[co]; /* :259-262 */
    if (classification != cls) continue;
    uint64_t po = idx * 12 + (uint64_t)h.offset_to_point_data; /* :265-269 */
    NEED(po, 12);
    int32_t px = rd_i32(file + po), py = rd_i32(file + po + 4), pz = rd_i32(file + po + 8);
    orc_point p;
    if (color_in_point >= 0) { /* :272-280 */
      uint64_t cc = idx * 6 + color_block;
      NEED(cc, 6);
      p.rgb[0] = rd_u16(file + cc);
      p.rgb[1] = rd_u16(file + cc + 2);
      p.rgb[2] = rd_u16(file + cc + 4);
    } else {
      p.rgb[0] = p.rgb[1] = p.rgb[2] = 0;
    }
    double mx = (double)px * h.scale[0]; /* :282-290 */
    double my = (double)py * h.scale[1];
    double mz = (double)pz * h.scale[2];
    p.pos[0] = mx + h.offset[0];
    p.pos[1] = my + h.offset[1];
    p.pos[2] = mz + h.offset[2];
    p.cls = classification;
    orc_collect_one(c, &p);
  }
  return ORC_OK;
}

/* searcher.rs:43-91, 104-152 (Optimized arms of "las" and "last"; other extensions are out of scope) */
int orc_search_file(const uint8_t* file, size_t n, const char* ext, int query_kind, const double qmin[3],
                    const double qmax[3], uint8_t cls, orc_collector* c) {
  if (strcmp(ext, "las") == 0) {
    return query_kind == 0 ? orc_search_las_file_by_bounds_optimized(file, n, qmin, qmax, c)
                           : orc_search_las_file_by_classification_optimized(file, n, cls, c);
  }
  if (strcmp(ext, "last") == 0) {
    return query_kind == 0 ? orc_search_last_file_by_bounds_optimized(file, n, qmin, qmax, c)
                           : orc_search_last_file_by_classification_optimized(file, n, cls, c);
  }
  return ORC_ERR_FORMAT; /* "Unsupported file extension" */
}

/* ---------------------------------------------------------------------------------------------
 * Chunk headers of the on-the-fly index the reference only describes (improvements.md:3-10).
 * Plain loops over the same fields the searches read: x, y, z i32 at +0/+4/+8 of a record (las.rs:106-119)
 * or of the positions column (last.rs:114-121); the class byte at +15 / +16 of a record (las.rs:202-212) or
 * in the class column at off + {15,16} * N (last.rs:245-259).
 * ------------------------------------------------------------------------------------------- */
int orc_chunk_headers(const uint8_t* file, size_t n, const char* ext, uint64_t first, uint64_t count,
                      uint32_t chunk_points, uint32_t* out) {
  const int is_las = strcmp(ext, "las") == 0;
  if (!is_las && strcmp(ext, "last") != 0) return ORC_ERR_FORMAT;
  if (chunk_points == 0) return ORC_ERR_FORMAT;
  orc_header h;
  int rc = orc_parse_header(file, n, 1, &h);
  if (rc != 0) return rc;
  const uint64_t N = h.n_points;
  if (first > N || count > N - first) return ORC_ERR_FORMAT;
  const uint64_t off = h.offset_to_point_data;
  const uint8_t fmt = (uint8_t)(h.format & 0x0F);
  const uint64_t cls_k = fmt <= 5 ? 15u : 16u;
  const uint64_t R = h.record_len;
  const uint64_t n_chunks = (count + chunk_points - 1) / chunk_points;
  for (uint64_t c = 0; c < n_chunks; ++c) {
    uint32_t* o = out + 16 * c;
    int32_t lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    uint32_t bits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint64_t a = first + c * chunk_points;
    const uint64_t b = a + chunk_points < first + count ? a + chunk_points : first + count;
    for (uint64_t i = a; i < b; ++i) {
      const uint8_t* p = is_las ? file + off + i * R : file + off + i * 12u;
      const uint8_t k = is_las ? file[off + i * R + cls_k] : file[off + cls_k * N + i];
      for (int ax = 0; ax < 3; ++ax) {
        const int32_t v = rd_i32(p + 4 * ax);
        if (v < lo[ax]) lo[ax] = v;
        if (v > hi[ax]) hi[ax] = v;
      }
      bits[k >> 5] |= 1u << (k & 31u);
    }
    for (int ax = 0; ax < 3; ++ax) {
      o[ax] = (uint32_t)lo[ax];
      o[3 + ax] = (uint32_t)hi[ax];
    }
    for (int j = 0; j < 8; ++j) o[6 + j] = bits[j];
    o[14] = (uint32_t)(b - a);
    o[15] = 0;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * run_search_parallel with CountCollector (main.rs:146-183): files.par_iter() — one task per
 * file, a fresh collector per file, counts summed by the caller.
 * ------------------------------------------------------------------------------------------- */
typedef struct par_job {
  const uint8_t* const* files;
  const size_t* sizes;
  const char* const* exts;
  size_t n_files;
  int query_kind;
  const double* qmin;
  const double* qmax;
  uint8_t cls;
  uint64_t* counts;
  size_t next;
  int err;
  pthread_mutex_t mu;
} par_job;

static void* par_worker(void* arg) {
  par_job* j = (par_job*)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    size_t i = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (i >= j->n_files) break;
    orc_collector* c = NULL;
    int rc = orc_collector_new(ORC_COLLECT_COUNT, NULL, NULL, 0.0, &c); /* main.rs:156 */
    if (rc == ORC_OK) {
      rc = orc_search_file(j->files[i], j->sizes[i], j->exts[i], j->query_kind, j->qmin, j->qmax, j->cls, c);
      j->counts[i] = (uint64_t)orc_collector_point_count(c);
      orc_collector_free(c);
    }
    if (rc != ORC_OK) {
      pthread_mutex_lock(&j->mu);
      if (j->err == ORC_OK) j->err = rc;
      pthread_mutex_unlock(&j->mu);
    }
  }
  return NULL;
}

int orc_count_parallel(const uint8_t* const* files, const size_t* sizes, const char* const* exts,
                       size_t n_files, int query_kind, const double qmin[3], const double qmax[3],
                       uint8_t cls, int n_threads, uint64_t* per_file_counts) {
  par_job j;
  j.files = files;
  j.sizes = sizes;
  j.exts = exts;
  j.n_files = n_files;
  j.query_kind = query_kind;
  j.qmin = qmin;
  j.qmax = qmax;
  j.cls = cls;
  j.counts = per_file_counts;
  j.next = 0;
  j.err = ORC_OK;
  pthread_mutex_init(&j.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > n_files) n_threads = (int)(n_files ? n_files : 1);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, par_worker, &j);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&j.mu);
  return j.err;
}
