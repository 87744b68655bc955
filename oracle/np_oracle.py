"""Independent numpy / pure-Python restatement of the same reference path (TEST INFRASTRUCTURE ONLY).

A second, differently-written statement of query/src/search/las.rs:52-148, 192-261,
query/src/search/last.rs:46-166, 213-293 and query/src/grid_sampling.rs:18-105, used to cross-check
the C oracle (oracle/pcq_oracle.c) since the reference itself cannot be built here.  numpy float64
arithmetic is IEEE and unfused, like Rust's.
"""
from __future__ import annotations

import math
import struct

import numpy as np

POINT_DTYPE = np.dtype([("pos", "<f8", (3,)), ("rgb", "<u2", (3,)), ("cls", "u1")])

I64_MAX, I64_MIN = (1 << 63) - 1, -(1 << 63)
U64_MAX = (1 << 64) - 1


class Panic(Exception):
    """a place where the reference panics"""


class FormatError(Exception):
    """an `Err(anyhow!(..))` of the reference"""


def f64_as_i64(v: float) -> int:
    if math.isnan(v):
        return 0
    if v >= 9223372036854775808.0:
        return I64_MAX
    if v <= -9223372036854775808.0:
        return I64_MIN
    return int(v)  # truncates toward zero


def f64_as_u64(v: float) -> int:
    if math.isnan(v) or v <= 0.0:
        return 0
    if v >= 18446744073709551616.0:
        return U64_MAX
    return int(v)


def parse_header(buf: np.ndarray, mask_format: bool = False) -> dict:
    b = bytes(buf[:375])
    if len(b) < 227:
        raise IOError("short header")
    if b[:4] != b"LASF":
        raise FormatError("signature")
    h = {
        "version": (b[24], b[25]),
        "header_size": struct.unpack_from("<H", b, 94)[0],
        "off": struct.unpack_from("<I", b, 96)[0],
        "format": b[104],
        "record_len": struct.unpack_from("<H", b, 105)[0],
        "legacy": struct.unpack_from("<I", b, 107)[0],
        "scale": struct.unpack_from("<3d", b, 131),
        "offset": struct.unpack_from("<3d", b, 155),
    }
    mx_x, mn_x, mx_y, mn_y, mx_z, mn_z = struct.unpack_from("<6d", b, 179)
    h["min"], h["max"] = (mn_x, mn_y, mn_z), (mx_x, mx_y, mx_z)
    if mask_format:
        h["format"] &= 0b1111
    if h["format"] > 10:
        raise FormatError("format")
    flen = (20, 28, 26, 34, 57, 63, 30, 36, 38, 59, 67)[h["format"]]
    if h["record_len"] < flen:
        raise FormatError("record length")
    v14 = h["version"] >= (1, 4)
    if h["format"] >= 6 and not v14:
        raise FormatError("extended format needs 1.4")
    h["n"] = h["legacy"] if h["legacy"] > 0 else (struct.unpack_from("<Q", b, 247)[0] if v14 else 0)
    return h


def local_bounds(h: dict, qmin, qmax):
    sx, sy, sz = h["scale"]
    ox, oy, oz = h["offset"]
    lo = [f64_as_i64((qmin[0] - ox) / sx), f64_as_i64((qmin[1] - oy) / sx), f64_as_i64((qmin[2] - oz) / sx)]
    hi = [f64_as_i64((qmax[0] - ox) / sx), f64_as_i64((qmax[1] - oy) / sy), f64_as_i64((qmax[2] - oz) / sz)]
    if any(a > b for a, b in zip(lo, hi)):
        raise Panic("local bounds inverted")
    return lo, hi


def _intersects(h, qmin, qmax) -> bool:
    if any(a > b for a, b in zip(h["min"], h["max"])):
        raise Panic("header bounds inverted")
    return all(h["min"][i] <= qmax[i] and h["max"][i] >= qmin[i] for i in range(3))


def _color_off(fmt):
    return {2: 20, 3: 28, 5: 28}.get(fmt)


def _columns(buf: np.ndarray, h: dict, layout: str):
    """-> xyz int32 (N,3), class u8 (N,), rgb u16 (N,3) or None — views decoded from the file image"""
    n, off, R, fmt = h["n"], h["off"], h["record_len"], h["format"]
    cls_k = 15 if fmt <= 5 else 16
    col_k = _color_off(fmt)
    if layout == "las":
        rec = buf[off: off + n * R].reshape(n, R)
        xyz = np.ascontiguousarray(rec[:, 0:12]).view("<i4").reshape(n, 3)
        cls = rec[:, cls_k].copy()
        rgb = np.ascontiguousarray(rec[:, col_k: col_k + 6]).view("<u2").reshape(n, 3) if col_k is not None else None
        return xyz, cls, rgb
    xyz = np.ascontiguousarray(buf[off: off + 12 * n]).view("<i4").reshape(n, 3)
    cls = buf[off + cls_k * n: off + cls_k * n + n].copy()
    rgb = None
    if col_k is not None:
        rgb = np.ascontiguousarray(buf[off + col_k * n: off + col_k * n + 6 * n]).view("<u2").reshape(n, 3)
    return xyz, cls, rgb


def _emit(h, xyz, cls, rgb, sel) -> np.ndarray:
    out = np.zeros(int(sel.sum()), dtype=POINT_DTYPE)
    v = xyz[sel].astype(np.float64)
    for a in range(3):
        out["pos"][:, a] = v[:, a] * np.float64(h["scale"][a]) + np.float64(h["offset"][a])  # two roundings
    out["cls"] = cls[sel]
    if rgb is not None:
        out["rgb"] = rgb[sel]
    return out


def search_bounds(buf: np.ndarray, layout: str, qmin, qmax) -> np.ndarray:
    """search_{las,last}_file_by_bounds_optimized -> matching Points in scan order"""
    h = parse_header(buf, mask_format=False)
    if not _intersects(h, qmin, qmax):
        return np.zeros(0, dtype=POINT_DTYPE)
    lo, hi = local_bounds(h, qmin, qmax)
    xyz, cls, rgb = _columns(buf, h, layout)
    if layout == "las":
        cls = buf[h["off"]: h["off"] + h["n"] * h["record_len"]].reshape(h["n"], h["record_len"])[:, 15].copy()  # las.rs:121-124
    x = xyz.astype(np.int64)
    sel = np.ones(h["n"], dtype=bool)
    for a in range(3):
        sel &= (x[:, a] >= lo[a]) & (x[:, a] <= hi[a])
    return _emit(h, xyz, cls, rgb, sel)


def search_class(buf: np.ndarray, layout: str, klass: int) -> np.ndarray:
    """search_{las,last}_file_by_classification_optimized -> matching Points in scan order"""
    h = parse_header(buf, mask_format=(layout == "last"))
    xyz, cls, rgb = _columns(buf, h, layout)
    sel = cls == np.uint8(klass)
    return _emit(h, xyz, cls, rgb, sel)


CHUNK_HEADER_DTYPE = np.dtype(
    [("lo", "<i4", (3,)), ("hi", "<i4", (3,)), ("cls_bits", "<u4", (8,)), ("n_points", "<u4"), ("pad", "<u4")]
)


def chunk_headers(buf: np.ndarray, layout: str, chunk_points: int = 8192, first: int = 0, count=None) -> np.ndarray:
    """The chunk headers improvements.md:3-10 proposes (the reference does not implement them): per chunk of
    consecutive points the min/max of the raw x/y/z fields and the set of the class bytes the class search compares
    (las.rs:202-212, last.rs:245-259).  Checker of pcq_file_build_index."""
    h = parse_header(buf, mask_format=True)
    xyz, cls, _ = _columns(buf, h, layout)
    n = h["n"] - first if count is None else count
    xyz, cls = xyz[first: first + n], cls[first: first + n]
    out = np.zeros((n + chunk_points - 1) // chunk_points, dtype=CHUNK_HEADER_DTYPE)
    for c in range(out.shape[0]):
        a, b = c * chunk_points, min((c + 1) * chunk_points, n)
        out["lo"][c] = xyz[a:b].min(axis=0)
        out["hi"][c] = xyz[a:b].max(axis=0)
        for v in np.unique(cls[a:b]):
            out["cls_bits"][c, int(v) >> 5] |= np.uint32(1 << (int(v) & 31))
        out["n_points"][c] = b - a
    return out


class SparseGrid:
    """grid_sampling.rs:9-105 with a dict as the HashMap"""

    def __init__(self, bmin, bmax, cell_size: float):
        self.bmin, self.bmax, self.cell = tuple(map(float, bmin)), tuple(map(float, bmax)), float(cell_size)
        self.dims, self.bits = [], []
        for a in range(3):
            extent = self.bmax[a] - self.bmin[a]
            with np.errstate(divide="ignore", invalid="ignore"):
                ncells = float(np.ceil(np.float64(extent) / np.float64(self.cell)))
                lg = float(np.ceil(np.log2(np.float64(ncells))))
            self.bits.append(f64_as_u64(lg))
            self.dims.append(f64_as_u64(ncells))
        if sum(self.bits) > 64:
            raise ValueError("Too many cells")
        self.cells = {}

    def _cell(self, pos):
        out = []
        for a in range(3):
            with np.errstate(divide="ignore", invalid="ignore"):
                r = (np.float64(pos[a]) - np.float64(self.bmin[a])) * np.float64(float(self.dims[a])) / (
                    np.float64(self.bmax[a]) - np.float64(self.bmin[a]))
            out.append(f64_as_u64(float(r)))
        return out

    def key_of(self, pos):
        c = self._cell(pos)
        m = [(1 << b) - 1 for b in self.bits]
        key = (c[0] & m[0]) | ((c[1] & m[1]) << self.bits[0]) | ((c[2] & m[2]) << (self.bits[0] + self.bits[1]))
        return key & U64_MAX, any(c[a] > m[a] for a in range(3))

    def insert_point(self, point) -> bool:
        pos = [float(point["pos"][a]) for a in range(3)]
        c = self._cell(pos)
        key, _ = self.key_of(pos)
        cur = self.cells.get(key)
        if cur is None:
            self.cells[key] = point
            return True
        centre = [np.float64(float(c[a]) + 0.5) * np.float64(self.cell) + np.float64(self.bmin[a]) for a in range(3)]

        def d2(p):
            d = [centre[a] - np.float64(float(p["pos"][a])) for a in range(3)]
            return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]

        if d2(point) < d2(cur):
            self.cells[key] = point
            return True
        return False

    def points(self) -> np.ndarray:
        out = np.zeros(len(self.cells), dtype=POINT_DTYPE)
        for i, p in enumerate(self.cells.values()):
            out[i] = p
        return out


# ---------------------------------------------------------------------------------------------------
# FileDumper::dump_points (query/src/dump_points.rs:63-116) — what `-o` writes.  TEST INFRASTRUCTURE ONLY.
# ---------------------------------------------------------------------------------------------------
def dump_points_plan(buffers):
    """The files FileDumper writes for a sequence of point buffers (one per collector, in the order main.rs hands
    them over): -> list of dicts {index, offset, scale, raw, cls, rgb}.

      :65-67   an empty buffer writes nothing and does NOT consume a file index
      :68-71   matching_points_{file_index}.las, file_index += 1
      :74-80   min / max position over the buffer; offset = min position
      :81-88   extent = max - min; max_extent = its largest component; min_scale = max_extent / i32::MAX;
               scale = 10 ^ ceil(log10(min_scale)); `if scale < 0.001 { scale = 0.001 }` (a zero extent gives
               log10(0) = -inf -> 10^-inf = 0 -> clamped; NaN compares false and would stay NaN)
      :90-106  LAS 1.2, point format 2, the same offset / scale on every axis
    Record quantisation lives in pasture-io's LASWriter (un-vendored, parity unpinned): assumed to be
    las-rs' Transform::inverse, round((p - offset) / scale) as i32.
    """
    out = []
    index = 0
    for pts in buffers:
        pts = np.asarray(pts, dtype=POINT_DTYPE)
        if len(pts) == 0:
            continue
        pos = pts["pos"].astype(np.float64)
        mn, mx = pos.min(axis=0), pos.max(axis=0)
        max_extent = float((mx - mn).max())
        min_scale = max_extent / 2147483647.0
        with np.errstate(divide="ignore"):
            scale = float(np.power(10.0, np.ceil(np.log10(min_scale)))) if min_scale > 0 else 0.0
        if scale < 0.001:
            scale = 0.001
        # f64::round rounds half AWAY from zero (np.rint would round half to even); the quotients are >= 0
        q = (pos - mn) / scale
        fl = np.floor(q)
        raw = np.where(q - fl >= 0.5, fl + 1.0, fl)
        out.append({"index": index, "offset": mn, "scale": scale, "raw": raw.astype(np.int64).astype(np.int32),
                    "cls": pts["cls"].copy(), "rgb": pts["rgb"].copy()})
        index += 1
    return out
