"""ctypes binding of oracle/liboracle.so — the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  See pcq_oracle.h for what the oracle restates and for its parity status ("parity
unpinned" for the bbox / class scans; pinned by the reference's three SparseGrid tests for the
density grid).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liboracle.so"

POINT_DTYPE = np.dtype([("pos", "<f8", (3,)), ("rgb", "<u2", (3,)), ("cls", "u1")])
assert POINT_DTYPE.itemsize == 31

ORC_OK, ORC_ERR_IO, ORC_ERR_FORMAT, ORC_ERR_PANIC, ORC_ERR_GRID = 0, -1, -2, -3, -6
COLLECT_COUNT, COLLECT_BUFFER, COLLECT_GRID = 0, 1, 2
_D3 = C.c_double * 3


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", str(_HERE), "liboracle.so"])


class Header(C.Structure):
    _fields_ = [
        ("version_major", C.c_uint8), ("version_minor", C.c_uint8), ("header_size", C.c_uint16),
        ("offset_to_point_data", C.c_uint32), ("n_vlrs", C.c_uint32), ("format", C.c_uint8),
        ("record_len", C.c_uint16), ("legacy_count", C.c_uint32),
        ("scale", _D3), ("offset", _D3), ("max", _D3), ("min", _D3),
        ("large_count", C.c_uint64), ("has_large", C.c_int), ("n_points", C.c_uint64),
    ]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        build()
    lib = C.CDLL(str(LIB_PATH))
    vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
    P = C.POINTER
    sig = {
        "orc_f64_as_i64": (C.c_int64, [C.c_double]),
        "orc_f64_as_u64": (C.c_uint64, [C.c_double]),
        "orc_parse_header": (C.c_int, [vp, sz, C.c_int, P(Header)]),
        "orc_local_bounds": (C.c_int, [P(Header), _D3, _D3, C.c_int64 * 3, C.c_int64 * 3]),
        "orc_grid_new": (C.c_int, [_D3, _D3, C.c_double, P(vp)]),
        "orc_grid_free": (None, [vp]),
        "orc_grid_insert_point": (C.c_int, [vp, vp]),
        "orc_grid_len": (sz, [vp]),
        "orc_grid_cells": (sz, [vp, vp, sz]),
        "orc_grid_points": (sz, [vp, vp, sz]),
        "orc_grid_key": (u64, [vp, _D3, P(C.c_int)]),
        "orc_collector_new": (C.c_int, [C.c_int, vp, vp, C.c_double, P(vp)]),
        "orc_collector_free": (None, [vp]),
        "orc_collect_one": (None, [vp, vp]),
        "orc_collector_point_count": (sz, [vp]),
        "orc_collector_points": (sz, [vp, vp, sz]),
        "orc_search_file": (C.c_int, [vp, sz, C.c_char_p, C.c_int, _D3, _D3, C.c_uint8, vp]),
        "orc_chunk_headers": (C.c_int, [vp, sz, C.c_char_p, u64, u64, C.c_uint32, vp]),
        "orc_count_parallel": (C.c_int, [P(vp), P(sz), P(C.c_char_p), sz, C.c_int, _D3, _D3, C.c_uint8, C.c_int, P(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


def d3(v):
    return _D3(float(v[0]), float(v[1]), float(v[2]))


def parse_header(file_bytes: np.ndarray, mask_format: bool = False) -> Header:
    h = Header()
    rc = lib.orc_parse_header(C.c_void_p(file_bytes.ctypes.data), file_bytes.nbytes, int(mask_format), C.byref(h))
    if rc != ORC_OK:
        raise OracleError(rc)
    return h


def local_bounds(h: Header, qmin, qmax):
    lo, hi = (C.c_int64 * 3)(), (C.c_int64 * 3)()
    rc = lib.orc_local_bounds(C.byref(h), d3(qmin), d3(qmax), lo, hi)
    if rc != ORC_OK:
        raise OracleError(rc)
    return list(lo), list(hi)


class Grid:
    """SparseGrid (grid_sampling.rs)."""

    def __init__(self, bmin, bmax, cell_size: float):
        h = C.c_void_p()
        rc = lib.orc_grid_new(d3(bmin), d3(bmax), float(cell_size), C.byref(h))
        if rc != ORC_OK:
            raise OracleError(rc)
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_grid_free(self.h)
            self.h = None

    def insert_point(self, pos, rgb=(0, 0, 0), cls=0) -> bool:
        p = np.zeros(1, dtype=POINT_DTYPE)
        p["pos"][0] = pos
        p["rgb"][0] = rgb
        p["cls"][0] = cls
        return bool(lib.orc_grid_insert_point(self.h, C.c_void_p(p.ctypes.data)))

    def cells(self) -> np.ndarray:
        n = lib.orc_grid_len(self.h)
        out = np.zeros(n, dtype=np.uint64)
        lib.orc_grid_cells(self.h, C.c_void_p(out.ctypes.data), n)
        return out

    def points(self) -> np.ndarray:
        n = lib.orc_grid_len(self.h)
        out = np.zeros(n, dtype=POINT_DTYPE)
        lib.orc_grid_points(self.h, C.c_void_p(out.ctypes.data), n)
        return out

    def key(self, pos):
        al = C.c_int()
        k = lib.orc_grid_key(self.h, d3(pos), C.byref(al))
        return int(k), bool(al.value)

    def dims_bits(self):
        class G(C.Structure):
            _fields_ = [("bmin", _D3), ("bmax", _D3), ("cell", C.c_double), ("dims", C.c_uint64 * 3), ("bits", C.c_uint64 * 3)]

        g = G.from_address(self.h.value)
        return list(g.dims), list(g.bits)


class Collector:
    """CountCollector / BufferCollector / GridSampledCollector (collect_points.rs)."""

    def __init__(self, kind: int, gmin=None, gmax=None, cell: float = 0.0):
        h = C.c_void_p()
        a = d3(gmin) if gmin is not None else None
        b = d3(gmax) if gmax is not None else None
        rc = lib.orc_collector_new(kind, a, b, float(cell), C.byref(h))
        if rc != ORC_OK:
            raise OracleError(rc)
        self.h = h
        self.kind = kind

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_collector_free(self.h)
            self.h = None

    def point_count(self) -> int:
        return int(lib.orc_collector_point_count(self.h))

    def points(self):
        if self.kind == COLLECT_COUNT:
            return None
        n = self.point_count()
        out = np.zeros(n, dtype=POINT_DTYPE)
        got = lib.orc_collector_points(self.h, C.c_void_p(out.ctypes.data), n)
        return out[:got]


def search_file(file_bytes: np.ndarray, ext: str, collector: Collector, bounds=None, cls=None) -> None:
    """Searcher::search_file with SearchImplementation::Optimized."""
    assert (bounds is None) != (cls is None)
    if bounds is not None:
        rc = lib.orc_search_file(C.c_void_p(file_bytes.ctypes.data), file_bytes.nbytes, ext.encode(), 0, d3(bounds[0]),
                                 d3(bounds[1]), 0, collector.h)
    else:
        z = d3((0, 0, 0))
        rc = lib.orc_search_file(C.c_void_p(file_bytes.ctypes.data), file_bytes.nbytes, ext.encode(), 1, z, z, int(cls), collector.h)
    if rc != ORC_OK:
        raise OracleError(rc)


def count_parallel(files, exts, n_threads: int, bounds=None, cls=None) -> np.ndarray:
    """run_search_parallel with CountCollectors: per-file match counts."""
    n = len(files)
    ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in files])
    sizes = (C.c_size_t * n)(*[f.nbytes for f in files])
    ex = (C.c_char_p * n)(*[e.encode() for e in exts])
    out = (C.c_uint64 * n)()
    if bounds is not None:
        rc = lib.orc_count_parallel(ptrs, sizes, ex, n, 0, d3(bounds[0]), d3(bounds[1]), 0, n_threads, out)
    else:
        z = d3((0, 0, 0))
        rc = lib.orc_count_parallel(ptrs, sizes, ex, n, 1, z, z, int(cls), n_threads, out)
    if rc != ORC_OK:
        raise OracleError(rc)
    return np.array(list(out), dtype=np.uint64)


CHUNK_HEADER_DTYPE = np.dtype(
    [("lo", "<i4", (3,)), ("hi", "<i4", (3,)), ("cls_bits", "<u4", (8,)), ("n_points", "<u4"), ("pad", "<u4")]
)


def chunk_headers(file_bytes: np.ndarray, ext: str, chunk_points: int = 8192, first: int = 0, count=None) -> np.ndarray:
    """Chunk headers of the on-the-fly index (improvements.md:3-10), C restatement (orc_chunk_headers)."""
    if count is None:
        count = int(parse_header(file_bytes, True).n_points) - first
    n_chunks = (count + chunk_points - 1) // chunk_points
    out = np.zeros(n_chunks, dtype=CHUNK_HEADER_DTYPE)
    rc = lib.orc_chunk_headers(C.c_void_p(file_bytes.ctypes.data), file_bytes.nbytes, ext.encode(), int(first), int(count),
                               int(chunk_points), C.c_void_p(out.ctypes.data))
    if rc != ORC_OK:
        raise OracleError(rc)
    return out
