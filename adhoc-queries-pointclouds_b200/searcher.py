"""Host-side mirror of the reference's operator interface for the scan path.

Same names, argument meaning and error behaviour as
  query/src/search/searcher.rs   SearchImplementation, Searcher, BoundsSearcher, ClassSearcher
  query/src/collect_points.rs    ResultCollector, CountCollector, BufferCollector, GridSampledCollector
  query/src/search/las.rs:52,192 search_las_file_by_{bounds,classification}_optimized
  query/src/search/last.rs:46,213 search_last_file_by_{bounds,classification}_optimized
  query/src/main.rs:122-183      run_search_sequential / run_search_parallel
so that the parity tests read like the reference's call sites.  All point work happens in libpcq's
CUDA kernels; this module only marshals arguments (it is the Python twin of the C++ mirror in
csrc/pcq_host.hpp, which the `query` CLI uses).
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from pathlib import Path
from typing import Callable, Iterable, Optional, Sequence, Union

import numpy as np

from . import binding as B
from .binding import PcqError, check, lib


class Context:
    """One GPU + stream (pcq_ctx).  One per process / rank."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib.pcq_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if self.handle:
            lib.pcq_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(lib.pcq_ctx_synchronize(self.handle))

    def set_stream(self, cuda_stream: int):
        check(lib.pcq_ctx_set_stream(self.handle, C.c_void_p(int(cuda_stream))))

    def set_scan_variant(self, variant: int):
        check(lib.pcq_ctx_set_scan_variant(self.handle, int(variant)))

    @property
    def launch_count(self) -> int:
        return int(lib.pcq_ctx_launch_count(self.handle))

    def set_auto_index(self, after_n_scans: int):
        """Build a resident file's chunk index when it is scanned for the (n+1)-th time (0 = never)."""
        check(lib.pcq_ctx_set_auto_index(self.handle, int(after_n_scans)))

    @property
    def last_scan_stats(self) -> B.ScanStats:
        st = B.ScanStats()
        check(lib.pcq_ctx_last_scan_stats(self.handle, C.byref(st)))
        return st


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def ext_of(path: Union[str, os.PathLike]) -> Optional[str]:
    s = Path(path).suffix
    return s[1:] if s else None


class DeviceFile:
    """A LAS / LAST file (or a point range of one) resident in HBM (pcq_file)."""

    def __init__(self, ctx: Context, handle: C.c_void_p, keepalive=None):
        self.ctx = ctx
        self.handle = handle
        self._keepalive = keepalive

    @classmethod
    def stage(cls, ctx: Context, file_bytes, ext: str, first_point: int = 0, n_points: Optional[int] = None) -> "DeviceFile":
        addr, n = B.buffer_address(file_bytes)
        h = C.c_void_p()
        npts = 0xFFFFFFFFFFFFFFFF if n_points is None else int(n_points)
        check(lib.pcq_file_stage_host(ctx.handle, C.c_void_p(addr), n, ext.encode(), int(first_point), npts, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def wrap(cls, ctx: Context, desc: B.FileDesc, dev_ptr: int, first_point_index: int = 0, keepalive=None) -> "DeviceFile":
        h = C.c_void_p()
        check(lib.pcq_file_wrap_device(ctx.handle, C.byref(desc), C.c_void_p(int(dev_ptr)), int(first_point_index), C.byref(h)))
        return cls(ctx, h, keepalive)

    def set_scan_base(self, base: int):
        check(lib.pcq_file_set_scan_base(self.handle, int(base)))

    @property
    def desc(self) -> B.FileDesc:
        d = B.FileDesc()
        check(lib.pcq_file_desc_get(self.handle, C.byref(d)))
        return d

    def build_index(self):
        """Chunk headers (improvements.md:3-10): later searches launch over the chunks that can hold a match."""
        check(lib.pcq_file_build_index(self.handle))

    def drop_index(self):
        lib.pcq_file_drop_index(self.handle)

    @property
    def index(self) -> np.ndarray:
        """Copy of the chunk headers (CHUNK_HEADER_DTYPE); empty when the file has no index."""
        p, n = C.c_void_p(), C.c_uint64()
        check(lib.pcq_file_index(self.handle, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=B.CHUNK_HEADER_DTYPE)
        raw = C.string_at(p.value, n.value * B.CHUNK_HEADER_DTYPE.itemsize)
        return np.frombuffer(raw, dtype=B.CHUNK_HEADER_DTYPE).copy()

    def release(self):
        if self.handle:
            lib.pcq_file_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------
# collectors (collect_points.rs)
# ---------------------------------------------------------------------------------------------------
class ResultCollector:
    """trait ResultCollector (collect_points.rs:7-12), filled in bulk on the device."""

    kind = -1

    def __init__(self, ctx: Optional[Context] = None, gmin=None, gmax=None, cell_size: float = 0.0):
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        a = B.d3(gmin) if gmin is not None else None
        b = B.d3(gmax) if gmax is not None else None
        check(lib.pcq_collector_create(self.ctx.handle, self.kind, a, b, float(cell_size), C.byref(h)))
        self.handle = h

    def collect_one(self, point):  # pragma: no cover - documented deviation
        raise NotImplementedError(
            "collect_one is a per-point callback; across the device boundary collectors are filled in bulk by search_file"
        )

    def point_count(self) -> int:
        n = C.c_uint64()
        check(lib.pcq_collector_point_count(self.handle, C.byref(n)))
        return int(n.value)

    def points(self) -> Optional[np.ndarray]:
        p = C.c_void_p()
        n = C.c_uint64()
        check(lib.pcq_collector_points(self.handle, C.byref(p), C.byref(n)))
        if self.kind == B.COLLECT_COUNT:
            return None
        if n.value == 0:
            return np.zeros(0, dtype=B.POINT_DTYPE)
        raw = (C.c_uint8 * (31 * n.value)).from_address(p.value)
        return np.frombuffer(raw, dtype=B.POINT_DTYPE).copy()

    def points_ref(self) -> Optional[np.ndarray]:
        return self.points() if self.kind == B.COLLECT_BUFFER else None

    def las_records(self):
        """FileDumper's output for this collector (dump_points.rs:63-116), reduced and quantised on the device:
        -> (offset[3], max[3], scale, records uint8[n, 26]) or None when there is nothing to write."""
        mn, mx = (C.c_double * 3)(), (C.c_double * 3)()
        scale, p, n = C.c_double(), C.c_void_p(), C.c_uint64()
        check(lib.pcq_collector_las_records(self.handle, mn, mx, C.byref(scale), C.byref(p), C.byref(n)))
        if n.value == 0:
            return None
        raw = (C.c_uint8 * (26 * n.value)).from_address(p.value)
        return np.array(mn[:]), np.array(mx[:]), float(scale.value), np.frombuffer(raw, dtype=np.uint8).reshape(-1, 26).copy()

    def reset(self):
        check(lib.pcq_collector_reset(self.handle))

    def close(self):
        if self.handle:
            lib.pcq_collector_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reset_collectors(collectors: Sequence[ResultCollector]) -> None:
    """pcq_collectors_reset: clear a list of collectors (one kernel launch per context for count / buffer collectors)."""
    n = len(collectors)
    if n:
        check(lib.pcq_collectors_reset((C.c_void_p * n)(*[c.handle for c in collectors]), n))


class CountCollector(ResultCollector):
    kind = B.COLLECT_COUNT


class BufferCollector(ResultCollector):
    kind = B.COLLECT_BUFFER


class GridSampledCollector(ResultCollector):
    """GridSampledCollector::new(bounds, cell_size) (collect_points.rs:104-108)."""

    kind = B.COLLECT_GRID

    def __init__(self, bounds_min, bounds_max, cell_size: float, ctx: Optional[Context] = None):
        super().__init__(ctx, bounds_min, bounds_max, cell_size)

    def export_candidates(self, n_parts: int):
        """Local winners partitioned by owner rank -> (device pointer, counts per part)."""
        p = C.c_void_p()
        counts = (C.c_uint64 * n_parts)()
        check(lib.pcq_grid_export_candidates(self.handle, n_parts, C.byref(p), counts))
        return p.value, [int(c) for c in counts]

    def import_candidates(self, dev_ptr: int, n: int):
        check(lib.pcq_grid_import_candidates(self.handle, C.c_void_p(int(dev_ptr or 0)), int(n)))


# ---------------------------------------------------------------------------------------------------
# searchers (searcher.rs)
# ---------------------------------------------------------------------------------------------------
class SearchImplementation(enum.Enum):
    Regular = 0
    Optimized = 1


FileLike = Union[str, os.PathLike, DeviceFile, tuple]


def _as_device_files(ctx: Context, files: Sequence[FileLike]):
    """-> (device files, the ones we staged ourselves and must release)"""
    out, owned = [], []
    for f in files:
        if isinstance(f, DeviceFile):
            out.append(f)
            continue
        if isinstance(f, tuple):  # (bytes, ext): an in-memory file image
            data, ext = f
        else:
            ext = ext_of(f)
            if ext is None:
                raise PcqError(B.PCQ_ERR_FORMAT, f"Invalid extension on file {f}")
            if ext not in ("las", "last"):
                # laz / lazer stay on the reference (arithmetic / LZ4 decoding is out of scope)
                raise PcqError(B.PCQ_ERR_FORMAT, f"Unsupported file extension in file {f}")
            data = np.fromfile(f, dtype=np.uint8)
        df = DeviceFile.stage(ctx, data, ext)
        out.append(df)
        owned.append(df)
    return out, owned


class Searcher:
    """trait Searcher (searcher.rs:24-31)."""

    def _query(self) -> B.Query:
        raise NotImplementedError

    def search_file(self, path: FileLike, search_impl: SearchImplementation, collector: ResultCollector) -> None:
        self.search_files([path], search_impl, [collector])

    def search_files(self, files: Sequence[FileLike], search_impl: SearchImplementation,
                     collectors: Sequence[ResultCollector]) -> None:
        """One launch for a batch: 1 collector = run_search_sequential, len(files) = run_search_parallel."""
        if search_impl is not SearchImplementation.Optimized:
            raise NotImplementedError("SearchImplementation::Regular (pasture readers + f64 contains) stays on the reference")
        ctx = collectors[0].ctx
        dfs, owned = _as_device_files(ctx, files)
        try:
            fh = (C.c_void_p * len(dfs))(*[d.handle for d in dfs])
            ch = (C.c_void_p * len(collectors))(*[c.handle for c in collectors])
            q = self._query()
            check(lib.pcq_search_files(ctx.handle, fh, len(dfs), C.byref(q), ch, len(collectors)))
            if owned:
                ctx.synchronize()
        finally:
            for d in owned:
                d.release()

    def search_host_files(self, images: Sequence[tuple], collectors: Sequence[ResultCollector]) -> None:
        """Host-staged scan of (buffer, ext) file images (pcq_search_host_files)."""
        ctx = collectors[0].ctx
        n = len(images)
        addrs = [B.buffer_address(b) for b, _ in images]
        ptrs = (C.c_void_p * n)(*[a for a, _ in addrs])
        sizes = (C.c_size_t * n)(*[s for _, s in addrs])
        exts = (C.c_char_p * n)(*[e.encode() for _, e in images])
        ch = (C.c_void_p * len(collectors))(*[c.handle for c in collectors])
        q = self._query()
        check(lib.pcq_search_host_files(ctx.handle, ptrs, sizes, exts, n, C.byref(q), ch, len(collectors)))


def index_filter(headers: np.ndarray, desc: B.FileDesc, searcher: "Searcher", join_gap: int = 1):
    """pcq_index_filter (host only): -> (runs as an (n, 2) array of [first_chunk, end_chunk), chunks that may match)"""
    headers = np.ascontiguousarray(headers, dtype=B.CHUNK_HEADER_DTYPE)
    cap = max(1, headers.shape[0])
    runs = np.zeros((cap, 2), dtype=np.uint64)
    n_runs, n_may = C.c_uint64(), C.c_uint64()
    q = searcher._query()
    check(lib.pcq_index_filter(C.c_void_p(headers.ctypes.data), headers.shape[0], C.byref(desc), C.byref(q), int(join_gap),
                               C.c_void_p(runs.ctypes.data), cap, C.byref(n_runs), C.byref(n_may)))
    return runs[: n_runs.value].astype(np.int64), int(n_may.value)


class HostIndex:
    """Chunk headers of a list of host file images (pcq_host_index): built as a by-product of the first host-staged
    pass, used by later passes to copy only the chunks that can hold a match."""

    def __init__(self, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        check(lib.pcq_host_index_create(self.ctx.handle, C.byref(h)))
        self.handle = h

    def info(self, file: int) -> tuple:
        """-> (chunks, has AABB part, has class part) of file `file` of the list"""
        n, hb, hc = C.c_uint64(), C.c_int(), C.c_int()
        check(lib.pcq_host_index_info(self.handle, int(file), C.byref(n), C.byref(hb), C.byref(hc)))
        return int(n.value), bool(hb.value), bool(hc.value)

    def close(self):
        if self.handle:
            lib.pcq_host_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def search_host_files_multi(images: Sequence[tuple], searchers: Sequence["Searcher"],
                            collectors_per_query: Sequence[Sequence[ResultCollector]],
                            index: Optional[HostIndex] = None) -> None:
    """One host-staged pass for several queries (pcq_search_host_files_multi): the file images cross PCIe once.
    With `index` (pcq_search_host_files_indexed) only the chunks that can hold a match cross it."""
    ctx = collectors_per_query[0][0].ctx
    n, nq, ncol = len(images), len(searchers), len(collectors_per_query[0])
    assert all(len(c) == ncol for c in collectors_per_query) and len(collectors_per_query) == nq
    addrs = [B.buffer_address(b) for b, _ in images]
    ptrs = (C.c_void_p * n)(*[a for a, _ in addrs])
    sizes = (C.c_size_t * n)(*[s for _, s in addrs])
    exts = (C.c_char_p * n)(*[e.encode() for _, e in images])
    qs = (B.Query * nq)(*[s._query() for s in searchers])
    ch = (C.c_void_p * (nq * ncol))(*[c.handle for cs in collectors_per_query for c in cs])
    if index is not None:
        check(lib.pcq_search_host_files_indexed(ctx.handle, ptrs, sizes, exts, n, qs, nq, ch, ncol, index.handle))
    else:
        check(lib.pcq_search_host_files_multi(ctx.handle, ptrs, sizes, exts, n, qs, nq, ch, ncol))


class BoundsSearcher(Searcher):
    """BoundsSearcher::new(bounds) (searcher.rs:33-41); bounds = (min xyz, max xyz)."""

    def __init__(self, bounds_min, bounds_max):
        self.bounds_min = tuple(float(v) for v in bounds_min)
        self.bounds_max = tuple(float(v) for v in bounds_max)

    def _query(self) -> B.Query:
        q = B.Query()
        q.kind = B.QUERY_BOUNDS
        q.qmin = B.d3(self.bounds_min)
        q.qmax = B.d3(self.bounds_max)
        return q


class ClassSearcher(Searcher):
    """ClassSearcher::new(class) (searcher.rs:94-102)."""

    def __init__(self, cls: int):
        if not 0 <= int(cls) <= 255:
            raise ValueError("class is a u8")
        self.cls = int(cls)

    def _query(self) -> B.Query:
        q = B.Query()
        q.kind = B.QUERY_CLASS
        q.cls = self.cls
        return q


def _expect_ext(path, want: str):
    if isinstance(path, (str, os.PathLike)) and ext_of(path) != want:
        raise PcqError(B.PCQ_ERR_FORMAT, f"{path} is not a .{want} file")


def search_las_file_by_bounds_optimized(path: FileLike, bounds, result_collector: ResultCollector) -> None:
    """las.rs:52-148."""
    _expect_ext(path, "las")
    BoundsSearcher(*bounds).search_file(path, SearchImplementation.Optimized, result_collector)


def search_las_file_by_classification_optimized(path: FileLike, cls: int, result_collector: ResultCollector) -> None:
    """las.rs:192-261."""
    _expect_ext(path, "las")
    ClassSearcher(cls).search_file(path, SearchImplementation.Optimized, result_collector)


def search_last_file_by_bounds_optimized(path: FileLike, bounds, result_collector: ResultCollector) -> None:
    """last.rs:46-166."""
    _expect_ext(path, "last")
    BoundsSearcher(*bounds).search_file(path, SearchImplementation.Optimized, result_collector)


def search_last_file_by_classification_optimized(path: FileLike, cls: int, result_collector: ResultCollector) -> None:
    """last.rs:213-293."""
    _expect_ext(path, "last")
    ClassSearcher(cls).search_file(path, SearchImplementation.Optimized, result_collector)


CollectorFactoryFn = Callable[[], ResultCollector]


def run_search_sequential(files: Sequence[FileLike], searcher: Searcher, search_impl: SearchImplementation,
                          collector_factory_fn: CollectorFactoryFn) -> ResultCollector:
    """main.rs:122-144: ONE collector across all files, in `files` order."""
    collector = collector_factory_fn()
    searcher.search_files(files, search_impl, [collector])
    return collector


def run_search_parallel(files: Sequence[FileLike], searcher: Searcher, search_impl: SearchImplementation,
                        collector_factory_fn: CollectorFactoryFn) -> list:
    """main.rs:146-183: one collector PER FILE (rayon par_iter); results stay in `files` order."""
    collectors = [collector_factory_fn() for _ in files]
    if files:
        searcher.search_files(files, search_impl, collectors)
    return collectors
