"""ctypes binding of libpcq.so (include/pcq.h, include/pcq_synth.h).

Harness-side only: tests and bench.py drive the C ABI through this module exactly as a Rust
``extern "C"`` block would.  Nothing here computes on points; if the shared
library is missing the import fails loudly — there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libpcq.so"

PCQ_OK = 0
PCQ_ERR_ARG = -1
PCQ_ERR_FORMAT = -2
PCQ_ERR_PANIC = -3
PCQ_ERR_CUDA = -4
PCQ_ERR_NOMEM = -5
PCQ_ERR_GRID = -6
PCQ_ERR_ALIASED = -7
PCQ_ERR_IO = -8

LAYOUT_LAS, LAYOUT_LAST = 0, 1
QUERY_BOUNDS, QUERY_CLASS = 0, 1
COLLECT_COUNT, COLLECT_BUFFER, COLLECT_GRID = 0, 1, 2
SHAPE_UNIFORM, SHAPE_TERRAIN, SHAPE_INDOOR, SHAPE_RELIEF = 0, 1, 2, 3

# readers::Point, readers/src/lib.rs:10-19 (repr(C, packed), 31 bytes)
POINT_DTYPE = np.dtype([("pos", "<f8", (3,)), ("rgb", "<u2", (3,)), ("cls", "u1")])
assert POINT_DTYPE.itemsize == 31

CANDIDATE_DTYPE = np.dtype(
    [("key", "<u8"), ("dist_bits", "<u8"), ("scan_idx", "<u8"), ("point", POINT_DTYPE), ("pad", "u1", (9,))]
)
assert CANDIDATE_DTYPE.itemsize == 64

_D3 = C.c_double * 3


class FileDesc(C.Structure):
    _fields_ = [
        ("layout", C.c_uint8),
        ("format", C.c_uint8),
        ("record_len", C.c_uint16),
        ("point_data_off", C.c_uint32),
        ("n_points", C.c_uint64),
        ("scale", _D3),
        ("offset", _D3),
        ("hdr_min", _D3),
        ("hdr_max", _D3),
    ]


class Query(C.Structure):
    _fields_ = [
        ("kind", C.c_uint8),
        ("cls", C.c_uint8),
        ("pad_", C.c_uint8 * 6),
        ("qmin", _D3),
        ("qmax", _D3),
    ]


class SynthSpec(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("n_points", C.c_uint64),
        ("layout", C.c_uint8),
        ("format", C.c_uint8),
        ("shape", C.c_uint8),
        ("n_classes", C.c_uint8),
        ("record_len", C.c_uint16),
        ("flag_per_64k", C.c_uint16),
        ("lo", C.c_int32 * 3),
        ("hi", C.c_int32 * 3),
        ("scale", _D3),
        ("offset", _D3),
        ("class_val", C.c_uint8 * 8),
        ("class_cum", C.c_uint16 * 8),
    ]


INDEX_CHUNK_POINTS = 8192
SHARD_RANGES, SHARD_FILES = 0, 1
GROUP_ID_BYTES = 128

SHARD_DTYPE = np.dtype([("file", "<u4"), ("rank", "<u4"), ("first_point", "<u8"), ("n_points", "<u8")])
assert SHARD_DTYPE.itemsize == 24

# pcq_chunk_header: integer AABB of the raw x/y/z fields + set of class bytes of one chunk (64 bytes)
CHUNK_HEADER_DTYPE = np.dtype(
    [("lo", "<i4", (3,)), ("hi", "<i4", (3,)), ("cls_bits", "<u4", (8,)), ("n_points", "<u4"), ("pad", "<u4")]
)
assert CHUNK_HEADER_DTYPE.itemsize == 64


class ScanStats(C.Structure):
    _fields_ = [
        ("points_total", C.c_uint64),
        ("points_scanned", C.c_uint64),
        ("chunks_total", C.c_uint64),
        ("chunks_skipped", C.c_uint64),
        ("segments", C.c_uint32),
        ("pad_", C.c_uint32),
    ]


class GroupStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("scan_ms", "rescan_ms", "export_ms", "exchange_ms", "import_ms", "finalize_ms")] + \
               [(n, C.c_uint64) for n in ("cells_sent", "log_entries_sent", "bytes_sent", "affected_keys")]


class PcqError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"pcq error {code}: {message}")
        self.code = code
        self.message = message


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL)
    vp, u64, u32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t
    P = C.POINTER
    sigs = {
        "pcq_last_error": (C.c_char_p, []),
        "pcq_version": (C.c_char_p, []),
        "pcq_parse_header": (C.c_int, [vp, sz, C.c_int, C.c_int, P(FileDesc)]),
        "pcq_local_bounds": (C.c_int, [P(FileDesc), _D3, _D3, C.c_int64 * 3, C.c_int64 * 3]),
        "pcq_file_intersects": (C.c_int, [P(FileDesc), _D3, _D3, P(C.c_int)]),
        "pcq_grid_params": (C.c_int, [_D3, _D3, C.c_double, u64 * 3, u64 * 3]),
        "pcq_ctx_create": (C.c_int, [C.c_int, P(vp)]),
        "pcq_ctx_destroy": (None, [vp]),
        "pcq_ctx_set_stream": (C.c_int, [vp, vp]),
        "pcq_ctx_synchronize": (C.c_int, [vp]),
        "pcq_ctx_set_scan_variant": (C.c_int, [vp, C.c_int]),
        "pcq_ctx_launch_count": (u64, [vp]),
        "pcq_file_stage_host": (C.c_int, [vp, vp, sz, C.c_char_p, u64, u64, P(vp)]),
        "pcq_file_wrap_device": (C.c_int, [vp, P(FileDesc), vp, u64, P(vp)]),
        "pcq_file_set_scan_base": (C.c_int, [vp, u64]),
        "pcq_file_desc_get": (C.c_int, [vp, P(FileDesc)]),
        "pcq_file_release": (None, [vp]),
        "pcq_collector_create": (C.c_int, [vp, C.c_int, vp, vp, C.c_double, P(vp)]),
        "pcq_collector_destroy": (None, [vp]),
        "pcq_collector_reset": (C.c_int, [vp]),
        "pcq_collectors_reset": (C.c_int, [P(vp), u32]),
        "pcq_grid_cells_under_box": (C.c_int, [_D3, _D3, C.c_double, _D3, _D3, u64 * 3, u64 * 3]),
        "pcq_collector_point_count": (C.c_int, [vp, P(u64)]),
        "pcq_collector_points": (C.c_int, [vp, P(vp), P(u64)]),
        "pcq_collector_points_device": (C.c_int, [vp, P(vp), P(u64)]),
        "pcq_collector_las_records": (C.c_int, [vp, P(C.c_double), P(C.c_double), P(C.c_double), P(vp), P(u64)]),
        "pcq_search_files": (C.c_int, [vp, P(vp), u32, P(Query), P(vp), u32]),
        "pcq_search_host_files": (C.c_int, [vp, P(vp), P(sz), P(C.c_char_p), u32, P(Query), P(vp), u32]),
        "pcq_search_host_files_multi": (C.c_int, [vp, P(vp), P(sz), P(C.c_char_p), u32, P(Query), u32, P(vp), u32]),
        "pcq_file_build_index": (C.c_int, [vp]),
        "pcq_file_drop_index": (None, [vp]),
        "pcq_file_index": (C.c_int, [vp, P(vp), P(u64)]),
        "pcq_index_filter": (C.c_int, [vp, u64, P(FileDesc), P(Query), u64, vp, u64, P(u64), P(u64)]),
        "pcq_ctx_set_auto_index": (C.c_int, [vp, u32]),
        "pcq_ctx_last_scan_stats": (C.c_int, [vp, P(ScanStats)]),
        "pcq_host_index_create": (C.c_int, [vp, P(vp)]),
        "pcq_host_index_destroy": (None, [vp]),
        "pcq_host_index_info": (C.c_int, [vp, u32, P(u64), P(C.c_int), P(C.c_int)]),
        "pcq_search_host_files_indexed": (C.c_int, [vp, P(vp), P(sz), P(C.c_char_p), u32, P(Query), u32, P(vp), u32, vp]),
        "pcq_host_alloc": (C.c_int, [sz, P(vp)]),
        "pcq_host_free": (None, [vp]),
        "pcq_grid_export_candidates": (C.c_int, [vp, u32, P(vp), P(u64)]),
        "pcq_grid_import_candidates": (C.c_int, [vp, vp, u64]),
        "pcq_host_register": (C.c_int, [vp, sz]),
        "pcq_host_unregister": (C.c_int, [vp]),
        "pcq_ctx_bind_host_thread": (C.c_int, [vp, P(C.c_int)]),
        "pcq_group_create": (C.c_int, [P(C.c_int), u32, P(vp)]),
        "pcq_group_unique_id": (C.c_int, [vp]),
        "pcq_group_create_rank": (C.c_int, [C.c_int, u32, u32, vp, P(vp)]),
        "pcq_group_destroy": (None, [vp]),
        "pcq_group_world": (u32, [vp]),
        "pcq_group_local_count": (u32, [vp]),
        "pcq_group_local_rank": (u32, [vp, u32]),
        "pcq_group_ctx": (vp, [vp, u32]),
        "pcq_group_launch_count": (u64, [vp]),
        "pcq_group_synchronize": (C.c_int, [vp]),
        "pcq_group_last_stats": (C.c_int, [vp, P(GroupStats)]),
        "pcq_shard_plan": (C.c_int, [P(u64), u32, u32, C.c_int, vp, u64, P(u64)]),
        "pcq_group_stage_host_files": (C.c_int, [vp, P(vp), P(sz), P(C.c_char_p), u32, C.c_int, P(vp)]),
        "pcq_group_wrap_files": (C.c_int, [vp, P(u64), u32, P(vp), P(u32), P(u32), u32, P(vp)]),
        "pcq_dataset_release": (None, [vp]),
        "pcq_group_search": (C.c_int, [vp, vp, P(Query), u32, C.c_int, vp, vp, C.c_double, C.c_int, P(vp)]),
        "pcq_group_search_host_files": (C.c_int, [vp, P(vp), P(sz), P(C.c_char_p), u32, P(Query), u32, C.c_int, vp, vp,
                                                  C.c_double, C.c_int, C.c_int, P(vp)]),
        "pcq_result_counts": (C.c_int, [vp, P(P(u64)), P(u32)]),
        "pcq_result_points": (C.c_int, [vp, u32, P(vp), P(u64)]),
        "pcq_result_release": (None, [vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here == an exported symbol is missing
        fn.restype = res
        fn.argtypes = args
    lib._pcq_symbols = tuple(sigs)
    return lib


lib = _load()

SYNTH_LIB_PATH = _HERE / "libpcq_synth.so"


def _load_synth() -> C.CDLL:
    """include/pcq_synth.h — the synthetic data generator is a library of its own (test / benchmark tooling)."""
    if not SYNTH_LIB_PATH.exists():
        raise ImportError(f"{SYNTH_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    sl = C.CDLL(str(SYNTH_LIB_PATH), mode=os.RTLD_LOCAL)
    vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
    P = C.POINTER
    sigs = {
        "pcq_synth_last_error": (C.c_char_p, []),
        "pcq_synth_file_size": (sz, [P(SynthSpec)]),
        "pcq_synth_host": (C.c_int, [P(SynthSpec), vp, sz]),
        "pcq_synth_host_points": (C.c_int, [P(SynthSpec), u64, u64, vp, C.c_int32 * 6]),
        "pcq_synth_device_points": (C.c_int, [C.c_int, P(SynthSpec), u64, u64, vp, C.c_int32 * 6]),
        "pcq_synth_device": (C.c_int, [C.c_int, P(SynthSpec), vp, C.c_int32 * 6]),
        "pcq_synth_header": (C.c_int, [P(SynthSpec), C.c_int32 * 6, vp]),
        "pcq_synth_desc": (C.c_int, [P(SynthSpec), C.c_int32 * 6, P(FileDesc)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(sl, name)
        fn.restype = res
        fn.argtypes = args
    sl._pcq_symbols = tuple(sigs)
    return sl


synth_lib = _load_synth()


def check_synth(rc: int) -> None:
    if rc != PCQ_OK:
        raise PcqError(rc, synth_lib.pcq_synth_last_error().decode("utf-8", "replace"))


def check(rc: int) -> None:
    if rc != PCQ_OK:
        raise PcqError(rc, lib.pcq_last_error().decode("utf-8", "replace"))


def d3(v) -> "_D3":
    return _D3(float(v[0]), float(v[1]), float(v[2]))


def buffer_address(buf) -> tuple[int, int]:
    """(address, nbytes) of a bytes / bytearray / numpy / memoryview buffer without copying; an (address, nbytes) pair
    (memory the caller manages, e.g. a registered mapping) is passed through."""
    if isinstance(buf, tuple):
        return int(buf[0]), int(buf[1])
    if isinstance(buf, np.ndarray):
        assert buf.flags["C_CONTIGUOUS"]
        return buf.ctypes.data, buf.nbytes
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value, len(buf)
    mv = memoryview(buf)
    arr = np.frombuffer(mv, dtype=np.uint8)
    return arr.ctypes.data, arr.nbytes
