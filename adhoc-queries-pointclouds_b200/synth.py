"""Synthetic datasets of the navvis / doc / ca13 shapes (include/pcq_synth.h).

The reference's datasets are private; sizes and query boxes come from
query/src/bin/run_query_experiments.rs:109-144, 320-343 and run_postgis_queries.rs:22-24.
Generation is integer-only and counter-based: `host_file` (small, for the oracle) and
`device_points` (billions of points, for the benchmark) produce identical bytes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import binding as B
from .binding import check_synth as check, synth_lib as lib

FORMAT_LEN = {0: 20, 1: 28, 2: 26, 3: 34}

# class mixes: (class value, probability)
DOC_CLASSES = ((1, 0.25), (2, 0.45), (5, 0.15), (6, 0.10), (7, 0.02), (9, 0.03))
CA13_CLASSES = ((2, 0.55), (6, 0.04), (1, 0.25), (5, 0.14), (9, 0.02))

# query boxes of run_query_experiments.rs:109-144
NAVVIS_S = ((0.0, 0.0, 0.0), (2.0, 2.0, 2.0))
NAVVIS_L = ((0.0, 0.0, 0.0), (20.0, 20.0, 5.0))
NAVVIS_XL = ((-23.108, -21.261, -10.029), (28.588, 27.123, 5.959))
DOC_S = ((390000.0, 130000.0, 0.0), (390500.0, 140000.0, 200.0))
DOC_L = ((390000.0, 130000.0, 0.0), (400000.0, 140000.0, 200.0))
DOC_XL = ((389400.0, 124200.0, -94.88), (406200.0, 148200.0, 760.03))
CA13_S = ((665000.0, 3910000.0, 0.0), (705000.0, 3950000.0, 480.0))
CA13_L = ((665000.0, 3910000.0, 0.0), (710000.0, 3950000.0, 480.0))
CA13_XL = ((643431.76, 3883547.565, -46194.145), (736910.93, 3977026.735, 47285.025))
NAVVIS_DENSITY, DOC_DENSITY, CA13_DENSITY = 0.1, 25.0, 100.0


def make_spec(seed: int, n_points: int, layout: int, fmt: int, shape: int, lo: Sequence[int], hi: Sequence[int],
              scale: Sequence[float], offset: Sequence[float], classes=DOC_CLASSES, flag_per_64k: int = 655,
              record_len: Optional[int] = None) -> B.SynthSpec:
    sp = B.SynthSpec()
    sp.seed = seed & 0xFFFFFFFFFFFFFFFF
    sp.n_points = int(n_points)
    sp.layout = layout
    sp.format = fmt
    sp.shape = shape
    sp.n_classes = len(classes)
    sp.record_len = record_len if record_len is not None else FORMAT_LEN[fmt]
    sp.flag_per_64k = flag_per_64k
    for a in range(3):
        sp.lo[a] = int(lo[a])
        sp.hi[a] = int(hi[a])
        sp.scale[a] = float(scale[a])
        sp.offset[a] = float(offset[a])
    cum = 0.0
    for k, (val, p) in enumerate(classes):
        cum += p
        sp.class_val[k] = int(val)
        sp.class_cum[k] = 65535 if k == len(classes) - 1 else min(65535, int(round(cum * 65536)) - 1)
    return sp


def file_size(spec: B.SynthSpec) -> int:
    return int(lib.pcq_synth_file_size(C.byref(spec)))


def host_file(spec: B.SynthSpec) -> np.ndarray:
    """Whole file image (227-byte LAS 1.2 header + point data) as a uint8 array."""
    out = np.empty(file_size(spec), dtype=np.uint8)
    check(lib.pcq_synth_host(C.byref(spec), C.c_void_p(out.ctypes.data), out.nbytes))
    return out


def device_points(ctx, spec: B.SynthSpec, dev_ptr: int):
    """Fill device memory with the point data of `spec`; -> (raw minmax[6], FileDesc)."""
    mm = (C.c_int32 * 6)()
    ctx.synchronize()
    check(lib.pcq_synth_device(int(ctx.device), C.byref(spec), C.c_void_p(int(dev_ptr)), mm))
    desc = B.FileDesc()
    check(lib.pcq_synth_desc(C.byref(spec), mm, C.byref(desc)))
    return [int(v) for v in mm], desc


def device_point_range(device: int, spec: B.SynthSpec, first_point: int, n_points: int, dev_ptr: int):
    """Point data of the range [first_point, first_point + n_points) of the file into device memory (a block of its
    own); -> raw minmax[6] of the range (INT_MAX / INT_MIN when it is empty)."""
    mm = (C.c_int32 * 6)()
    check(lib.pcq_synth_device_points(int(device), C.byref(spec), int(first_point), int(n_points), C.c_void_p(int(dev_ptr)), mm))
    return [int(v) for v in mm]


def host_point_range(spec: B.SynthSpec, first_point: int, n_points: int, out: np.ndarray):
    """Same bytes on the host, written into `out` (uint8, C-contiguous); -> raw minmax[6] of the range.  Releases the
    GIL: callers split files over threads."""
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
    mm = (C.c_int32 * 6)()
    check(lib.pcq_synth_host_points(C.byref(spec), int(first_point), int(n_points), C.c_void_p(out.ctypes.data), mm))
    return [int(v) for v in mm]


def desc_of(spec: B.SynthSpec, minmax: Sequence[int]) -> B.FileDesc:
    mm = (C.c_int32 * 6)(*[int(v) for v in minmax])
    desc = B.FileDesc()
    check(lib.pcq_synth_desc(C.byref(spec), mm, C.byref(desc)))
    return desc


def header_bytes(spec: B.SynthSpec, minmax: Sequence[int]) -> np.ndarray:
    mm = (C.c_int32 * 6)(*[int(v) for v in minmax])
    out = np.empty(227, dtype=np.uint8)
    check(lib.pcq_synth_header(C.byref(spec), mm, C.c_void_p(out.ctypes.data)))
    return out


def ext_of_layout(layout: int) -> str:
    return "las" if layout == B.LAYOUT_LAS else "last"


# ---------------------------------------------------------------------------------------------------
# dataset shapes
# ---------------------------------------------------------------------------------------------------
def doc_specs(n_files: int = 64, pts_per_file: int = 31_250_000, layout: int = B.LAYOUT_LAS, fmt: int = 1,
              seed: int = 0x5EED0000 + 2000, fma_sensitive: bool = False) -> list:
    """doc-shape tiles (C2): a g x g grid over the doc-XL footprint, scale 0.01, offset = tile corner."""
    g = max(1, int(round(n_files ** 0.5)))
    assert g * g == n_files, "doc dataset is a square tile grid"
    (x0, y0, _), (x1, y1, _) = DOC_XL
    w, h = (x1 - x0) / g, (y1 - y0) / g
    specs = []
    for j in range(g):
        for i in range(g):
            k = j * g + i
            scale = 0.001 if fma_sensitive else 0.01
            ox = x0 + i * w + (0.37 if fma_sensitive else 0.0)
            oy = y0 + j * h + (0.565 if fma_sensitive else 0.0)
            oz = -23.108 if fma_sensitive else 0.0
            hi_x, hi_y = int(round(w / scale)) - 1, int(round(h / scale)) - 1
            lo_z, hi_z = int(round((-94.88 - oz) / scale)), int(round((194.88 - oz) / scale))
            specs.append(make_spec(seed + k, pts_per_file, layout, fmt, B.SHAPE_TERRAIN, (0, 0, lo_z), (hi_x, hi_y, hi_z),
                                   (scale,) * 3, (ox, oy, oz), DOC_CLASSES))
    return specs


def ca13_specs(n_files: int = 64, pts_per_file: int = 40_750_000, layout: int = B.LAYOUT_LAST, fmt: int = 1,
               seed: int = 0x5EED0000 + 3000) -> list:
    """ca13-shape tiles (C3): LAST files over the ca13-XL footprint, terrain z in [0, 480]."""
    g = max(1, int(round(n_files ** 0.5)))
    assert g * g == n_files
    (x0, y0, _), (x1, y1, _) = CA13_XL
    w, h = (x1 - x0) / g, (y1 - y0) / g
    specs = []
    for j in range(g):
        for i in range(g):
            k = j * g + i
            ox, oy = float(int(x0 + i * w)), float(int(y0 + j * h))
            hi_x, hi_y = int(w / 0.01), int(h / 0.01)
            specs.append(make_spec(seed + k, pts_per_file, layout, fmt, B.SHAPE_RELIEF, (0, 0, 0), (hi_x, hi_y, 48000),
                                   (0.01,) * 3, (ox, oy, 0.0), CA13_CLASSES))
    return specs


def navvis_spec(n_points: int = 56_200_000, layout: int = B.LAYOUT_LAS, fmt: int = 3, seed: int = 0x5EED0000 + 4000,
                fma_sensitive: bool = False) -> B.SynthSpec:
    """navvis-shape single dense indoor file (C4): scale 0.001 inside the navvis-XL box."""
    (x0, y0, z0), (x1, y1, z1) = NAVVIS_XL
    off = (x0, y0, z0) if fma_sensitive else (0.0, 0.0, 0.0)
    lo = [int(round((a - o) * 1000)) for a, o in zip((x0, y0, z0), off)]
    hi = [int(round((a - o) * 1000)) for a, o in zip((x1, y1, z1), off)]
    return make_spec(seed, n_points, layout, fmt, B.SHAPE_INDOOR, lo, hi, (0.001,) * 3, off, ((1, 0.6), (2, 0.2), (6, 0.2)))


def uniform_spec(n_points: int, layout: int, fmt: int, seed: int = 0x5EED0000 + 5000, extent: int = 1_000_000,
                 scale=(0.01, 0.01, 0.01), offset=(0.0, 0.0, 0.0), record_len: Optional[int] = None) -> B.SynthSpec:
    """C5 sweep: uniform points in a cube of `extent` raw units."""
    return make_spec(seed, n_points, layout, fmt, B.SHAPE_UNIFORM, (0, 0, 0), (extent - 1,) * 3, scale, offset,
                     DOC_CLASSES, record_len=record_len)


def strips_device(ctx, n_points: int, n_strips: int = 64, layout: int = B.LAYOUT_LAS, fmt: int = 1,
                  rare_class_every: int = 4, seed: int = 0x5EED0000 + 6000, extent: int = 1_000_000):
    """One device-resident file in acquisition order (the chunk-index workload): `n_strips` flight strips stored one
    after the other, strip k covering its own band of x (10 % overlap with its neighbours) over the whole of y;
    class 6 occurs only in every `rare_class_every`-th strip.  -> (torch uint8 buffer, FileDesc, 227-byte header).
    Harness-side plumbing (torch holds the memory and, for LAST, transposes the records into columns)."""
    import torch

    R = FORMAT_LEN[fmt]
    per = n_points // n_strips
    n_points = per * n_strips
    dev = f"cuda:{ctx.device}"
    buf = torch.empty(n_points * R + 256, dtype=torch.uint8, device=dev)
    w = extent // n_strips
    lo_all, hi_all = [2**31 - 1] * 3, [-(2**31)] * 3
    for k in range(n_strips):
        classes = DOC_CLASSES if k % rare_class_every == 0 else tuple((v, p) for v, p in DOC_CLASSES if v != 6)
        tot = sum(p for _, p in classes)
        classes = tuple((v, p / tot) for v, p in classes)
        x0, x1 = max(0, k * w - w // 20), min(extent - 1, (k + 1) * w + w // 20)
        sp = make_spec(seed + k, per, B.LAYOUT_LAS, fmt, B.SHAPE_TERRAIN, (x0, 0, -9488), (x1, extent - 1, 19488),
                       (0.01,) * 3, (390000.0, 130000.0, 0.0), classes)
        mm, _ = device_points(ctx, sp, buf.data_ptr() + k * per * R)
        lo_all = [min(a, b) for a, b in zip(lo_all, mm[:3])]
        hi_all = [max(a, b) for a, b in zip(hi_all, mm[3:])]
    whole = make_spec(seed, n_points, layout, fmt, B.SHAPE_TERRAIN, (0, 0, -9488), (extent - 1, extent - 1, 19488),
                      (0.01,) * 3, (390000.0, 130000.0, 0.0), DOC_CLASSES)
    mm = (C.c_int32 * 6)(*lo_all, *hi_all)
    desc = B.FileDesc()
    check(lib.pcq_synth_desc(C.byref(whole), mm, C.byref(desc)))
    if layout == B.LAYOUT_LAST:
        rec = buf[: n_points * R].view(n_points, R)
        fields = [(0, 12), (12, 2), (14, 1), (15, 1), (16, 1), (17, 1), (18, 2)]
        if fmt in (1, 3):
            fields.append((20, 8))
        if fmt in (2, 3):
            fields.append((20 if fmt == 2 else 28, 6))
        cols = torch.cat([rec[:, o: o + s].contiguous().view(-1) for o, s in fields])
        del rec, buf
        buf = torch.cat([cols, torch.zeros(256, dtype=torch.uint8, device=dev)])
        del cols
    return buf, desc, header_bytes(whole, list(lo_all) + list(hi_all))
