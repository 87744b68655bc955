"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol,
its host-only entry points agree with the oracle, and the synthetic generators are sound.
No compute entry point is called here (no GPU in this container)."""
import ctypes as C
import json
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as npo
from oracle import oracle as orc
from hypothesis import given, settings
from hypothesis import strategies as st

from tests.helpers import make_file

ROOT = Path(__file__).resolve().parent.parent
KAV = json.loads((Path(__file__).parent / "golden" / "kav.json").read_text())


def _desc(pcq, f, layout=0, mask=0):
    d = pcq.FileDesc()
    pcq.binding.check(pcq.lib.pcq_parse_header(C.c_void_p(f.ctypes.data), f.nbytes, layout, mask, C.byref(d)))
    return d


def test_library_exports_every_declared_symbol(pcq):
    # include/pcq.h -> libpcq.so (the product), include/pcq_synth.h -> libpcq_synth.so (test / bench tooling)
    for hdr, so, bound, least in (("pcq.h", pcq.binding.LIB_PATH, pcq.lib._pcq_symbols, 40),
                                  ("pcq_synth.h", pcq.binding.SYNTH_LIB_PATH, pcq.binding.synth_lib._pcq_symbols, 6)):
        text = (ROOT / "include" / hdr).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared = set(re.findall(r"\b(pcq_[a-z0-9_]+)\s*\(", text))
        assert len(declared) >= least
        out = subprocess.check_output(["nm", "-D", "--defined-only", str(so)], text=True)
        exported = set(re.findall(r"\bT (pcq_[a-z0-9_]+)", out))
        assert declared <= exported, f"{hdr}: missing exports: {sorted(declared - exported)}"
        assert set(bound) <= exported


def test_synth_library_does_not_map_the_product_library(pcq):
    """The CPU reference arm of bench.py generates its inputs with libpcq_synth.so alone."""
    out = subprocess.check_output(["readelf", "-d", str(pcq.binding.SYNTH_LIB_PATH)], text=True)
    assert "libpcq.so" not in out
    und = subprocess.check_output(["nm", "-D", "--undefined-only", str(pcq.binding.SYNTH_LIB_PATH)], text=True)
    assert not re.findall(r"\bpcq_(?!synth)", und)


def test_library_is_sm100a_native(pcq):
    out = subprocess.run(["cuobjdump", "--list-elf", str(pcq.binding.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly(pcq):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pcq.PcqError) as e:
        pcq.Context(0)
    assert e.value.code == pcq.binding.PCQ_ERR_CUDA and "no CPU fallback" in e.value.message


def test_point_layout_is_readers_point(pcq):
    dt = pcq.POINT_DTYPE  # readers/src/lib.rs:10-19
    assert dt.itemsize == 31 and dt.fields["pos"][1] == 0 and dt.fields["rgb"][1] == 24 and dt.fields["cls"][1] == 30


@pytest.mark.parametrize("k", KAV["local_bounds"], ids=lambda k: k["name"])
def test_local_bounds_kav(pcq, k):
    f = make_file(np.zeros((1, 3), np.int32), [0], scale=k["scale"], offset=k["offset"])
    d = _desc(pcq, f)
    lo, hi = (C.c_int64 * 3)(), (C.c_int64 * 3)()
    pcq.binding.check(pcq.lib.pcq_local_bounds(C.byref(d), pcq.binding.d3(k["qmin"]), pcq.binding.d3(k["qmax"]), lo, hi))
    assert list(lo) == k["lo"] and list(hi) == k["hi"]


def test_local_bounds_match_oracle_randomised(pcq):
    rng = np.random.default_rng(11)
    for _ in range(300):
        scale = tuple(float(s) for s in rng.choice([0.001, 0.01, 0.00025, 0.0123456789, 1.0, 0.5], size=3))
        offset = tuple(float(o) for o in rng.uniform(-1e6, 4e6, size=3).round(3))
        f = make_file(np.zeros((1, 3), np.int32), [0], scale=scale, offset=offset)
        qmin = rng.uniform(-1e6, 4e6, size=3)
        qmax = qmin + rng.uniform(0, 1e5, size=3)
        d = _desc(pcq, f)
        lo, hi = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        rc = pcq.lib.pcq_local_bounds(C.byref(d), pcq.binding.d3(qmin), pcq.binding.d3(qmax), lo, hi)
        try:
            want = orc.local_bounds(orc.parse_header(f), qmin, qmax)
        except orc.OracleError as e:
            assert e.code == orc.ORC_ERR_PANIC and rc == pcq.binding.PCQ_ERR_PANIC
            continue
        assert rc == 0 and (list(lo), list(hi)) == want


@pytest.mark.parametrize("k", KAV["grid_params"], ids=lambda k: k["name"])
def test_grid_params_kav(pcq, k):
    dims, bits = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
    pcq.binding.check(pcq.lib.pcq_grid_params(pcq.binding.d3(k["min"]), pcq.binding.d3(k["max"]), k["cell"], dims, bits))
    assert list(dims) == k["dims"] and list(bits) == k["bits"]


def test_grid_params_too_many_cells(pcq):
    dims, bits = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
    rc = pcq.lib.pcq_grid_params(pcq.binding.d3((0, 0, 0)), pcq.binding.d3((1e9, 1e9, 1e9)), 1e-3, dims, bits)
    assert rc == pcq.binding.PCQ_ERR_GRID and b"Too many cells" in pcq.lib.pcq_last_error()


def test_parse_header_matches_oracle_and_rejects_like_it(pcq):
    xyz = np.arange(30, dtype=np.int32).reshape(10, 3)
    for fmt, ver in ((0, (1, 2)), (1, (1, 2)), (2, (1, 2)), (3, (1, 3)), (6, (1, 4)), (7, (1, 4))):
        f = make_file(xyz, np.arange(10), fmt=fmt, version=ver, scale=(0.001, 0.01, 0.1), offset=(1.5, -2.5, 3.25))
        d, h = _desc(pcq, f), orc.parse_header(f)
        assert (d.n_points, d.record_len, d.format, d.point_data_off) == (h.n_points, h.record_len, h.format, h.offset_to_point_data)
        assert list(d.scale) == list(h.scale) and list(d.offset) == list(h.offset)
        assert list(d.hdr_min) == list(h.min) and list(d.hdr_max) == list(h.max)
    f = make_file(xyz, np.arange(10), fmt=1)
    bad = f.copy()
    bad[:4] = np.frombuffer(b"LASX", np.uint8)
    d = pcq.FileDesc()
    assert pcq.lib.pcq_parse_header(C.c_void_p(bad.ctypes.data), bad.nbytes, 0, 0, C.byref(d)) == pcq.binding.PCQ_ERR_FORMAT
    assert pcq.lib.pcq_parse_header(C.c_void_p(f.ctypes.data), 100, 0, 0, C.byref(d)) == pcq.binding.PCQ_ERR_IO
    g = make_file(xyz, np.arange(10), fmt=1, fmt_byte=0x81)
    assert pcq.lib.pcq_parse_header(C.c_void_p(g.ctypes.data), g.nbytes, 1, 0, C.byref(d)) == pcq.binding.PCQ_ERR_FORMAT
    assert pcq.lib.pcq_parse_header(C.c_void_p(g.ctypes.data), g.nbytes, 1, 1, C.byref(d)) == 0 and d.format == 1
    short = make_file(xyz, np.arange(10), fmt=3, record_len=34)
    short[105] = 20  # record length smaller than format 3 needs
    assert pcq.lib.pcq_parse_header(C.c_void_p(short.ctypes.data), short.nbytes, 0, 0, C.byref(d)) == pcq.binding.PCQ_ERR_FORMAT


def test_file_intersects(pcq):
    f = make_file(np.array([[0, 0, 0], [100, 100, 100]], np.int32), [1, 1], scale=(1.0,) * 3)
    d = _desc(pcq, f)
    out = C.c_int()
    for qmin, qmax, want in (((100, 100, 100), (200, 200, 200), 1),  # touching faces overlap (closed intervals)
                             ((100.5, 0, 0), (200, 200, 200), 0), ((-5, -5, -5), (-1, 200, 200), 0), ((50, 50, 50), (60, 60, 60), 1)):
        pcq.binding.check(pcq.lib.pcq_file_intersects(C.byref(d), pcq.binding.d3(qmin), pcq.binding.d3(qmax), C.byref(out)))
        assert out.value == want
    inv = make_file(np.zeros((1, 3), np.int32), [1], hdr_min=(5, 0, 0), hdr_max=(1, 1, 1))
    rc = pcq.lib.pcq_file_intersects(C.byref(_desc(pcq, inv)), pcq.binding.d3((0, 0, 0)), pcq.binding.d3((1, 1, 1)), C.byref(out))
    assert rc == pcq.binding.PCQ_ERR_PANIC


# ---- synthetic data --------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_synth_las_and_last_hold_the_same_points(pcq, fmt):
    B, S = pcq.binding, pcq.synth
    a = S.host_file(S.uniform_spec(2000, B.LAYOUT_LAS, fmt, seed=5))
    b = S.host_file(S.uniform_spec(2000, B.LAYOUT_LAST, fmt, seed=5))
    assert np.array_equal(a[:227], b[:227])
    for klass in (2, 6):
        assert np.array_equal(npo.search_class(a, "las", klass).view(np.uint8), npo.search_class(b, "last", klass).view(np.uint8))
    h = orc.parse_header(a)
    xyz, cls, rgb = npo._columns(a, npo.parse_header(a), "las")
    pos = xyz.astype(np.float64) * np.array(h.scale[:]) + np.array(h.offset[:])
    assert np.array_equal(pos.min(axis=0), np.array(h.min[:])) and np.array_equal(pos.max(axis=0), np.array(h.max[:]))
    assert h.header_size == 227 and h.offset_to_point_data == 227 and (h.version_major, h.version_minor) == (1, 2)


def test_synth_is_deterministic_and_seed_sensitive(pcq):
    B, S = pcq.binding, pcq.synth
    a = S.host_file(S.uniform_spec(1000, B.LAYOUT_LAS, 1, seed=5))
    assert np.array_equal(a, S.host_file(S.uniform_spec(1000, B.LAYOUT_LAS, 1, seed=5)))
    assert not np.array_equal(a, S.host_file(S.uniform_spec(1000, B.LAYOUT_LAS, 1, seed=6)))


def test_synth_shapes_and_class_mix(pcq):
    B, S = pcq.binding, pcq.synth
    sp = S.doc_specs(n_files=4, pts_per_file=20000)[3]
    f = S.host_file(sp)
    h = npo.parse_header(f)
    xyz, cls, _ = npo._columns(f, h, "las")
    base = cls & 0x1F
    frac2 = float((base == 2).mean())
    assert 0.40 < frac2 < 0.50  # DOC_CLASSES: 45 % ground
    assert 0.003 < float((cls > 0x1F).mean()) < 0.03  # ~1 % carry a flag bit: whole-byte compare must see them
    z = xyz[:, 2] * 0.01
    assert -95 <= z.min() and z.max() <= 195 and 30 < float(np.median(z)) < 70  # bell around 50 m
    nav = S.host_file(S.navvis_spec(n_points=30000))
    nx, _, nrgb = npo._columns(nav, npo.parse_header(nav), "las")
    assert nrgb is not None and len(np.unique(nx[:, 2])) < 20000  # floors: many points share a z level
    ca = S.host_file(S.ca13_specs(n_files=4, pts_per_file=5000)[0])
    assert npo.parse_header(ca)["format"] == 1 and len(npo.search_class(ca, "last", 2)) > 2000


def test_doc_tiles_cover_the_xl_box_and_boxes_select_tiles(pcq):
    S = pcq.synth
    specs = S.doc_specs(n_files=64, pts_per_file=0)
    hits = {"S": 0, "L": 0, "XL": 0}
    for sp in specs:
        x0, y0 = sp.offset[0], sp.offset[1]
        x1, y1 = x0 + sp.hi[0] * sp.scale[0], y0 + sp.hi[1] * sp.scale[1]
        for name, (qmin, qmax) in (("S", S.DOC_S), ("L", S.DOC_L), ("XL", S.DOC_XL)):
            if x0 <= qmax[0] and x1 >= qmin[0] and y0 <= qmax[1] and y1 >= qmin[1]:
                hits[name] += 1
    assert hits == {"S": 5, "L": 30, "XL": 64}


# ---- chunk index: the host-side filter (pcq_index_filter) against the per-point search of the oracle --------------
def _strip_file(rng, n, layout, ch):
    from tests.helpers import make_file

    i = np.arange(n)
    xyz = np.stack([i * 40 + rng.integers(-3000, 3000, n), rng.integers(-50_000, 150_000, n), rng.integers(-500, 5000, n)], axis=1).astype(np.int32)
    cls = rng.choice(np.array([1, 2, 2, 5, 6 | 0x20], np.uint8), size=n)
    cls[((i // ch) % 7 == 3) & (rng.random(n) < 0.02)] = 6  # class 6 only in every 7th chunk
    return make_file(xyz, cls, fmt=1, scale=(0.01, 0.01, 0.01), offset=(390000.0, 130000.0, 0.0), layout=layout, seed=5), xyz, cls


@pytest.mark.parametrize("layout", ["las", "last"])
def test_index_filter_is_sound_and_joins_runs(pcq, layout):
    from oracle import np_oracle as npo

    ch = pcq.binding.INDEX_CHUNK_POINTS
    rng = np.random.default_rng(17)
    n = 40 * ch + 321
    f, xyz, cls = _strip_file(rng, n, layout, ch)
    headers = npo.chunk_headers(f, layout, chunk_points=ch)
    desc = _desc(pcq, f, 0 if layout == "las" else 1, 1)
    n_chunks = headers.shape[0]
    assert n_chunks == 41

    def chunks_of(runs):
        keep = np.zeros(n_chunks, dtype=bool)
        for a, b in runs:
            assert 0 <= a < b <= n_chunks
            keep[a:b] = True
        return keep

    # bounds: a slab in x; every chunk that holds a match of the per-point search must be inside a run
    for x0, x1 in ((5.2 * ch * 40, 7.9 * ch * 40), (0, 1), (-1e9, 1e9), (39.5 * ch * 40, 1e9)):
        qmin, qmax = (390000.0 + x0 * 0.01, 0.0, -1e6), (390000.0 + x1 * 0.01, 1e7, 1e6)
        s = pcq.BoundsSearcher(qmin, qmax)
        runs, may = pcq.index_filter(headers, desc, s, join_gap=1)
        hd = npo.parse_header(f)
        lo, hi = npo.local_bounds(hd, qmin, qmax)
        m = np.all((xyz.astype(np.int64) >= np.array(lo)) & (xyz.astype(np.int64) <= np.array(hi)), axis=1)
        assert m.sum() == npo.search_bounds(f, layout, qmin, qmax).shape[0]
        keep = chunks_of(runs)
        assert keep.sum() == may
        assert not m[~np.repeat(keep, ch)[:n]].any(), "a matching point lies in a chunk the filter dropped"
        if x1 - x0 < 3 * ch * 40:
            assert 0 < may < 6
        # runs are maximal and disjoint at gap 1
        assert all(runs[i][1] < runs[i + 1][0] for i in range(len(runs) - 1))
    # class: chunks 3, 10, 17, 24, 31, 38 hold class 6; the flag-bit variant 6|0x20 is another byte (las.rs:229)
    runs, may = pcq.index_filter(headers, desc, pcq.ClassSearcher(6), join_gap=1)
    assert may == 6 and [tuple(r) for r in runs] == [(c, c + 1) for c in (3, 10, 17, 24, 31, 38)]
    runs, may = pcq.index_filter(headers, desc, pcq.ClassSearcher(6), join_gap=8)   # 7 chunks apart: joined
    assert may == 6 and [tuple(r) for r in runs] == [(3, 39)]
    runs, may = pcq.index_filter(headers, desc, pcq.ClassSearcher(6), join_gap=7)   # gap of 6 chunks < 7: joined too
    assert [tuple(r) for r in runs] == [(3, 39)]
    runs, may = pcq.index_filter(headers, desc, pcq.ClassSearcher(6), join_gap=6)
    assert len(runs) == 6
    runs, may = pcq.index_filter(headers, desc, pcq.ClassSearcher(19), join_gap=4)
    assert may == 0 and len(runs) == 0
    runs, may = pcq.index_filter(headers, desc, pcq.ClassSearcher(6 | 0x20), join_gap=4)
    assert may == n_chunks and [tuple(r) for r in runs] == [(0, n_chunks)]
    # a query the header box already excludes: no runs at all (the early-out of las.rs:82-84)
    runs, may = pcq.index_filter(headers, desc, pcq.BoundsSearcher((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)))
    assert may == 0 and len(runs) == 0
    # inverted box: the reference panics in AABB::from_min_max
    with pytest.raises(pcq.PcqError) as e:
        pcq.index_filter(headers, desc, pcq.BoundsSearcher((5.0, 0.0, 0.0), (1.0, 1.0, 1.0)))
    assert e.value.code == pcq.binding.PCQ_ERR_PANIC


def test_index_filter_property_random_boxes(pcq):
    """Random boxes (tiny, huge, far outside the i32 range, anisotropic scales where the reference's min-y / min-z quirk
    inverts the local box): pcq_index_filter either fails exactly where the per-point search fails, or keeps every
    chunk that holds a match."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from oracle import np_oracle as npo
    from tests.helpers import make_file

    ch = 256
    rng = np.random.default_rng(99)
    n = 20 * ch + 7
    i = np.arange(n)
    xyz = np.stack([i * 13 - 20000 + rng.integers(-400, 400, n), rng.integers(-30000, 30000, n), (i % 977) * 31 - 9000], axis=1).astype(np.int32)
    cls = rng.integers(0, 8, size=n).astype(np.uint8)
    files = {}
    for name, scale in (("iso", (0.01, 0.01, 0.01)), ("aniso", (0.001, 0.002, 0.00025))):
        f = make_file(xyz, cls, fmt=1, scale=scale, offset=(1000.0, -2000.0, 50.0), layout="las", seed=2)
        files[name] = (f, npo.chunk_headers(f, "las", chunk_points=ch), _desc(pcq, f, 0, 1), npo.parse_header(f))
    def axis(scale_a, off_a, lo_raw, hi_raw):
        w0, w1 = off_a + lo_raw * scale_a, off_a + hi_raw * scale_a
        span = w1 - w0
        return st.one_of(st.floats(w0 - span, w1 + span), st.floats(-1e13, 1e13), st.sampled_from([off_a, w0, w1, 1e300, -1e300]))

    def corners(scale):
        return st.tuples(axis(scale[0], 1000.0, -20400, 49000), axis(scale[1], -2000.0, -30000, 30000), axis(scale[2], 50.0, -9000, 21300))

    cases = st.one_of(st.tuples(st.just("iso"), corners((0.01, 0.01, 0.01)), corners((0.01, 0.01, 0.01))),
                      st.tuples(st.just("aniso"), corners((0.001, 0.002, 0.00025)), corners((0.001, 0.002, 0.00025))))
    seen = {"panic": 0, "match": 0, "skipped": 0}

    @settings(max_examples=400, deadline=None, derandomize=True, database=None)
    @given(cases, st.integers(1, 5))
    def run(case, gap):
        which, a, b = case
        f, headers, desc, hd = files[which]
        qmin = tuple(min(x, y) for x, y in zip(a, b))
        qmax = tuple(max(x, y) for x, y in zip(a, b))
        try:
            want = npo.search_bounds(f, "las", qmin, qmax)
            panicked = False
        except npo.Panic:
            panicked = True
        try:
            runs, may = pcq.index_filter(headers, desc, pcq.BoundsSearcher(qmin, qmax), join_gap=gap)
        except pcq.PcqError as e:
            assert panicked and e.code == pcq.binding.PCQ_ERR_PANIC
            seen["panic"] += 1
            return
        assert not panicked
        keep = np.zeros(headers.shape[0], dtype=bool)
        for r0, r1 in runs:
            keep[r0:r1] = True
        if want.shape[0]:
            lo, hi = npo.local_bounds(hd, qmin, qmax)
            x = xyz.astype(np.int64)
            lo64 = np.array([max(min(int(v), 1 << 62), -(1 << 62)) for v in lo], dtype=np.int64)
            hi64 = np.array([max(min(int(v), 1 << 62), -(1 << 62)) for v in hi], dtype=np.int64)
            m = np.all((x >= lo64) & (x <= hi64), axis=1)
            assert m.sum() == want.shape[0]
            assert not m[~np.repeat(keep, ch)[:n]].any()
            seen["match"] += 1
            seen["skipped"] += int(keep.sum() < headers.shape[0])
        assert may <= keep.sum()

    run()
    assert seen["match"] > 30 and seen["skipped"] > 10, seen  # the strategy reaches boxes with matches and with skipped chunks
    # a box that intersects the anisotropic file but whose local y range inverts (min uses the x scale, las.rs:91-92)
    f, headers, desc, hd = files["aniso"]
    qmin, qmax = (990.0, -1990.0, 40.0), (1040.0, -1985.0, 60.0)
    with pytest.raises(npo.Panic):
        npo.search_bounds(f, "las", qmin, qmax)
    with pytest.raises(pcq.PcqError) as e:
        pcq.index_filter(headers, desc, pcq.BoundsSearcher(qmin, qmax))
    assert e.value.code == pcq.binding.PCQ_ERR_PANIC


# ---- the sub-box a per-file grid's table covers (pcq_grid_cells_under_box) ---------------------------------------------
@settings(max_examples=150, deadline=None, derandomize=True)
@given(st.tuples(st.floats(-1e5, 1e5), st.floats(-1e5, 1e5), st.floats(-1e3, 1e3)),
       st.tuples(st.floats(0.5, 5e4), st.floats(0.5, 5e4), st.floats(0.5, 2e3)),
       st.sampled_from([0.01, 0.1, 0.37, 1.0, 25.0, 100.0]),
       st.tuples(st.floats(-0.3, 1.2), st.floats(-0.3, 1.2), st.floats(-0.3, 1.2)),
       st.tuples(st.floats(0.0, 0.8), st.floats(0.0, 0.8), st.floats(0.0, 0.8)),
       st.integers(0, 2**31 - 1))
def test_cells_under_a_box_hold_every_position_of_the_box(pcq, gmin, ext, cell, rel_lo, rel_len, seed):
    """Soundness of the table sizing: a position inside the box falls into a cell of the reported range on every axis,
    unless the cell exceeds the axis' key mask (aliased: such points bypass the table).  Cells as the numpy restatement
    of SparseGrid computes them (grid_sampling.rs:51-60)."""
    from oracle import np_oracle

    gmax = tuple(gmin[a] + ext[a] for a in range(3))
    try:
        grid = np_oracle.SparseGrid(gmin, gmax, cell)
    except ValueError:
        return
    if any(b > 62 for b in grid.bits):
        return
    box_min = tuple(gmin[a] + rel_lo[a] * ext[a] for a in range(3))
    box_max = tuple(box_min[a] + rel_len[a] * ext[a] for a in range(3))
    B = pcq.binding
    lo, n = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
    B.check(B.lib.pcq_grid_cells_under_box(B.d3(gmin), B.d3(gmax), float(cell), B.d3(box_min), B.d3(box_max), lo, n))
    rng = np.random.default_rng(seed)
    pts = [box_min, box_max] + [tuple(box_min[a] + rng.random() * (box_max[a] - box_min[a]) for a in range(3)) for _ in range(40)]
    for p in pts:
        c = grid._cell(p)
        for a in range(3):
            if c[a] > (1 << grid.bits[a]) - 1:
                continue
            assert lo[a] <= c[a] < lo[a] + n[a], (a, p, c, list(lo), list(n))
    # a box that covers a small part of the grid gets a small range
    for a in range(3):
        if 0.0 <= rel_lo[a] and rel_lo[a] + rel_len[a] <= 1.0 and grid.dims[a] > 0:
            assert n[a] <= rel_len[a] * grid.dims[a] + 4
