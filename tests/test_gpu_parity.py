"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bit-exact everywhere: counts, scan-ordered record streams (BufferCollector) and record sets
(GridSampledCollector, whose reference order is HashMap order)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import make_file, same_point_seq, same_point_set

pytestmark = pytest.mark.gpu

SI = None  # SearchImplementation.Optimized, bound in _impl()


def _impl(pcq):
    return pcq.SearchImplementation.Optimized


def oracle_run(files, exts, kind, bounds=None, cls=None, grid=None, per_file=False):
    """run_search_sequential / run_search_parallel on the oracle -> list of collectors"""
    def new():
        if kind == orc.COLLECT_GRID:
            return orc.Collector(kind, grid[0], grid[1], grid[2])
        return orc.Collector(kind)

    if per_file:
        cols = [new() for _ in files]
        for f, e, c in zip(files, exts, cols):
            orc.search_file(f, e, c, bounds=bounds, cls=cls)
        return cols
    c = new()
    for f, e in zip(files, exts):
        orc.search_file(f, e, c, bounds=bounds, cls=cls)
    return [c]


def gpu_run(pcq, ctx, files, exts, kind, bounds=None, cls=None, grid=None, per_file=False, host_stream=False):
    def new():
        if kind == orc.COLLECT_GRID:
            return pcq.GridSampledCollector(grid[0], grid[1], grid[2], ctx=ctx)
        return (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)

    searcher = pcq.BoundsSearcher(*bounds) if bounds is not None else pcq.ClassSearcher(cls)
    cols = [new() for _ in files] if per_file else [new()]
    images = list(zip(files, exts))
    if host_stream:
        searcher.search_host_files(images, cols)
    else:
        searcher.search_files(images, _impl(pcq), cols)
    return cols


def assert_same(kind, got_cols, want_cols):
    assert len(got_cols) == len(want_cols)
    for g, w in zip(got_cols, want_cols):
        assert g.point_count() == w.point_count()
        if kind == orc.COLLECT_COUNT:
            assert g.points() is None
        elif kind == orc.COLLECT_BUFFER:
            assert same_point_seq(g.points(), w.points())
        else:
            assert same_point_set(g.points(), w.points())


HEADERS = [((0.01, 0.01, 0.01), (390000.0, 130000.0, 0.0)),
           ((0.001, 0.001, 0.001), (-23.108, -21.261, -10.029)),        # FMA-sensitive (KAV-4b)
           ((0.001, 0.002, 0.00025), (643431.76, 3883547.565, 1234.5))]  # FMA-sensitive + anisotropic


def random_file(rng, n, fmt, layout, hdr=0, record_len=None):
    scale, offset = HEADERS[hdr]
    xyz = rng.integers(-50_000, 150_000, size=(n, 3), dtype=np.int32)
    cls = rng.choice(np.array([1, 2, 2, 2, 5, 6, 6 | 0x20, 2 | 0x80, 9], np.uint8), size=n)
    return make_file(xyz, cls, fmt=fmt, scale=scale, offset=offset, layout=layout, record_len=record_len,
                     seed=int(rng.integers(1 << 30)))


def box(hdr, lo=100, hi=90_000):
    scale, offset = HEADERS[hdr]
    return ([offset[a] + lo * scale[0] for a in range(3)], [offset[a] + hi * scale[a] for a in range(3)])


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [1, 2], ids=["direct", "staged"])
@pytest.mark.parametrize("layout", ["las", "last"])
@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_bounds_and_class_all_collectors(pcq, ctx, variant, layout, fmt):
    ctx.set_scan_variant(variant)
    rng = np.random.default_rng(100 * fmt + (layout == "last"))
    hdr = fmt % 3
    files = [random_file(rng, n, fmt, layout, hdr) for n in (5000, 513, 1)]
    exts = [layout] * 3
    b = box(hdr)
    grid = (b[0], b[1], (b[1][0] - b[0][0]) / 37.0)
    for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER, orc.COLLECT_GRID):
        for per_file in (False, True):
            want = oracle_run(files, exts, kind, bounds=b, grid=grid, per_file=per_file)
            got = gpu_run(pcq, ctx, files, exts, kind, bounds=b, grid=grid, per_file=per_file)
            assert_same(kind, got, want)
    hmin = np.min([orc.parse_header(f).min[:] for f in files], axis=0)
    hmax = np.max([orc.parse_header(f).max[:] for f in files], axis=0)
    cgrid = (list(hmin), list(hmax + 1.0), float((hmax[0] - hmin[0]) / 29.0))  # union of header bounds (main.rs:94-120), padded
    for klass in (2, 6, 19, 2 | 0x80):
        for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER, orc.COLLECT_GRID):
            want = oracle_run(files, exts, kind, cls=klass, grid=cgrid, per_file=True)
            got = gpu_run(pcq, ctx, files, exts, kind, cls=klass, grid=cgrid, per_file=True)
            assert_same(kind, got, want)
    ctx.set_scan_variant(0)


@pytest.mark.parametrize("n", [0, 1, 31, 255, 256, 511, 512, 513, 1023, 1025, 100_003])
def test_ragged_sizes(pcq, ctx, n):
    rng = np.random.default_rng(n)
    for layout in ("las", "last"):
        f = random_file(rng, n, 1, layout, 1)
        b = box(1, -60_000, 160_000)  # everything matches
        for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER):
            want = oracle_run([f], [layout], kind, bounds=b)
            got = gpu_run(pcq, ctx, [f], [layout], kind, bounds=b)
            assert want[0].point_count() == n
            assert_same(kind, got, want)


@pytest.mark.parametrize("layout", ["las", "last"])
def test_select_many_small_files(pcq, ctx, layout):
    """look-back units never straddle files: 150 files of ragged sizes (empty ones included), one collector over all
    of them and one collector per file, bounds and class queries (LAST class queries take k_select_bytes)"""
    rng = np.random.default_rng(2024 + (layout == "last"))
    sizes = [int(v) for v in rng.integers(0, 6000, size=146)] + [0, 1, 40_000, 70_001]
    files = [random_file(rng, n, 1 + 2 * (k % 2), layout, 1) for k, n in enumerate(sizes)]
    exts = [layout] * len(files)
    b = box(1, 20_000, 120_000)
    for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER):
        for per_file in (False, True):
            want = oracle_run(files, exts, kind, bounds=b, per_file=per_file)
            got = gpu_run(pcq, ctx, files, exts, kind, bounds=b, per_file=per_file)
            assert_same(kind, got, want)
            for klass in (2, 19):
                want = oracle_run(files, exts, kind, cls=klass, per_file=per_file)
                got = gpu_run(pcq, ctx, files, exts, kind, cls=klass, per_file=per_file)
                assert_same(kind, got, want)


def test_last_class_select_multi_unit(pcq, ctx):
    """k_select_bytes: several 32768-point units per CTA, a ragged tail, host-streamed chunks and a class column
    whose address is not 16-byte aligned (falls back to k_select)"""
    import torch

    rng = np.random.default_rng(77)
    n = 3_000_017
    f = random_file(rng, n, 3, "last", 2)
    for klass in (2, 6, 19):
        want = oracle_run([f], ["last"], orc.COLLECT_BUFFER, cls=klass)
        got = gpu_run(pcq, ctx, [f], ["last"], orc.COLLECT_BUFFER, cls=klass)
        assert_same(orc.COLLECT_BUFFER, got, want)
    want = oracle_run([f], ["last"], orc.COLLECT_BUFFER, cls=6)
    got = gpu_run(pcq, ctx, [f], ["last"], orc.COLLECT_BUFFER, cls=6, host_stream=True)
    assert_same(orc.COLLECT_BUFFER, got, want)
    # the whole transposed block wrapped in place: the class column starts at base + 15 N (odd)
    desc = pcq.FileDesc()
    pcq.binding.check(pcq.lib.pcq_parse_header(C.c_void_p(f.ctypes.data), f.nbytes, 1, 1, C.byref(desc)))
    body = torch.from_numpy(f[desc.point_data_off:].copy()).to(f"cuda:{ctx.device}")
    df = pcq.DeviceFile.wrap(ctx, desc, body.data_ptr(), keepalive=body)
    c = pcq.BufferCollector(ctx)
    pcq.ClassSearcher(6).search_files([df], _impl(pcq), [c])
    assert same_point_seq(c.points(), want[0].points())


@pytest.mark.parametrize("fmt", [0, 1])
def test_last_class_select_dense_colourless(pcq, ctx, fmt):
    """k_select_bytes' cp.async gather emit (colourless LAST, more than 128 matches in a warp's 4096 points): density
    varies along the file from 0 to 100 %, so warps of one unit take different paths; ragged tail; small output
    buffer first (overflow re-run); several files on one collector and one per file."""
    rng = np.random.default_rng(300 + fmt)
    files = []
    for n in (1_000_003, 70_001, 129):
        xyz = rng.integers(-50_000, 150_000, size=(n, 3), dtype=np.int32)
        p = np.clip(np.sin(np.arange(n) / 9000.0) * 0.7 + 0.5, 0.0, 1.0)  # match probability along the file
        cls = np.where(rng.random(n) < p, 2, rng.choice(np.array([1, 5, 6, 2 | 0x40], np.uint8), size=n)).astype(np.uint8)
        files.append(make_file(xyz, cls, fmt=fmt, scale=HEADERS[1][0], offset=HEADERS[1][1], layout="last", seed=n))
    exts = ["last"] * 3
    for per_file in (False, True):
        for klass in (2, 6):
            want = oracle_run(files, exts, orc.COLLECT_BUFFER, cls=klass, per_file=per_file)
            got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_BUFFER, cls=klass, per_file=per_file)
            assert_same(orc.COLLECT_BUFFER, got, want)
    want = oracle_run(files, exts, orc.COLLECT_BUFFER, cls=2)
    got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_BUFFER, cls=2, host_stream=True)
    assert_same(orc.COLLECT_BUFFER, got, want)


@pytest.mark.parametrize("layout,fmt,unit", [("las", 0, 4096), ("las", 1, 3072), ("las", 2, 3072), ("las", 3, 2048),
                                             ("last", 1, 7168), ("last", 3, 7168)])
def test_select_around_unit_boundaries(pcq, ctx, layout, fmt, unit):
    """file sizes just below / at / above multiples of the look-back unit of k_select_ring (and of k_select's 2048),
    sparse and dense boxes, every file its own collector and all files in one collector"""
    rng = np.random.default_rng(unit + fmt)
    sizes = [unit - 1, unit, unit + 1, 2 * unit + 17, 3 * unit - 1, 2047, 2049, 5 * unit]
    files = [random_file(rng, n, fmt, layout, fmt % 2) for n in sizes]
    exts = [layout] * len(files)
    for lo, hi in ((100, 90_000), (40_000, 60_000), (-60_000, 160_000)):  # ~60 %, ~1 %, everything
        b = box(fmt % 2, lo, hi)  # (isotropic headers: the x-scale quirk would turn these boxes inside out)
        for per_file in (False, True):
            want = oracle_run(files, exts, orc.COLLECT_BUFFER, bounds=b, per_file=per_file)
            for variant in (2, 1):
                ctx.set_scan_variant(variant)
                got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_BUFFER, bounds=b, per_file=per_file)
                assert_same(orc.COLLECT_BUFFER, got, want)
    ctx.set_scan_variant(0)
    want = oracle_run(files, exts, orc.COLLECT_BUFFER, cls=2)
    got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_BUFFER, cls=2)
    assert_same(orc.COLLECT_BUFFER, got, want)


def test_odd_record_lengths_and_extended_formats(pcq, ctx):
    rng = np.random.default_rng(5)
    for fmt, rl, ver in ((0, 23, (1, 2)), (1, 31, (1, 2)), (2, 29, (1, 2)), (3, 40, (1, 3)), (6, 30, (1, 4)), (7, 36, (1, 4))):
        for layout in ("las", "last"):
            xyz = rng.integers(0, 1000, size=(3000, 3), dtype=np.int32)
            cls = rng.integers(0, 8, size=3000).astype(np.uint8)
            f = make_file(xyz, cls, fmt=fmt, layout=layout, record_len=rl, version=ver, scale=(0.5, 0.25, 0.125))
            b = ([100.0, 50.0, 20.0], [400.0, 200.0, 100.0])
            for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER):
                assert_same(kind, gpu_run(pcq, ctx, [f], [layout], kind, bounds=b), oracle_run([f], [layout], kind, bounds=b))
                assert_same(kind, gpu_run(pcq, ctx, [f], [layout], kind, cls=3), oracle_run([f], [layout], kind, cls=3))


def test_kav_on_device(pcq, ctx):
    # KAV-2: raw z = 5959 sits on the query's max face and is rejected (5.959/0.001 truncates to 5958)
    xyz = np.array([[0, 0, 5959], [0, 0, 5958]], np.int32)
    f = make_file(xyz, [1, 1], fmt=0, scale=(0.001,) * 3)
    got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_BUFFER, bounds=((-23.108, -21.261, -10.029), (28.588, 27.123, 5.959)))
    assert got[0].point_count() == 1 and got[0].points()["pos"][0][2] == 5958 * 0.001
    # KAV-4: reconstruction must not be fused
    for scale, offset, raw, want in ((0.001, 643431.76, 74891984, "0x1.5ebe77ced9168p+19"), (0.001, -23.108, 22435, "-0x1.589374bc6a800p-1"),
                                     (0.00025, 1234.5, -69419388, "-0x1.f7c2c6a7ef9dcp+13"), (0.0123456789, 390000.0, 1840728, "0x1.930d425b6d384p+18")):
        f = make_file(np.array([[raw] * 3], np.int32), [2], fmt=0, scale=(scale,) * 3, offset=(offset,) * 3)
        for variant in (1, 2):
            ctx.set_scan_variant(variant)
            got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_BUFFER, cls=2)
            assert float(got[0].points()["pos"][0][0]).hex() == float.fromhex(want).hex()
    ctx.set_scan_variant(0)


def test_error_behaviour_matches_reference(pcq, ctx):
    B = pcq.binding
    xyz = np.array([[1, 1, 1]], np.int32)
    inv = make_file(xyz, [2], fmt=1, hdr_min=(5, 0, 0), hdr_max=(1, 1, 1))
    with pytest.raises(pcq.PcqError) as e:
        gpu_run(pcq, ctx, [inv], ["las"], orc.COLLECT_COUNT, bounds=((0, 0, 0), (1, 1, 1)))
    assert e.value.code == B.PCQ_ERR_PANIC
    gpu_run(pcq, ctx, [inv], ["las"], orc.COLLECT_COUNT, cls=2)  # class search never looks at the bounds
    aniso = make_file(xyz, [2], fmt=1, scale=(0.01, 1.0, 1.0), hdr_min=(-100,) * 3, hdr_max=(100,) * 3)
    with pytest.raises(pcq.PcqError) as e:  # local min (x-scaled) > local max: AABB::<i64>::from_min_max panics
        gpu_run(pcq, ctx, [aniso], ["las"], orc.COLLECT_COUNT, bounds=((0, 5, 5), (10, 10, 10)))
    assert e.value.code == B.PCQ_ERR_PANIC
    with pytest.raises(pcq.PcqError) as e:
        gpu_run(pcq, ctx, [aniso], ["las"], orc.COLLECT_COUNT, bounds=((5, 0, 0), (1, 1, 1)))
    assert e.value.code == B.PCQ_ERR_PANIC
    masked = make_file(xyz, [2], fmt=1, layout="last", fmt_byte=0x81)
    assert gpu_run(pcq, ctx, [masked], ["last"], orc.COLLECT_COUNT, cls=2)[0].point_count() == 1  # last.rs:222
    with pytest.raises(pcq.PcqError) as e:
        gpu_run(pcq, ctx, [masked], ["last"], orc.COLLECT_COUNT, bounds=((0, 0, 0), (1, 1, 1)))
    assert e.value.code == B.PCQ_ERR_FORMAT
    with pytest.raises(pcq.PcqError) as e:
        pcq.DeviceFile.stage(ctx, make_file(xyz, [2], fmt=1)[:-20].copy(), "las")
    assert e.value.code == B.PCQ_ERR_IO
    with pytest.raises(pcq.PcqError) as e:
        pcq.DeviceFile.stage(ctx, make_file(xyz, [2], fmt=1), "laz")
    assert e.value.code == B.PCQ_ERR_FORMAT
    with pytest.raises(pcq.PcqError) as e:
        pcq.GridSampledCollector((0, 0, 0), (1e9, 1e9, 1e9), 1e-3, ctx=ctx)
    assert e.value.code == B.PCQ_ERR_GRID
    with pytest.raises(NotImplementedError):
        pcq.BoundsSearcher((0, 0, 0), (1, 1, 1)).search_file((inv, "las"), pcq.SearchImplementation.Regular, pcq.CountCollector(ctx))


def test_file_outside_query_and_out_of_i32_range(pcq, ctx):
    rng = np.random.default_rng(9)
    f = random_file(rng, 2000, 1, "las", 0)
    far = ([0.0, 0.0, 0.0], [10.0, 10.0, 10.0])
    assert gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_COUNT, bounds=far)[0].point_count() == 0
    huge = make_file(rng.integers(0, 100, size=(100, 3), dtype=np.int32), [1] * 100, scale=(1e-9,) * 3,
                     hdr_min=(-1e12,) * 3, hdr_max=(1e12,) * 3)
    b = ([1e6, 1e6, 1e6], [1e7, 1e7, 1e7])  # local bounds far beyond i32: nothing can match, must not wrap
    assert_same(orc.COLLECT_COUNT, gpu_run(pcq, ctx, [huge], ["las"], orc.COLLECT_COUNT, bounds=b),
                oracle_run([huge], ["las"], orc.COLLECT_COUNT, bounds=b))
    b = ([-1e6, -1e6, -1e6], [1e7, 1e7, 1e7])  # local bounds clamp to the whole i32 range: everything matches
    want = oracle_run([huge], ["las"], orc.COLLECT_COUNT, bounds=b)
    assert want[0].point_count() == 100
    assert_same(orc.COLLECT_COUNT, gpu_run(pcq, ctx, [huge], ["las"], orc.COLLECT_COUNT, bounds=b), want)


def test_collectors_accumulate_across_calls(pcq, ctx):
    rng = np.random.default_rng(21)
    files = [random_file(rng, 3000 + i, 3, "las", 1) for i in range(4)]
    b = box(1)
    grid = (b[0], b[1], (b[1][0] - b[0][0]) / 23.0)
    for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER, orc.COLLECT_GRID):
        want = oracle_run(files, ["las"] * 4, kind, bounds=b, grid=grid)
        if kind == orc.COLLECT_GRID:
            col = pcq.GridSampledCollector(grid[0], grid[1], grid[2], ctx=ctx)
        else:
            col = (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)
        s = pcq.BoundsSearcher(*b)
        for f in files:  # one search_file call per file, as run_search_sequential does (main.rs:131-133)
            s.search_file((f, "las"), _impl(pcq), col)
        assert_same(kind, [col], want)
        col.reset()
        assert col.point_count() == 0
        s.search_files([(f, "las") for f in files], _impl(pcq), [col])
        assert_same(kind, [col], want)


def test_buffer_grows_when_everything_matches(pcq, ctx):
    rng = np.random.default_rng(33)
    n = 3_000_000  # beyond the initial capacity guess -> exercises the grow-and-rerun path
    xyz = rng.integers(0, 1000, size=(n, 3), dtype=np.int32)
    f = make_file(xyz, np.full(n, 2, np.uint8), fmt=2, scale=(0.01,) * 3)
    want = oracle_run([f], ["las"], orc.COLLECT_BUFFER, cls=2)
    got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_BUFFER, cls=2)
    assert want[0].point_count() == n
    assert_same(orc.COLLECT_BUFFER, got, want)


def test_host_streaming_equals_resident(pcq, ctx):
    rng = np.random.default_rng(41)
    os.environ["PCQ_CHUNK_MB"] = "1"  # many chunks per file
    try:
        for layout in ("las", "last"):
            files = [random_file(rng, 150_000 + 777 * i, 3, layout, 1) for i in range(3)]
            exts = [layout] * 3
            b = box(1)
            grid = (b[0], b[1], (b[1][0] - b[0][0]) / 19.0)
            for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER, orc.COLLECT_GRID):
                for per_file in (False, True):
                    want = oracle_run(files, exts, kind, bounds=b, grid=grid, per_file=per_file)
                    got = gpu_run(pcq, ctx, files, exts, kind, bounds=b, grid=grid, per_file=per_file, host_stream=True)
                    assert_same(kind, got, want)
            want = oracle_run(files, exts, orc.COLLECT_COUNT, cls=2, per_file=True)
            assert_same(orc.COLLECT_COUNT, gpu_run(pcq, ctx, files, exts, orc.COLLECT_COUNT, cls=2, per_file=True, host_stream=True), want)
            # a batch of queries sharing one staging pass (pcq_search_host_files_multi)
            b2 = box(1, 20_000, 60_000)
            searchers = [pcq.BoundsSearcher(*b), pcq.ClassSearcher(6), pcq.BoundsSearcher(*b2)]
            wants = [oracle_run(files, exts, orc.COLLECT_BUFFER, bounds=b, per_file=True), oracle_run(files, exts, orc.COLLECT_BUFFER, cls=6, per_file=True),
                     oracle_run(files, exts, orc.COLLECT_BUFFER, bounds=b2, per_file=True)]
            cols = [[pcq.BufferCollector(ctx) for _ in files] for _ in searchers]
            pcq.search_host_files_multi(list(zip(files, exts)), searchers, cols)
            for got, want in zip(cols, wants):
                assert_same(orc.COLLECT_BUFFER, got, want)
    finally:
        del os.environ["PCQ_CHUNK_MB"]


# ---- density ---------------------------------------------------------------------------------------
def test_reference_sparse_grid_tests_on_device(pcq, ctx):
    # grid_sampling.rs:121-208 driven through the scan: points sit exactly on the tested coordinates
    gmin, gmax = (-5.0, -5.0, -5.0), (5.0, 5.0, 5.0)
    cases = [([[-45, -46, -47]], 1), ([[-45, -46, -47], [-35, -45, -44]], 2), ([[-48, -46, -47], [-45, -44, -46]], 1)]
    for pts, ncells in cases:
        f = make_file(np.array(pts, np.int32), [0] * len(pts), fmt=0, scale=(0.1,) * 3)
        want = oracle_run([f], ["las"], orc.COLLECT_GRID, bounds=(gmin, gmax), grid=(gmin, gmax, 1.0))
        got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_GRID, bounds=(gmin, gmax), grid=(gmin, gmax, 1.0))
        assert want[0].point_count() == ncells
        assert_same(orc.COLLECT_GRID, got, want)
    p = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_GRID, bounds=(gmin, gmax), grid=(gmin, gmax, 1.0))[0].points()
    assert p["pos"][0].tolist() == [-45 * 0.1, -44 * 0.1, -46 * 0.1]  # the closer-to-centre point wins (test 3)


def test_density_ties_keep_the_first_point(pcq, ctx):
    # many points at exactly the same distance from the cell centre: the strict `<` keeps the first in scan order
    rng = np.random.default_rng(77)
    xyz = np.zeros((50_000, 3), np.int32)
    xyz[:, 0] = rng.integers(0, 8, size=50_000) * 100 + rng.choice([40, 60], size=50_000)  # mirror-symmetric about the centre
    xyz[:, 1] = 50
    xyz[:, 2] = 50
    rgb = rng.integers(0, 65536, size=(50_000, 3)).astype(np.uint16)  # distinguishes otherwise identical points
    f = make_file(xyz, np.arange(50_000) % 7, rgb=rgb, fmt=2, scale=(0.01,) * 3)
    g = ((0.0, 0.0, 0.0), (8.0, 1.0, 1.0), 1.0)
    for variant in (1, 2):
        ctx.set_scan_variant(variant)
        want = oracle_run([f], ["las"], orc.COLLECT_GRID, bounds=(g[0], g[1]), grid=g)
        got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_GRID, bounds=(g[0], g[1]), grid=g)
        assert want[0].point_count() == 8
        assert_same(orc.COLLECT_GRID, got, want)
    ctx.set_scan_variant(0)


def test_density_navvis_shape(pcq, ctx):
    S, B = pcq.synth, pcq.binding
    for fma in (False, True):
        f = S.host_file(S.navvis_spec(n_points=400_000, fma_sensitive=fma))
        for (qmin, qmax) in (S.NAVVIS_S, S.NAVVIS_L, S.NAVVIS_XL):
            grid = (qmin, qmax, S.NAVVIS_DENSITY)
            want = oracle_run([f], ["las"], orc.COLLECT_GRID, bounds=(qmin, qmax), grid=grid)
            got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_GRID, bounds=(qmin, qmax), grid=grid)
            assert_same(orc.COLLECT_GRID, got, want)


def test_density_hashed_table(pcq, ctx):
    rng = np.random.default_rng(55)
    f = random_file(rng, 200_000, 3, "las", 1)
    b = box(1)
    grid = (b[0], b[1], (b[1][0] - b[0][0]) / 300.0)
    want = oracle_run([f], ["las"], orc.COLLECT_GRID, bounds=b, grid=grid)
    os.environ["PCQ_DENSE_MAX_BITS"] = "0"
    os.environ["PCQ_HASH_SLOTS_LOG2"] = "10"  # far too small: forces several rehash rounds
    try:
        got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_GRID, bounds=b, grid=grid)
        assert want[0].point_count() > 10_000
        assert_same(orc.COLLECT_GRID, got, want)
    finally:
        del os.environ["PCQ_DENSE_MAX_BITS"], os.environ["PCQ_HASH_SLOTS_LOG2"]


def test_density_aliasing_kav6_all_orders(pcq, ctx):
    """KAV-6: doc-S grid has 8 z-cells (3 bits); z = 200.00 lies on the inclusive max face -> cell 8 -> masked to 0,
    so A, B, C share key 0 while B is measured against another centre: the reference's result depends on insertion
    order (grid_sampling.rs:62-70 vs 78-82).  The ordered replay must reproduce every order's winner."""
    import itertools
    import json
    from pathlib import Path

    kav = json.loads((Path(__file__).parent / "golden" / "kav.json").read_text())["alias_order"]
    pts = {"A": [1000, 1000, 500], "B": [1100, 1100, 20000], "C": [1200, 1200, 1200]}
    qmin, qmax = pcq.synth.DOC_S
    for order in itertools.permutations("ABC"):
        xyz = np.array([pts[k] for k in order], np.int32)
        f = make_file(xyz, [2, 2, 2], fmt=1, scale=(0.01,) * 3, offset=(390000.0, 130000.0, 0.0))
        want = oracle_run([f], ["las"], orc.COLLECT_GRID, bounds=(qmin, qmax), grid=(qmin, qmax, 25.0))
        got = gpu_run(pcq, ctx, [f], ["las"], orc.COLLECT_GRID, bounds=(qmin, qmax), grid=(qmin, qmax, 25.0))
        assert_same(orc.COLLECT_GRID, got, want)
        winner = kav["winners"]["".join(order)]
        assert got[0].points()["pos"][0].tolist() == [float(v) for v in kav[winner]]


def _alias_files(rng, n_files, n, layout="las", fmt=1):
    """points of a 128 x 128 x 8-cell grid (16 x 16 x 1 at cell 0.125: power-of-two dims on every axis); a tenth of
    them sit exactly on a max face (raw 16384 / 1024 at scale 2^-10), where the cell index exceeds its mask"""
    files = []
    for _ in range(n_files):
        xyz = rng.integers(0, 16_384, size=(n, 3), dtype=np.int32)
        xyz[:, 2] = rng.integers(0, 1025, size=n)
        face = rng.random(n) < 0.1
        axis = rng.integers(0, 3, size=n)
        for a, top in ((0, 16_384), (1, 16_384), (2, 1024)):
            xyz[face & (axis == a), a] = top
        cls = rng.choice(np.array([2, 2, 6], np.uint8), size=n)
        files.append(make_file(xyz, cls, fmt=fmt, scale=(2.0 ** -10,) * 3, offset=(0.0, 0.0, 0.0), layout=layout,
                               seed=int(rng.integers(1 << 30))))
    return files


ALIAS_BOX = ([0.0, 0.0, 0.0], [16.0, 16.0, 1.0])
ALIAS_CELL = 0.125


@pytest.mark.parametrize("layout,fmt", [("las", 1), ("las", 3), ("last", 2)])
def test_density_aliasing_replay_matches_reference_fold(pcq, ctx, layout, fmt):
    rng = np.random.default_rng(7 + fmt)
    files = _alias_files(rng, 3, 20_000, layout, fmt)
    exts = [layout] * 3
    grid = (ALIAS_BOX[0], ALIAS_BOX[1], ALIAS_CELL)
    assert orc.Grid(*grid).dims_bits() == ([128, 128, 8], [7, 7, 3])
    # one collector over all files in one call, one collector per file, and file after file into one collector
    for per_file in (False, True):
        want = oracle_run(files, exts, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid, per_file=per_file)
        got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid, per_file=per_file)
        assert_same(orc.COLLECT_GRID, got, want)
    want = oracle_run(files, exts, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid)
    g = pcq.GridSampledCollector(*grid, ctx=ctx)
    s = pcq.BoundsSearcher(*ALIAS_BOX)
    for f in files:
        s.search_file((f, layout), _impl(pcq), g)
    assert_same(orc.COLLECT_GRID, [g], want)
    # class query: every point of the class is a match, the grid is the query-independent box (main.rs:253-264)
    want = oracle_run(files, exts, orc.COLLECT_GRID, cls=2, grid=grid)
    got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_GRID, cls=2, grid=grid)
    assert_same(orc.COLLECT_GRID, got, want)


def test_density_aliasing_replay_host_streamed_chunks(pcq, ctx):
    """host-staged search: every 1 MB chunk is its own launch, so affected keys appear launch after launch"""
    rng = np.random.default_rng(17)
    files = _alias_files(rng, 2, 150_000)
    grid = (ALIAS_BOX[0], ALIAS_BOX[1], ALIAS_CELL)
    want = oracle_run(files, ["las"] * 2, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid)
    os.environ["PCQ_CHUNK_MB"] = "1"
    try:
        got = gpu_run(pcq, ctx, files, ["las"] * 2, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid, host_stream=True)
    finally:
        del os.environ["PCQ_CHUNK_MB"]
    assert_same(orc.COLLECT_GRID, got, want)


def test_density_aliasing_refuses_cross_gpu_merge(pcq, ctx):
    rng = np.random.default_rng(3)
    files = _alias_files(rng, 1, 5_000)
    grid = (ALIAS_BOX[0], ALIAS_BOX[1], ALIAS_CELL)
    got = gpu_run(pcq, ctx, files, ["las"], orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid)
    with pytest.raises(pcq.PcqError) as e:
        got[0].export_candidates(2)
    assert e.value.code == pcq.binding.PCQ_ERR_ALIASED


def test_density_exchange_two_ranks_emulated(pcq, ctx):
    """Two 'ranks' on one GPU: point ranges of one file are scanned into two local grids, winners are
    exported by owner, swapped, imported into per-owner merge grids — the union must equal one grid."""
    rng = np.random.default_rng(91)
    f = random_file(rng, 300_000, 3, "las", 1)
    b = box(1)
    grid = (b[0], b[1], (b[1][0] - b[0][0]) / 61.0)
    want = oracle_run([f], ["las"], orc.COLLECT_GRID, bounds=b, grid=grid)
    n = 300_000
    cut = 137 * 512
    parts = [(0, cut), (cut, n - cut)]
    local = []
    for first, cnt in parts:
        df = pcq.DeviceFile.stage(ctx, f, "las", first_point=first, n_points=cnt)
        df.set_scan_base(first)
        col = pcq.GridSampledCollector(grid[0], grid[1], grid[2], ctx=ctx)
        pcq.BoundsSearcher(*b).search_files([df], _impl(pcq), [col])
        local.append((df, col))
    exports = [col.export_candidates(2) for _, col in local]
    merged = []
    for owner in range(2):
        m = pcq.GridSampledCollector(grid[0], grid[1], grid[2], ctx=ctx)
        for ptr, counts in exports:
            off = sum(counts[:owner])
            m.import_candidates(ptr + 64 * off if ptr else 0, counts[owner])
        merged.append(m.points())
    allp = np.concatenate(merged)
    assert same_point_set(allp, want[0].points())


# ---- synthetic data: device generator == host generator ------------------------------------------------
def test_device_synth_matches_host_synth(pcq, ctx):
    import torch

    B, S = pcq.binding, pcq.synth
    specs = [S.uniform_spec(10_007, B.LAYOUT_LAS, 0), S.uniform_spec(10_007, B.LAYOUT_LAST, 3), S.doc_specs(4, 9_999)[2],
             S.navvis_spec(12_345, fma_sensitive=True), S.ca13_specs(4, 8_191)[1], S.uniform_spec(5000, B.LAYOUT_LAS, 1, record_len=31)]
    for sp in specs:
        host = S.host_file(sp)
        buf = torch.empty(host.nbytes - 227, dtype=torch.uint8, device=f"cuda:{ctx.device}")
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        assert np.array_equal(buf.cpu().numpy(), host[227:])
        assert np.array_equal(S.header_bytes(sp, mm), host[:227])
        # and a wrapped device file answers like the staged host file
        layout = S.ext_of_layout(sp.layout)
        df = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
        for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER):
            col = (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)
            pcq.ClassSearcher(2).search_files([df], _impl(pcq), [col])
            assert_same(kind, [col], oracle_run([host], [layout], kind, cls=2))


def test_wrapped_unaligned_device_memory(pcq, ctx):
    """Point data at an odd device address (a whole file image copied verbatim: records start at +227)."""
    import torch

    rng = np.random.default_rng(61)
    for layout, fmt in (("las", 1), ("las", 2), ("last", 3)):
        f = random_file(rng, 20_011, fmt, layout, 2)
        t = torch.from_numpy(f).to(f"cuda:{ctx.device}")
        d = pcq.FileDesc()
        pcq.binding.check(pcq.lib.pcq_parse_header(C.c_void_p(f.ctypes.data), f.nbytes, 0 if layout == "las" else 1, 0, C.byref(d)))
        df = pcq.DeviceFile.wrap(ctx, d, t.data_ptr() + d.point_data_off, keepalive=t)
        b = box(2)
        for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER):
            col = (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)
            pcq.BoundsSearcher(*b).search_files([df], _impl(pcq), [col])
            assert_same(kind, [col], oracle_run([f], [layout], kind, bounds=b))
            col = (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)
            pcq.ClassSearcher(6).search_files([df], _impl(pcq), [col])
            assert_same(kind, [col], oracle_run([f], [layout], kind, cls=6))


# ---- golden fixtures + size-independent properties at benchmark-like sizes ------------------------------
def test_golden_fixtures(pcq, ctx):
    import hashlib
    import json
    from pathlib import Path

    from tests.golden.make_golden import CASES, build_case

    golden = json.loads((Path(__file__).parent / "golden" / "scan_golden.json").read_text())
    for name in CASES:
        files, exts, kw = build_case(pcq, name)
        g = golden[name]
        for kind, key in ((orc.COLLECT_COUNT, "count"), (orc.COLLECT_BUFFER, "buffer_sha256")):
            cols = gpu_run(pcq, ctx, files, exts, kind, per_file=True, **kw)
            if kind == orc.COLLECT_COUNT:
                assert [c.point_count() for c in cols] == g["counts"]
            else:
                h = hashlib.sha256()
                for c in cols:
                    h.update(np.ascontiguousarray(c.points()).view(np.uint8).tobytes())
                assert h.hexdigest() == g["buffer_sha256"]


def test_large_device_resident_properties(pcq, ctx):
    """C2-shaped tiles at 4 M points each, generated on the device: properties that need no oracle."""
    import torch

    S, B = pcq.synth, pcq.binding
    specs = S.doc_specs(n_files=4, pts_per_file=4_000_000)
    dfs, descs = [], []
    for sp in specs:
        buf = torch.empty(sp.n_points * sp.record_len + 256, dtype=torch.uint8, device=f"cuda:{ctx.device}")
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        dfs.append(pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf))
    total = sum(sp.n_points for sp in specs)
    xl = pcq.BoundsSearcher(*S.DOC_XL)
    cols = pcq.run_search_parallel(dfs, xl, _impl(pcq), lambda: pcq.CountCollector(ctx))
    assert [c.point_count() for c in cols] == [sp.n_points for sp in specs]  # XL covers every tile entirely
    # class counts over a partition of the class byte sum to N (checksum of checksums)
    per_class = 0
    for klass in range(256):
        c = pcq.run_search_sequential(dfs, pcq.ClassSearcher(klass), _impl(pcq), lambda: pcq.CountCollector(ctx))
        per_class += c.point_count()
    assert per_class == total
    # a box split in two along x: counts add up; select returns exactly count records, in order, idempotently
    (x0, y0, z0), (x1, y1, z1) = S.DOC_XL
    xm = 395123.45
    whole = pcq.run_search_sequential(dfs, pcq.BoundsSearcher((x0, y0, 0.0), (x1, y1, 100.0)), _impl(pcq), lambda: pcq.CountCollector(ctx)).point_count()
    left = pcq.run_search_sequential(dfs, pcq.BoundsSearcher((x0, y0, 0.0), (xm, y1, 100.0)), _impl(pcq), lambda: pcq.CountCollector(ctx)).point_count()
    right_col = pcq.run_search_sequential(dfs, pcq.BoundsSearcher((xm + 0.01, y0, 0.0), (x1, y1, 100.0)), _impl(pcq), lambda: pcq.BufferCollector(ctx))
    assert 0 < left < whole and left + right_col.point_count() == whole
    pts = right_col.points()
    assert len(pts) == right_col.point_count()
    assert pts["pos"][:, 0].min() >= xm and pts["pos"][:, 2].max() <= 100.0 + 1e-9
    again = pcq.run_search_sequential(dfs, pcq.BoundsSearcher((xm + 0.01, y0, 0.0), (x1, y1, 100.0)), _impl(pcq), lambda: pcq.BufferCollector(ctx))
    assert same_point_seq(pts, again.points())
    for variant in (1, 2):  # both kernel variants return the same stream
        ctx.set_scan_variant(variant)
        v = pcq.run_search_sequential(dfs, pcq.BoundsSearcher((xm + 0.01, y0, 0.0), (x1, y1, 100.0)), _impl(pcq), lambda: pcq.BufferCollector(ctx))
        assert same_point_seq(pts, v.points())
    ctx.set_scan_variant(0)
    # density: one cell can hold at most one point, every selected point is a matching point
    g = pcq.GridSampledCollector(S.DOC_XL[0], S.DOC_XL[1], S.DOC_DENSITY, ctx=ctx)
    xl.search_files(dfs, _impl(pcq), [g])
    gp = g.points()
    assert len(gp) == g.point_count() and len(gp) > 1000
    dims, bits = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
    B.check(B.lib.pcq_grid_params(B.d3(S.DOC_XL[0]), B.d3(S.DOC_XL[1]), S.DOC_DENSITY, dims, bits))
    gmin, gmax = np.array(S.DOC_XL[0]), np.array(S.DOC_XL[1])
    cell = ((gp["pos"] - gmin) * np.array(list(dims), dtype=np.float64) / (gmax - gmin)).astype(np.int64)  # grid_sampling.rs:51-60
    assert len(np.unique(cell, axis=0)) == len(gp)


def test_more_than_2_32_points_in_one_file(pcq):
    """maximum sizes: one resident file of 4.5 G points (90 GB, more than a u32 can index), checked through
    size-independent properties (tools/big_check.py) in its own process"""
    import subprocess
    import sys

    import torch

    free, _ = torch.cuda.mem_get_info(0)
    if free < 120 * (1 << 30):
        pytest.skip("needs 120 GB of free HBM")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "big_check.py"), "--points", "4500000000", "--layout", "las"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


def test_density_exchange_over_nccl_when_two_gpus(pcq):
    """Real multi-rank run (one process per GPU, NCCL all-to-all) — needs >= 2 GPUs on the box."""
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (covered on CPU by tests/test_sharding_gloo.py)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(root, "tests", "dist_density_nccl.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout and r.stdout.count("OK") == 4


# ---- -o output prepared on the device (pcq_collector_las_records <- dump_points.rs:63-116) -----------------------------
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 70_001])
def test_las_records_on_device_equal_the_restated_file_dumper(pcq, ctx, n):
    """The device-side min / max reduction + quantisation writes what the oracle's restatement of FileDumper writes from
    the oracle's points: offset = min position, scale, raw coordinates, class and colour of every record, in scan order."""
    from oracle import np_oracle

    rng = np.random.default_rng(700 + n)
    xyz = rng.integers(-2_000_000, 2_000_000, size=(n, 3), dtype=np.int64).astype(np.int32)
    cls = rng.integers(0, 4, size=n).astype(np.uint8)
    rgb = rng.integers(0, 65536, size=(n, 3)).astype(np.uint16)
    f = make_file(xyz, cls, rgb=rgb, fmt=3, scale=(0.001, 0.001, 0.001), offset=(-23.108, -21.261, -10.029))
    box = ((-1e9,) * 3, (1e9,) * 3)
    for kind, kw in ((orc.COLLECT_BUFFER, dict(bounds=box)), (orc.COLLECT_BUFFER, dict(cls=1)),
                     (orc.COLLECT_GRID, dict(bounds=box, grid=((-2100.0,) * 3, (2100.0,) * 3, 50.0)))):
        got = gpu_run(pcq, ctx, [f], ["las"], kind, **kw)[0]
        want = oracle_run([f], ["las"], kind, **kw)[0]
        res = got.las_records()
        if want.point_count() == 0:
            assert res is None
            continue
        pts = got.points()  # GRID: the device's own order (HashMap order in the reference); the set is checked elsewhere
        if kind == orc.COLLECT_BUFFER:
            assert pts.tobytes() == want.points().tobytes()
        (e,) = np_oracle.dump_points_plan([pts])
        mn, mx, scale, rec = res
        assert np.array_equal(mn, e["offset"]) and scale == e["scale"] and np.array_equal(mx, pts["pos"].max(axis=0))
        assert rec.shape == (len(pts), 26)
        assert np.array_equal(np.ascontiguousarray(rec[:, :12]).view("<i4").reshape(-1, 3), e["raw"])
        assert np.array_equal(rec[:, 15], e["cls"]) and (rec[:, 14] == 0x09).all()
        assert np.array_equal(np.ascontiguousarray(rec[:, 20:26]).view("<u2").reshape(-1, 3), e["rgb"])
        assert not rec[:, [12, 13, 16, 17, 18, 19]].any()


def test_batched_collector_reset(pcq, ctx):
    """pcq_collectors_reset clears count, buffer and grid collectors like pcq_collector_reset does one by one."""
    rng = np.random.default_rng(41)
    files = [make_file(rng.integers(0, 100_000, size=(5000 + 100 * k, 3)), rng.integers(1, 4, size=5000 + 100 * k), fmt=1, seed=k)
             for k in range(5)]
    exts = ["las"] * len(files)
    box1, box2 = ((0.0, 0.0, 0.0), (600.0, 900.0, 1000.0)), ((100.0, 0.0, 0.0), (1000.0, 300.0, 500.0))
    grid = ((0.0, 0.0, 0.0), (1000.0, 1000.0, 1000.0), 50.0)
    images = list(zip(files, exts))
    for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER, orc.COLLECT_GRID):
        def new():
            if kind == orc.COLLECT_GRID:
                return pcq.GridSampledCollector(*grid, ctx=ctx)
            return (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)

        cols = [new() for _ in files]
        pcq.BoundsSearcher(*box1).search_files(images, _impl(pcq), cols)
        assert_same(kind, cols, oracle_run(files, exts, kind, bounds=box1, grid=grid, per_file=True))
        pcq.reset_collectors(cols)
        assert [c.point_count() for c in cols] == [0] * len(cols)
        pcq.BoundsSearcher(*box2).search_files(images, _impl(pcq), cols)
        assert_same(kind, cols, oracle_run(files, exts, kind, bounds=box2, grid=grid, per_file=True))


def test_many_per_file_grids_share_the_hbm(pcq, ctx):
    """`query --parallel --density` makes one grid per file (main.rs:253-273).  A 2^30-cell grid is an 8 GB dense table:
    twelve of them must not be allocated side by side — the tables are sized when the collectors are first used, by how
    many of them the call brings (ensure_grid_tables), and the result does not depend on dense or hashed."""
    rng = np.random.default_rng(77)
    n_files = 12
    files = [make_file(rng.integers(0, 1_000_000, size=(3000, 3)), rng.integers(1, 4, size=3000), fmt=1, seed=k) for k in range(n_files)]
    exts = ["las"] * n_files
    box = ((0.0, 0.0, 0.0), (10_000.0, 10_000.0, 10_000.0))
    grid = (box[0], box[1], 10.0)  # 1000^3 cells: 10 + 10 + 10 key bits
    got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_GRID, bounds=box, grid=grid, per_file=True)
    want = oracle_run(files, exts, orc.COLLECT_GRID, bounds=box, grid=grid, per_file=True)
    assert_same(orc.COLLECT_GRID, got, want)


def test_per_file_grid_tables_cover_the_file_box_and_fall_back(pcq, ctx):
    """A per-file grid's table covers only the cells under the file's header box (cut by the query box).  A header that
    understates its bounds sends points outside that sub-box: the collector moves to a table over the whole grid and the
    launch runs again — the result is the oracle's either way, resident and host-streamed, and after a reset."""
    rng = np.random.default_rng(91)
    files, exts = [], []
    for k in range(6):
        lo = np.array([k * 150_000, 0, 0])
        xyz = lo + rng.integers(0, 150_000, size=(4000, 3))
        honest = dict()
        if k % 2 == 1:  # lying header: claims a tenth of the real extent (still intersects the query)
            honest = dict(hdr_min=(lo[0] * 0.01, 0.0, 0.0), hdr_max=(lo[0] * 0.01 + 150.0, 150.0, 150.0))
        files.append(make_file(xyz, rng.integers(1, 4, size=4000), fmt=3 if k % 3 == 0 else 1, seed=k, **honest))
        exts.append("las")
    box = ((0.0, 0.0, 0.0), (9000.0, 1500.0, 1500.0))
    grid = (box[0], box[1], 5.0)  # 1800 x 300 x 300 cells (11 + 9 + 9 key bits); a file's honest box is 300^3 of them
    want = oracle_run(files, exts, orc.COLLECT_GRID, bounds=box, grid=grid, per_file=True)
    for host_stream in (False, True):
        got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_GRID, bounds=box, grid=grid, per_file=True, host_stream=host_stream)
        assert_same(orc.COLLECT_GRID, got, want)
        # the same collectors again after a reset, fed the files in reverse: other boxes than the tables were made for
        pcq.reset_collectors(got)
        searcher = pcq.BoundsSearcher(*box)
        searcher.search_files(list(zip(files[::-1], exts)), _impl(pcq), got)
        assert_same(orc.COLLECT_GRID, got, want[::-1])
    # one grid over all files (run_search_sequential): the box is the union of the header boxes
    got = gpu_run(pcq, ctx, files, exts, orc.COLLECT_GRID, bounds=box, grid=grid)
    assert_same(orc.COLLECT_GRID, got, oracle_run(files, exts, orc.COLLECT_GRID, bounds=box, grid=grid))
