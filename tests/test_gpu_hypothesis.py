"""Property-driven device tests: random formats x headers (incl. FMA-sensitive ones) x query boxes
(incl. boxes that saturate i64, infinite and NaN bounds, boxes of zero extent) x density grids (incl. an axis of zero
extent, whose every point falls into cell 0) — the CUDA path through the C ABI must equal the oracle every time, and
must fail where it fails (AABB::from_min_max panics on min > max)."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle as orc
from tests.helpers import make_file, same_point_seq, same_point_set
from tests.test_gpu_parity import HEADERS

pytestmark = pytest.mark.gpu

FORMATS = [0, 1, 2, 3, 6, 7]
SPECIAL = [0.0, -0.0, 1e300, -1e300, float("inf"), float("-inf"), float("nan"), 9.3e18, -9.3e18, 2147483647.0, -2147483648.0]


@st.composite
def boxes(draw, hdr):
    scale, offset = HEADERS[hdr]
    lo, hi = [], []
    for a in range(3):
        def coord():
            kind = draw(st.integers(0, 9))
            if kind == 0:
                return draw(st.sampled_from(SPECIAL))
            if kind == 1:  # exactly on a raw lattice value: `<=` on both sides matters
                return offset[a] + draw(st.integers(-60_000, 160_000)) * scale[a]
            return offset[a] + draw(st.floats(-700.0, 1700.0)) * (scale[a] * 100.0)
        x, y = coord(), coord()
        if draw(st.integers(0, 19)) != 0 and not (x != x or y != y) and x > y:
            x, y = y, x  # mostly well-formed; one in twenty keeps min > max (a panic in the reference)
        lo.append(x)
        hi.append(y)
    return lo, hi


@st.composite
def cases(draw):
    fmt = draw(st.sampled_from(FORMATS))
    layout = draw(st.sampled_from(["las", "last"]))
    hdr = draw(st.integers(0, len(HEADERS) - 1))
    n = draw(st.sampled_from([0, 1, 33, 600, 2500]))
    seed = draw(st.integers(0, 2**31 - 1))
    box = draw(boxes(hdr))
    extra = draw(st.sampled_from([0, 0, 3]))
    return fmt, layout, hdr, n, seed, box, extra


def build(fmt, layout, hdr, n, seed, extra):
    from tests.helpers import FORMAT_LEN

    rng = np.random.default_rng(seed)
    scale, offset = HEADERS[hdr]
    xyz = rng.integers(-50_000, 150_000, size=(n, 3), dtype=np.int32)
    cls = rng.choice(np.array([1, 2, 2, 2, 5, 6, 6 | 0x20, 2 | 0x80, 9], np.uint8), size=n)
    return make_file(xyz, cls, fmt=fmt, scale=scale, offset=offset, layout=layout, record_len=FORMAT_LEN[fmt] + extra, seed=seed,
                     version=(1, 4) if fmt >= 6 else (1, 2))


def both(pcq, ctx, f, ext, kind, bounds=None, cls=None, grid=None):
    """-> ('ok', gpu collector, oracle collector) or ('err', gpu code, oracle code)"""
    try:
        oc = orc.Collector(kind, *(grid or ()))
        orc.search_file(f, ext, oc, bounds=bounds, cls=cls)
        oerr = None
    except orc.OracleError as e:
        oerr = e
    try:
        if kind == orc.COLLECT_GRID:
            gc = pcq.GridSampledCollector(grid[0], grid[1], grid[2], ctx=ctx)
        else:
            gc = (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)
        s = pcq.BoundsSearcher(*bounds) if bounds is not None else pcq.ClassSearcher(cls)
        s.search_files([(f, ext)], pcq.SearchImplementation.Optimized, [gc])
        gerr = None
    except pcq.PcqError as e:
        gerr = e
    assert (oerr is None) == (gerr is None), f"oracle: {oerr!r}, device: {gerr!r}"
    if oerr is not None:
        return "err", gerr, oerr
    return "ok", gc, oc


@settings(max_examples=120, deadline=None, derandomize=True)
@given(cases())
def test_bounds_queries_equal_the_oracle(pcq, ctx, case):
    fmt, layout, hdr, n, seed, box, extra = case
    f = build(fmt, layout, hdr, n, seed, extra)
    if any(box[0][a] > box[1][a] for a in range(3)):
        # such a BoundsSearcher cannot exist in the reference: AABB::from_min_max panics when main.rs:80 builds the box
        # (NaN compares false and passes, as in Rust); the oracle's search functions take the box as given
        with pytest.raises(pcq.PcqError) as e:
            pcq.BoundsSearcher(*box).search_files([(f, layout)], pcq.SearchImplementation.Optimized, [pcq.CountCollector(ctx)])
        assert e.value.code == pcq.binding.PCQ_ERR_PANIC
        return
    for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER):
        r = both(pcq, ctx, f, layout, kind, bounds=box)
        if r[0] == "err":
            assert r[1].code == pcq.binding.PCQ_ERR_PANIC or r[1].code == pcq.binding.PCQ_ERR_FORMAT
            continue
        _, g, o = r
        assert g.point_count() == o.point_count()
        if kind == orc.COLLECT_BUFFER:
            assert same_point_seq(g.points(), o.points())


@settings(max_examples=80, deadline=None, derandomize=True)
@given(cases(), st.integers(0, 255), st.integers(0, 3), st.sampled_from([0.37, 1.0, 25.0, 1e-3]))
def test_class_and_density_queries_equal_the_oracle(pcq, ctx, case, klass, flat_axis, cell):
    fmt, layout, hdr, n, seed, box, extra = case
    f = build(fmt, layout, hdr, n, seed, extra)
    klass = klass if klass % 3 == 0 else [1, 2, 6, 6 | 0x20][klass % 4]
    r = both(pcq, ctx, f, layout, orc.COLLECT_BUFFER, cls=klass)
    if r[0] == "ok":
        assert same_point_seq(r[1].points(), r[2].points())
    # a density grid over a finite box; `flat_axis` < 3 gives that axis zero extent (every point -> cell 0 on it)
    scale, offset = HEADERS[hdr]
    gmin = [offset[a] + 100 * scale[a] for a in range(3)]
    gmax = [offset[a] + 90_000 * scale[a] for a in range(3)]
    if flat_axis < 3:
        gmax[flat_axis] = gmin[flat_axis]
    cell = cell * scale[0] * 1000.0
    try:
        orc.Grid(gmin, gmax, cell)
    except orc.OracleError:
        with pytest.raises(pcq.PcqError):
            pcq.GridSampledCollector(gmin, gmax, cell, ctx=ctx)
        return
    r = both(pcq, ctx, f, layout, orc.COLLECT_GRID, bounds=(gmin, gmax), grid=(gmin, gmax, cell))
    if r[0] == "ok":
        assert r[1].point_count() == r[2].point_count()
        assert same_point_set(r[1].points(), r[2].points())
    r = both(pcq, ctx, f, layout, orc.COLLECT_GRID, cls=2, grid=(gmin, gmax, cell))
    if r[0] == "ok":
        assert same_point_set(r[1].points(), r[2].points())
