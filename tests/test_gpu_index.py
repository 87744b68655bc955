"""GPU tests of the on-the-fly chunk index (improvements.md:3-10; include/pcq.h "chunk index").

The reference only describes the idea, so there is no reference output to pin: the bar is (1) the headers equal the
numpy restatement (oracle/np_oracle.py chunk_headers) bit for bit, and (2) a search over an indexed file returns
exactly what the full scan and the CPU oracle return — counts, scan-ordered record streams, density winners — while
the statistics prove that chunks really were skipped."""
import ctypes as C

import numpy as np
import pytest

from oracle import np_oracle as npo
from oracle import oracle as orc
from tests.helpers import make_file, same_point_set
from tests.test_gpu_parity import HEADERS, assert_same, oracle_run

pytestmark = pytest.mark.gpu

CH = 8192


def strip_arrays(rng, n, rare_every=5):
    """Acquisition-order-like points: x grows along the file (with some jitter), y/z random; class 7 occurs only in
    every `rare_every`-th group of five chunks, classes 9 and 11 nowhere."""
    i = np.arange(n)
    x = (i * 40 + rng.integers(-3000, 3000, size=n)).astype(np.int32)
    y = rng.integers(-50_000, 150_000, size=n, dtype=np.int32)
    z = rng.integers(-500, 5000, size=n, dtype=np.int32)
    cls = rng.choice(np.array([1, 2, 2, 5, 6, 2 | 0x80], np.uint8), size=n)
    rare = ((i // (5 * CH)) % rare_every == 0) & (rng.random(n) < 0.01)
    cls[rare] = 7
    return np.stack([x, y, z], axis=1), cls


def strip_file(rng, n, fmt, layout, hdr=0, record_len=None, rare_every=5):
    scale, offset = HEADERS[hdr]
    xyz, cls = strip_arrays(rng, n, rare_every)
    return make_file(xyz, cls, fmt=fmt, scale=scale, offset=offset, layout=layout,
                     record_len=record_len, seed=int(rng.integers(1 << 30)), version=(1, 4) if fmt >= 6 else (1, 2))


def xbox(hdr, x0, x1):
    """query box selecting raw x in about [x0, x1], everything in y and z"""
    scale, offset = HEADERS[hdr]
    lo = [offset[0] + x0 * scale[0], offset[1] - 1e6 * scale[0], offset[2] - 1e6 * scale[0]]
    hi = [offset[0] + x1 * scale[0], offset[1] + 1e6 * scale[1], offset[2] + 1e6 * scale[2]]
    return lo, hi


def new_collector(pcq, ctx, kind, grid=None):
    if kind == orc.COLLECT_GRID:
        return pcq.GridSampledCollector(grid[0], grid[1], grid[2], ctx=ctx)
    return (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)


def search(pcq, ctx, dfs, kind, bounds=None, cls=None, grid=None, per_file=False):
    s = pcq.BoundsSearcher(*bounds) if bounds is not None else pcq.ClassSearcher(cls)
    cols = [new_collector(pcq, ctx, kind, grid) for _ in dfs] if per_file else [new_collector(pcq, ctx, kind, grid)]
    s.search_files(dfs, pcq.SearchImplementation.Optimized, cols)
    return cols


@pytest.mark.parametrize("layout,fmt,record_len", [("las", 0, None), ("las", 1, None), ("las", 3, 37), ("las", 6, None),
                                                   ("las", 2, 29), ("last", 1, None), ("last", 3, None), ("last", 6, None)])
def test_chunk_headers_equal_the_restatement(pcq, ctx, layout, fmt, record_len):
    rng = np.random.default_rng(7 + fmt)
    for n in (1, CH - 1, CH, 3 * CH + 17):
        f = strip_file(rng, n, fmt, layout, fmt % 2, record_len)
        df = pcq.DeviceFile.stage(ctx, f, layout)
        assert df.index.shape[0] == 0
        df.build_index()
        df.build_index()  # idempotent
        got, want, want_c = df.index, npo.chunk_headers(f, layout), orc.chunk_headers(f, layout)
        assert got.shape == want.shape == want_c.shape == ((n + CH - 1) // CH,)
        for k in ("lo", "hi", "cls_bits", "n_points"):
            assert np.array_equal(got[k], want[k]) and np.array_equal(got[k], want_c[k]), k
        df.drop_index()
        assert df.index.shape[0] == 0
        df.release()


def test_chunk_headers_of_a_point_range_and_unaligned_memory(pcq, ctx):
    import torch

    rng = np.random.default_rng(11)
    for layout in ("las", "last"):
        f = strip_file(rng, 5 * CH + 100, 1, layout, 0)
        df = pcq.DeviceFile.stage(ctx, f, layout, first_point=CH + 3, n_points=3 * CH)
        df.build_index()
        want = npo.chunk_headers(f, layout, first=CH + 3, count=3 * CH)
        got = df.index
        for k in ("lo", "hi", "cls_bits", "n_points"):
            assert np.array_equal(got[k], want[k]), (layout, k)
        df.release()
        # whole file image copied verbatim: point data starts at the odd address +227
        t = torch.from_numpy(f).to(f"cuda:{ctx.device}")
        d = pcq.FileDesc()
        pcq.binding.check(pcq.lib.pcq_parse_header(C.c_void_p(f.ctypes.data), f.nbytes, 0 if layout == "las" else 1, 1, C.byref(d)))
        dw = pcq.DeviceFile.wrap(ctx, d, t.data_ptr() + d.point_data_off, keepalive=t)
        dw.build_index()
        want = npo.chunk_headers(f, layout)
        for k in ("lo", "hi", "cls_bits", "n_points"):
            assert np.array_equal(dw.index[k], want[k]), (layout, k)
        col = pcq.BufferCollector(ctx)
        b = xbox(0, 2 * CH * 40, 3 * CH * 40)
        pcq.BoundsSearcher(*b).search_files([dw], pcq.SearchImplementation.Optimized, [col])
        assert ctx.last_scan_stats.chunks_skipped > 0
        assert_same(orc.COLLECT_BUFFER, [col], oracle_run([f], [layout], orc.COLLECT_BUFFER, bounds=b))
        dw.release()


@pytest.mark.parametrize("layout,fmt", [("las", 1), ("las", 2), ("las", 3), ("last", 1), ("last", 3)])
def test_indexed_search_equals_full_scan_and_oracle(pcq, ctx, layout, fmt):
    rng = np.random.default_rng(40 + fmt + (layout == "last"))
    hdr = fmt % 2
    n = 61 * CH + 1234
    files = [strip_file(rng, n, fmt, layout, hdr), strip_file(rng, 7 * CH, fmt, layout, hdr)]
    exts = [layout, layout]
    plain = [pcq.DeviceFile.stage(ctx, f, layout) for f in files]
    indexed = [pcq.DeviceFile.stage(ctx, f, layout) for f in files]
    for d in indexed:
        d.build_index()
    span = n * 40
    queries = [dict(bounds=xbox(hdr, int(0.40 * span), int(0.45 * span))),   # one narrow run of chunks
               dict(bounds=xbox(hdr, -10_000, 2 * span)),                     # everything
               dict(bounds=xbox(hdr, int(0.02 * span), int(0.97 * span))),    # > 90 % survive: scanned whole
               dict(bounds=xbox(hdr, 5 * span, 6 * span)),                    # beyond the data but inside no header: none
               dict(cls=7), dict(cls=9), dict(cls=2), dict(cls=2 | 0x80)]
    scale, offset = HEADERS[hdr]
    gmin = [offset[0] - 4000 * scale[0], offset[1] - 60_000 * scale[0], offset[2] - 600 * scale[0]]
    gmax = [offset[0] + (span + 4000) * scale[0], offset[1] + 160_000 * scale[1], offset[2] + 6000 * scale[2]]
    grid = (gmin, gmax, (gmax[0] - gmin[0]) / 300.0)
    for q in queries:
        for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER, orc.COLLECT_GRID):
            for per_file in (False, True):
                want = oracle_run(files, exts, kind, grid=grid, per_file=per_file, **q)
                full = search(pcq, ctx, plain, kind, grid=grid, per_file=per_file, **q)
                full_stats = ctx.last_scan_stats
                got = search(pcq, ctx, indexed, kind, grid=grid, per_file=per_file, **q)
                st = ctx.last_scan_stats
                assert_same(kind, full, want)
                assert_same(kind, got, want)
                if kind == orc.COLLECT_GRID:  # same winners AND the same tie-breaks as the full scan on this device
                    for a, b in zip(got, full):
                        assert same_point_set(a.points(), b.points())
                assert full_stats.chunks_total == 0 and full_stats.points_scanned == full_stats.points_total
                if st.points_total:
                    assert st.chunks_total > 0
        # what the statistics must show for the queries whose answer is known by construction
        if q.get("cls") == 9:
            assert st.chunks_skipped == st.chunks_total and st.points_scanned == 0 and st.segments == 0
        if q.get("cls") == 7:
            assert 0 < st.chunks_skipped < st.chunks_total and st.segments >= 3  # several runs of the big file
        if q.get("cls") == 2:
            assert st.chunks_skipped == 0
    # the narrow box again, for the numbers
    b = queries[0]["bounds"]
    c = search(pcq, ctx, indexed, orc.COLLECT_COUNT, bounds=b)
    st = ctx.last_scan_stats
    assert st.chunks_skipped > 0.8 * st.chunks_total and st.points_scanned < 0.2 * st.points_total
    assert c[0].point_count() == oracle_run(files, exts, orc.COLLECT_COUNT, bounds=b)[0].point_count()
    for d in plain + indexed:
        d.release()


def test_index_with_many_short_runs(pcq, ctx):
    """Chunks that can match alternate with chunks that cannot, further apart than the join gap: the launch holds
    hundreds of short point ranges (segment look-up gallops, kernels.cu seg_forward)."""
    rng = np.random.default_rng(77)
    n = 599 * CH + 99
    xyz, cls = strip_arrays(rng, n)
    marks = np.arange(n)[(np.arange(n) // CH) % 6 == 0]  # class 11 in one chunk out of six: chunks 0, 6, ..., 594
    cls[marks[rng.random(marks.shape[0]) < 0.02]] = 11
    for layout in ("las", "last"):
        f = make_file(xyz, cls, fmt=1, scale=HEADERS[0][0], offset=HEADERS[0][1], layout=layout, seed=3)
        df = pcq.DeviceFile.stage(ctx, f, layout)
        df.build_index()
        for kind in (orc.COLLECT_COUNT, orc.COLLECT_BUFFER):
            got = search(pcq, ctx, [df], kind, cls=11)
            st = ctx.last_scan_stats
            if layout == "last" and kind == orc.COLLECT_COUNT:
                # counting a class in a 1-byte column is cheaper whole than fragmented unless > 7/8 can be skipped
                assert st.segments == 1 and st.chunks_total == 600 and st.chunks_skipped == 0
            else:
                assert st.segments == 100 and st.chunks_total == 600 and st.chunks_skipped == 500
            assert_same(kind, got, oracle_run([f], [layout], kind, cls=11))
        df.release()


def test_auto_index_builds_on_the_second_scan(pcq):
    ctx = pcq.Context(0)
    try:
        ctx.set_auto_index(1)
        rng = np.random.default_rng(5)
        f = strip_file(rng, 20 * CH, 1, "las", 0)
        df = pcq.DeviceFile.stage(ctx, f, "las")
        b = xbox(0, 3 * CH * 40, 5 * CH * 40)
        want = oracle_run([f], ["las"], orc.COLLECT_BUFFER, bounds=b)
        first = search(pcq, ctx, [df], orc.COLLECT_BUFFER, bounds=b)
        assert df.index.shape[0] == 0 and ctx.last_scan_stats.chunks_total == 0
        second = search(pcq, ctx, [df], orc.COLLECT_BUFFER, bounds=b)
        assert df.index.shape[0] == 20 and ctx.last_scan_stats.chunks_skipped >= 14
        assert_same(orc.COLLECT_BUFFER, first, want)
        assert_same(orc.COLLECT_BUFFER, second, want)
        df.release()
    finally:
        ctx.close()


def test_indexed_density_keeps_scan_order_ties(pcq, ctx):
    """Two points at the same distance from a cell centre in different chunks: the earlier one must win with and
    without the index (strict `<`, grid_sampling.rs:84-102), including when the chunks between them are skipped."""
    n = 12 * CH
    xyz = np.zeros((n, 3), dtype=np.int32)
    xyz[:, 0] = 1_000_000 + np.arange(n)  # far outside the query box
    cls = np.full(n, 1, np.uint8)
    # cell [0,100)^3 raw, centre 50: equidistant points 40 and 60 on x, in chunk 1 and chunk 9; closer one in chunk 10 of cell 2
    for i, x in ((CH + 5, 40), (9 * CH + 7, 60), (10 * CH + 1, 250), (2 * CH + 9, 230)):
        xyz[i] = (x, 50, 50)
    rgb = np.zeros((n, 3), np.uint16)
    rgb[:, 0] = np.arange(n) % 65536
    f = make_file(xyz, cls, rgb=rgb, fmt=2, scale=(1.0, 1.0, 1.0), offset=(0.0, 0.0, 0.0), layout="las")
    b = ([0.0, 0.0, 0.0], [399.0, 99.0, 99.0])
    grid = ([0.0, 0.0, 0.0], [400.0, 100.0, 100.0], 100.0)
    want = oracle_run([f], ["las"], orc.COLLECT_GRID, bounds=b, grid=grid)
    df = pcq.DeviceFile.stage(ctx, f, "las")
    df.build_index()
    got = search(pcq, ctx, [df], orc.COLLECT_GRID, bounds=b, grid=grid)
    st = ctx.last_scan_stats
    assert st.chunks_skipped >= 6
    assert_same(orc.COLLECT_GRID, got, want)
    pts = got[0].points()
    assert sorted(pts["pos"][:, 0].tolist()) == [40.0, 250.0]
    df.release()


# ---- host-staged passes with a host index (pcq_search_host_files_indexed) ---------------------------------------
def host_pass(pcq, ctx, images, queries, kind, index, grid=None, per_file=False):
    """one batch of queries over host file images -> list (per query) of collector lists"""
    searchers = [pcq.BoundsSearcher(*q["bounds"]) if "bounds" in q else pcq.ClassSearcher(q["cls"]) for q in queries]
    cols = [[new_collector(pcq, ctx, kind, grid) for _ in (images if per_file else [0])] for _ in queries]
    pcq.search_host_files_multi(images, searchers, cols, index=index)
    return cols


@pytest.mark.parametrize("layout,fmt", [("las", 1), ("las", 3), ("last", 1), ("last", 3)])
def test_host_index_is_a_by_product_and_later_passes_copy_less(pcq, ctx, layout, fmt, monkeypatch):
    monkeypatch.setenv("PCQ_CHUNK_MB", "2")  # several pieces per file, several runs per piece
    rng = np.random.default_rng(90 + fmt + (layout == "last"))
    hdr = fmt % 2
    n = 61 * CH + 1234
    files = [strip_file(rng, n, fmt, layout, hdr), strip_file(rng, 7 * CH, fmt, layout, hdr)]
    images = [(f, layout) for f in files]
    exts = [layout, layout]
    span = n * 40
    narrow = dict(bounds=xbox(hdr, int(0.40 * span), int(0.45 * span)))
    two = dict(bounds=xbox(hdr, int(0.80 * span), int(0.82 * span)))
    scale, offset = HEADERS[hdr]
    gmin = [offset[0] - 4000 * scale[0], offset[1] - 60_000 * scale[0], offset[2] - 600 * scale[0]]
    gmax = [offset[0] + (span + 4000) * scale[0], offset[1] + 160_000 * scale[1], offset[2] + 6000 * scale[2]]
    grid = (gmin, gmax, (gmax[0] - gmin[0]) / 300.0)
    ix = pcq.HostIndex(ctx)
    assert ix.info(0) == (0, False, False)

    def check(queries, kind, per_file=False):
        got = host_pass(pcq, ctx, images, queries, kind, ix, grid=grid, per_file=per_file)
        st = ctx.last_scan_stats
        for q, cols in zip(queries, got):
            assert_same(kind, cols, oracle_run(files, exts, kind, grid=grid, per_file=per_file, **q))
        return st

    # pass 1, a bounds count: everything crosses PCIe; LAS records carry both attributes, LAST only the positions
    st = check([narrow], orc.COLLECT_COUNT)
    assert st.chunks_total == 0 and st.points_scanned == st.points_total == n  # (the small file fails the header test)
    assert ix.info(0) == (62, True, layout == "las")
    # pass 2: the same box again, then another one, then both in one batch with output
    st = check([narrow], orc.COLLECT_COUNT)
    assert st.chunks_total == 62 and st.chunks_skipped > 40 and st.points_scanned < 0.35 * st.points_total
    st = check([two], orc.COLLECT_BUFFER)
    assert st.chunks_skipped > 40
    st = check([narrow, two], orc.COLLECT_BUFFER, per_file=True)
    assert st.chunks_skipped > 30 and st.segments >= 2  # two separate runs of the big file
    st = check([narrow, two], orc.COLLECT_GRID)
    assert st.chunks_skipped > 30
    # class queries: LAST has no class part yet -> full pass that builds it, then filtered passes
    st = check([dict(cls=7)], orc.COLLECT_BUFFER)
    if layout == "last":
        assert st.chunks_total == 0 and ix.info(0) == (62, True, True) and ix.info(1) == (7, True, True)  # output needs positions too
    else:
        assert st.chunks_skipped > 0
    st = check([dict(cls=7)], orc.COLLECT_BUFFER, per_file=True)
    assert 0 < st.chunks_skipped < st.chunks_total
    st = check([dict(cls=9)], orc.COLLECT_COUNT)
    assert st.points_scanned == 0 and st.chunks_skipped == st.chunks_total == 69
    st = check([dict(cls=9), narrow], orc.COLLECT_BUFFER)  # mixed batch: union of what the two queries need
    assert 0 < st.chunks_skipped < st.chunks_total
    # an unfiltered pass over the same list (index = None) still gives the same answers
    got = host_pass(pcq, ctx, images, [narrow], orc.COLLECT_BUFFER, None)
    assert_same(orc.COLLECT_BUFFER, got[0], oracle_run(files, exts, orc.COLLECT_BUFFER, **narrow))
    ix.close()


def test_host_index_from_pageable_memory_and_list_mismatch(pcq, ctx):
    rng = np.random.default_rng(123)
    f = strip_file(rng, 30 * CH + 5, 1, "las", 0)  # a numpy array: pageable memory -> pinned bounce ring
    ix = pcq.HostIndex(ctx)
    b = xbox(0, 10 * CH * 40, 12 * CH * 40)
    want = oracle_run([f], ["las"], orc.COLLECT_BUFFER, bounds=b)
    for k in range(3):
        got = host_pass(pcq, ctx, [(f, "las")], [dict(bounds=b)], orc.COLLECT_BUFFER, ix)
        assert_same(orc.COLLECT_BUFFER, got[0], want)
        st = ctx.last_scan_stats
        assert (st.chunks_skipped > 20) == (k > 0)
    with pytest.raises(pcq.PcqError) as e:
        host_pass(pcq, ctx, [(f, "las"), (f, "las")], [dict(bounds=b)], orc.COLLECT_COUNT, ix)
    assert e.value.code == pcq.binding.PCQ_ERR_ARG
    ix.close()


def test_last_class_count_over_hundreds_of_runs(pcq, ctx):
    """More than 256 runs: a LAST class count leaves k_class_count_soa (which walks the segments one by one with the
    whole grid) for the tile-scheduled scan."""
    rng = np.random.default_rng(8)
    n_chunks = 9 * 260
    n = n_chunks * CH
    xyz = rng.integers(-1000, 1000, size=(n, 3), dtype=np.int32)
    cls = rng.choice(np.array([1, 2, 5], np.uint8), size=n)
    idx = np.flatnonzero((np.arange(n) // CH) % 9 == 4)  # class 11 in one chunk out of nine
    idx = idx[rng.random(idx.shape[0]) < 0.01]
    cls[idx] = 11
    f = make_file(xyz, cls, fmt=0, scale=HEADERS[0][0], offset=HEADERS[0][1], layout="last", seed=4)
    df = pcq.DeviceFile.stage(ctx, f, "last")
    df.build_index()
    got = search(pcq, ctx, [df], orc.COLLECT_COUNT, cls=11)
    st = ctx.last_scan_stats
    assert st.segments == 260 and st.chunks_total == n_chunks and st.chunks_skipped == n_chunks - 260
    assert got[0].point_count() == idx.shape[0] > 10_000
    assert search(pcq, ctx, [df], orc.COLLECT_COUNT, cls=2)[0].point_count() == int((cls == 2).sum())
    df.release()


def test_chunk_headers_golden_fixtures(pcq, ctx):
    """The committed fixtures (tests/golden/scan_golden.json, generator make_golden.py): SHA-256 of the chunk headers of
    every file of the seeded datasets."""
    import hashlib
    import json
    from pathlib import Path

    from tests.golden.make_golden import CASES, build_case

    golden = json.loads((Path(__file__).parent / "golden" / "scan_golden.json").read_text())
    for name in CASES:
        files, exts, _ = build_case(pcq, name)
        h = hashlib.sha256()
        for f, e in zip(files, exts):
            df = pcq.DeviceFile.stage(ctx, f, e)
            df.build_index()
            h.update(np.ascontiguousarray(df.index).tobytes())
            df.release()
        assert h.hexdigest() == golden[name]["chunk_headers_sha256"], name
