"""GPU-vs-oracle parity at the REAL sizes of BASELINE.json's configurations ("no parity test at a BASELINE
configuration's real size"): C1 proper (10 M points), one full C2 tile (31.25 M points) and one full C3 file
(40.75 M points, LAST).  Counts, SHA-256 of the scan-ordered record streams and density record sets must be the
oracle's.  The files are generated on the host by libpcq_synth.so (threads), scanned by the C oracle on the host and by
the CUDA path through the C ABI."""
import hashlib
import threading

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import same_point_set

pytestmark = pytest.mark.gpu


def host_file_threads(S, sp, n_threads=16):
    R = sp.record_len
    img = np.empty(227 + sp.n_points * R, dtype=np.uint8)
    N = sp.n_points
    mms = [None] * n_threads
    if sp.layout == 0:
        cuts = [N * i // n_threads for i in range(n_threads)] + [N]

        def work(i):
            mms[i] = S.host_point_range(sp, cuts[i], cuts[i + 1] - cuts[i], img[227 + cuts[i] * R: 227 + cuts[i + 1] * R])

        ths = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        mm = [min(m[a] for m in mms) for a in range(3)] + [max(m[a] for m in mms) for a in range(3, 6)]
        img[:227] = S.header_bytes(sp, mm)
        return img
    return S.host_file(sp)  # LAST: columns span the whole file, one thread


def sha(points) -> str:
    return hashlib.sha256(np.ascontiguousarray(points).view(np.uint8).tobytes()).hexdigest()


def run_both(pcq, ctx, img, ext, kind, bounds=None, cls=None, grid=None):
    oc = orc.Collector(kind, *(grid or ()))
    orc.search_file(img, ext, oc, bounds=bounds, cls=cls)
    if kind == orc.COLLECT_GRID:
        gc = pcq.GridSampledCollector(grid[0], grid[1], grid[2], ctx=ctx)
    else:
        gc = (pcq.CountCollector if kind == orc.COLLECT_COUNT else pcq.BufferCollector)(ctx)
    s = pcq.BoundsSearcher(*bounds) if bounds is not None else pcq.ClassSearcher(cls)
    df = run_both.files.get(id(img))
    if df is None:
        df = run_both.files[id(img)] = pcq.DeviceFile.stage(ctx, img, ext)
    s.search_files([df], pcq.SearchImplementation.Optimized, [gc])
    return gc, oc


run_both.files = {}


def release_files():
    for df in run_both.files.values():
        df.release()
    run_both.files.clear()


def test_c1_ten_million_points(pcq, ctx):
    S, B = pcq.synth, pcq.binding
    sp = S.make_spec(0x5EED0000 + 1000, 10_000_000, B.LAYOUT_LAS, 1, B.SHAPE_TERRAIN, (0, 0, -9500), (99_999, 99_999, 76_000),
                     (0.01,) * 3, (390000.0, 130000.0, 0.0), S.DOC_CLASSES)
    img = host_file_threads(S, sp)
    g, o = run_both(pcq, ctx, img, "las", orc.COLLECT_COUNT, bounds=S.DOC_S)
    assert g.point_count() == o.point_count() > 0
    g, o = run_both(pcq, ctx, img, "las", orc.COLLECT_BUFFER, bounds=S.DOC_S)
    assert g.point_count() == o.point_count() and sha(g.points()) == sha(o.points())
    g, o = run_both(pcq, ctx, img, "las", orc.COLLECT_BUFFER, cls=6)
    assert g.point_count() == o.point_count() and sha(g.points()) == sha(o.points())
    grid = (S.DOC_S[0], S.DOC_S[1], S.DOC_DENSITY)
    g, o = run_both(pcq, ctx, img, "las", orc.COLLECT_GRID, bounds=S.DOC_S, grid=grid)
    assert g.point_count() == o.point_count() and same_point_set(g.points(), o.points())
    release_files()


def test_c2_one_full_tile(pcq, ctx):
    S = pcq.synth
    import bench

    specs = S.doc_specs(n_files=64, pts_per_file=31_250_000, seed=bench.SEED)
    k = bench.tiles_hit(specs, S.DOC_S)[0]  # a tile all three boxes touch
    img = host_file_threads(S, specs[k])
    for box in (S.DOC_S, S.DOC_L, S.DOC_XL):
        g, o = run_both(pcq, ctx, img, "las", orc.COLLECT_COUNT, bounds=box)
        assert g.point_count() == o.point_count() > 0
    g, o = run_both(pcq, ctx, img, "las", orc.COLLECT_BUFFER, bounds=S.DOC_S)
    assert g.point_count() == o.point_count() and sha(g.points()) == sha(o.points())
    grid = (S.DOC_L[0], S.DOC_L[1], S.DOC_DENSITY)
    g, o = run_both(pcq, ctx, img, "las", orc.COLLECT_GRID, bounds=S.DOC_L, grid=grid)
    assert g.point_count() == o.point_count() and same_point_set(g.points(), o.points())
    release_files()


def test_c3_one_full_last_file(pcq, ctx):
    S = pcq.synth
    sp = S.ca13_specs()[27]
    assert sp.n_points == 40_750_000
    img = host_file_threads(S, sp)
    for klass in (2, 6, 19):
        g, o = run_both(pcq, ctx, img, "last", orc.COLLECT_COUNT, cls=klass)
        assert g.point_count() == o.point_count()
    for klass in (6, 2, 19):  # compacted LAS output of the class queries: sparse, dense, empty
        g, o = run_both(pcq, ctx, img, "last", orc.COLLECT_BUFFER, cls=klass)
        assert g.point_count() == o.point_count() and sha(g.points()) == sha(o.points())
    g, o = run_both(pcq, ctx, img, "last", orc.COLLECT_COUNT, bounds=S.CA13_XL)
    assert g.point_count() == o.point_count() == sp.n_points
    release_files()
