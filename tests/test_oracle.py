"""CPU tests of the oracle: pinned against the reference's own SparseGrid tests, the survey's
known-answer vectors (tests/golden/kav.json), and an independent numpy restatement."""
import itertools
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as npo
from oracle import oracle as orc
from tests.helpers import make_file, same_point_seq, same_point_set

KAV = json.loads((Path(__file__).parent / "golden" / "kav.json").read_text())


def _hdr(scale, offset, fmt=1):
    f = make_file(np.zeros((1, 3), np.int32), [0], fmt=fmt, scale=scale, offset=offset)
    return orc.parse_header(f)


# ---- Rust cast semantics ---------------------------------------------------------------------------
@pytest.mark.parametrize("v,want", [(100.7, 100), (-100.7, -100), (float("nan"), 0), (1e300, 2**63 - 1), (-1e300, -(2**63)),
                                    (5958.999999999999, 5958), (-0.9, 0)])
def test_f64_as_i64(v, want):
    assert orc.lib.orc_f64_as_i64(v) == want
    assert npo.f64_as_i64(v) == want


@pytest.mark.parametrize("v,want", [(3.9, 3), (-3.9, 0), (float("nan"), 0), (1e300, 2**64 - 1), (float("inf"), 2**64 - 1), (-0.0, 0)])
def test_f64_as_u64(v, want):
    assert orc.lib.orc_f64_as_u64(v) == want
    assert npo.f64_as_u64(v) == want


# ---- KAV-1/2: local integer bounds (las.rs:88-99) ----------------------------------------------------
@pytest.mark.parametrize("k", KAV["local_bounds"], ids=lambda k: k["name"])
def test_local_bounds_kav(k):
    h = _hdr(k["scale"], k["offset"])
    lo, hi = orc.local_bounds(h, k["qmin"], k["qmax"])
    assert lo == k["lo"] and hi == k["hi"]
    f = make_file(np.zeros((1, 3), np.int32), [0], scale=k["scale"], offset=k["offset"])
    lo2, hi2 = npo.local_bounds(npo.parse_header(f), k["qmin"], k["qmax"])
    assert lo2 == k["lo"] and hi2 == k["hi"]


def test_local_bounds_use_x_scale_for_min_y_z():
    # las.rs:91-92 divide min.y / min.z by x_scale_factor; with anisotropic scales that is observable
    h = _hdr((0.01, 0.1, 0.5), (0, 0, 0))
    lo, hi = orc.local_bounds(h, (1, 1, 1), (100, 100, 100))
    assert lo == [100, 100, 100] and hi == [10000, 1000, 200]


def test_local_bounds_inverted_panics():
    # anisotropic scales can push the (x-scaled) min above the max: AABB::from_min_max panics
    h = _hdr((0.01, 1.0, 1.0), (0, 0, 0))
    with pytest.raises(orc.OracleError) as e:
        orc.local_bounds(h, (0, 5, 5), (10, 10, 10))
    assert e.value.code == orc.ORC_ERR_PANIC


def test_kav2_point_on_max_face_rejected():
    # navvis-XL: 5.959 / 0.001 = 5958.999999999999 -> 5958, so raw z = 5959 is rejected by Optimized
    xyz = np.array([[0, 0, 5959], [0, 0, 5958]], np.int32)
    f = make_file(xyz, [1, 1], fmt=0, scale=(0.001,) * 3, offset=(0, 0, 0))
    c = orc.Collector(orc.COLLECT_BUFFER)
    orc.search_file(f, "las", c, bounds=((-23.108, -21.261, -10.029), (28.588, 27.123, 5.959)))
    pts = c.points()
    assert len(pts) == 1 and pts["pos"][0][2] == 5958 * 0.001


def test_kav3_truncation_asymmetry():
    # quotient +100.7 -> 100 admits raw 100 (position below q.min); -100.7 -> -100 admits raw -100 above q.max
    xyz = np.array([[100, 0, 0], [101, 0, 0], [-100, 0, 0], [-101, 0, 0]], np.int32)
    f = make_file(xyz, [0] * 4, fmt=0, scale=(1.0,) * 3, offset=(0, 0, 0), hdr_min=(-1000,) * 3, hdr_max=(1000,) * 3)
    c = orc.Collector(orc.COLLECT_BUFFER)
    orc.search_file(f, "las", c, bounds=((100.7, -1, -1), (500, 1, 1)))
    assert list(c.points()["pos"][:, 0]) == [100.0, 101.0]
    c = orc.Collector(orc.COLLECT_BUFFER)
    orc.search_file(f, "las", c, bounds=((-500, -1, -1), (-100.7, 1, 1)))
    assert list(c.points()["pos"][:, 0]) == [-100.0, -101.0]


# ---- KAV-4: position reconstruction is two roundings, never an FMA (las.rs:139-141) ------------------
@pytest.mark.parametrize("k", KAV["reconstruct"], ids=lambda k: k["name"])
def test_reconstruct_no_fma(k):
    xyz = np.array([[k["raw"], k["raw"], k["raw"]]], np.int32)
    f = make_file(xyz, [2], fmt=0, scale=(k["scale"],) * 3, offset=(k["offset"],) * 3)
    c = orc.Collector(orc.COLLECT_BUFFER)
    orc.search_file(f, "las", c, cls=2)
    got = float(c.points()["pos"][0][0])
    assert got.hex() == float.fromhex(k["hex"]).hex()
    assert got.hex() != float.fromhex(k["fused_hex"]).hex()
    assert float(npo.search_class(f, "las", 2)["pos"][0][1]).hex() == float.fromhex(k["hex"]).hex()


# ---- KAV-5: SparseGrid::new (grid_sampling.rs:18-47) -------------------------------------------------
@pytest.mark.parametrize("k", KAV["grid_params"], ids=lambda k: k["name"])
def test_grid_params_kav(k):
    g = orc.Grid(k["min"], k["max"], k["cell"])
    dims, bits = g.dims_bits()
    assert dims == k["dims"] and bits == k["bits"]
    g2 = npo.SparseGrid(k["min"], k["max"], k["cell"])
    assert g2.dims == k["dims"] and g2.bits == k["bits"]


def test_grid_too_many_cells():
    with pytest.raises(orc.OracleError) as e:
        orc.Grid((0, 0, 0), (1e9, 1e9, 1e9), 1e-3)  # 40 bits per axis
    assert e.value.code == orc.ORC_ERR_GRID


# ---- KAV-7: the reference's own unit tests (grid_sampling.rs:121-208) --------------------------------
@pytest.mark.parametrize("k", KAV["grid_tests"], ids=lambda k: k["name"])
def test_reference_sparse_grid_tests(k):
    g = orc.Grid(k["min"], k["max"], k["cell"])
    for p in k["insert"]:
        g.insert_point(p)
    assert sorted(int(c) for c in g.cells()) == k["cells"]
    got = sorted(tuple(p) for p in g.points()["pos"].tolist())
    assert got == sorted(tuple(p) for p in k["points"])
    g2 = npo.SparseGrid(k["min"], k["max"], k["cell"])
    for p in k["insert"]:
        q = np.zeros(1, npo.POINT_DTYPE)[0]
        q["pos"] = p
        g2.insert_point(q)
    assert sorted(g2.cells.keys()) == k["cells"]
    assert sorted(tuple(p) for p in g2.points()["pos"].tolist()) == got


# ---- KAV-6: masked key vs unmasked centre makes the fold order-dependent ------------------------------
def test_alias_order_dependence():
    k = KAV["alias_order"]
    pts = {"A": k["A"], "B": k["B"], "C": k["C"]}
    for order, winner in k["winners"].items():
        g = orc.Grid(k["min"], k["max"], k["cell"])
        for name in order:
            g.insert_point(pts[name])
        assert list(g.cells()) == [0]
        assert g.points()["pos"][0].tolist() == [float(v) for v in pts[winner]], order
    g = orc.Grid(k["min"], k["max"], k["cell"])
    assert g.key(k["B"]) == (0, True) and g.key(k["A"]) == (0, False)


# ---- C oracle vs the independent numpy restatement ---------------------------------------------------
def _random_file(rng, n, fmt, layout, scale, offset, record_len=None):
    xyz = rng.integers(-50_000, 150_000, size=(n, 3), dtype=np.int32)
    cls = rng.choice(np.array([1, 2, 2, 2, 5, 6, 6 | 0x20, 2 | 0x80, 9], np.uint8), size=n)
    return make_file(xyz, cls, fmt=fmt, scale=scale, offset=offset, layout=layout, record_len=record_len,
                     seed=int(rng.integers(1 << 30)))


HEADERS = [((0.01, 0.01, 0.01), (390000.0, 130000.0, 0.0)), ((0.001, 0.001, 0.001), (-23.108, -21.261, -10.029)),
           ((0.001, 0.002, 0.00025), (643431.76, 3883547.565, 1234.5))]


@pytest.mark.parametrize("layout", ["las", "last"])
@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("hdr", range(len(HEADERS)))
def test_c_oracle_matches_numpy(layout, fmt, hdr):
    rng = np.random.default_rng(1000 * hdr + 10 * fmt + (layout == "last"))
    scale, offset = HEADERS[hdr]
    f = _random_file(rng, 3001, fmt, layout, scale, offset, record_len=None if fmt != 1 else 31)
    qmin = [offset[a] + 100 * scale[0] for a in range(3)]
    qmax = [offset[a] + 90_000 * scale[a] for a in range(3)]
    c = orc.Collector(orc.COLLECT_BUFFER)
    orc.search_file(f, layout, c, bounds=(qmin, qmax))
    want = npo.search_bounds(f, layout, qmin, qmax)
    assert 0 < len(want) < 3001
    assert same_point_seq(c.points(), want)
    cc = orc.Collector(orc.COLLECT_COUNT)
    orc.search_file(f, layout, cc, bounds=(qmin, qmax))
    assert cc.point_count() == len(want) and cc.points() is None
    for klass in (2, 6, 19, 6 | 0x20):
        c = orc.Collector(orc.COLLECT_BUFFER)
        orc.search_file(f, layout, c, cls=klass)
        assert same_point_seq(c.points(), npo.search_class(f, layout, klass))


def test_grid_collector_matches_numpy_fold():
    rng = np.random.default_rng(7)
    f = _random_file(rng, 4000, 3, "las", (0.001,) * 3, (-23.108, -21.261, -10.029))
    qmin, qmax = (-40.0, -40.0, -30.0), (60.0, 70.0, 80.0)
    c = orc.Collector(orc.COLLECT_GRID, qmin, qmax, 7.5)
    orc.search_file(f, "las", c, bounds=(qmin, qmax))
    g = npo.SparseGrid(qmin, qmax, 7.5)
    for p in npo.search_bounds(f, "las", qmin, qmax):
        g.insert_point(p)
    assert c.point_count() == len(g.cells) > 10
    assert same_point_set(c.points(), g.points())


def test_file_outside_query_is_skipped_and_bad_headers_fail():
    xyz = np.array([[1, 1, 1]], np.int32)
    f = make_file(xyz, [2], fmt=1)
    c = orc.Collector(orc.COLLECT_COUNT)
    orc.search_file(f, "las", c, bounds=((10, 10, 10), (20, 20, 20)))
    assert c.point_count() == 0
    bad = f.copy()
    bad[0] = ord("X")
    with pytest.raises(orc.OracleError) as e:
        orc.search_file(bad, "las", orc.Collector(orc.COLLECT_COUNT), cls=2)
    assert e.value.code == orc.ORC_ERR_FORMAT
    # LAZ-style "compressed" bit in the format byte: LAST class search masks it (last.rs:222), the others do not
    g = make_file(xyz, [2], fmt=1, layout="last", fmt_byte=0x80 | 1)
    c = orc.Collector(orc.COLLECT_COUNT)
    orc.search_file(g, "last", c, cls=2)
    assert c.point_count() == 1
    with pytest.raises(orc.OracleError):
        orc.search_file(g, "last", orc.Collector(orc.COLLECT_COUNT), bounds=((0, 0, 0), (1, 1, 1)))
    # inverted header bounds: AABB::from_min_max panics (las.rs:61)
    inv = make_file(xyz, [2], fmt=1, hdr_min=(5, 0, 0), hdr_max=(1, 1, 1))
    with pytest.raises(orc.OracleError) as e:
        orc.search_file(inv, "las", orc.Collector(orc.COLLECT_COUNT), bounds=((0, 0, 0), (1, 1, 1)))
    assert e.value.code == orc.ORC_ERR_PANIC
    # truncated file: reads are lazy (Cursor), so only a read that really passes the end is an io error
    c = orc.Collector(orc.COLLECT_COUNT)
    orc.search_file(f[:-3].copy(), "las", c, cls=2)  # only the GPS-time tail of the record is missing
    assert c.point_count() == 1
    with pytest.raises(orc.OracleError) as e:
        orc.search_file(f[:-20].copy(), "las", orc.Collector(orc.COLLECT_COUNT), cls=2)
    assert e.value.code == orc.ORC_ERR_IO


def test_extended_format_class_offset():
    xyz = np.zeros((5, 3), np.int32)
    f = make_file(xyz, [2, 6, 2, 6, 6], fmt=6, version=(1, 4), layout="las")
    c = orc.Collector(orc.COLLECT_COUNT)
    orc.search_file(f, "las", c, cls=6)  # class byte at +16 for formats 6..10 (las.rs:202-205)
    assert c.point_count() == 3


def test_count_parallel_matches_sequential():
    rng = np.random.default_rng(3)
    files = [_random_file(rng, 500 + 37 * i, i % 4, "las" if i % 2 else "last", (0.01,) * 3, (0.0, 0.0, 0.0)) for i in range(7)]
    exts = ["las" if i % 2 else "last" for i in range(7)]
    bounds = ((0.0, 0.0, 0.0), (900.0, 800.0, 700.0))
    got = orc.count_parallel(files, exts, 4, bounds=bounds)
    for i, f in enumerate(files):
        c = orc.Collector(orc.COLLECT_COUNT)
        orc.search_file(f, exts[i], c, bounds=bounds)
        assert c.point_count() == int(got[i])


# ---- committed golden fixtures (tests/golden/scan_golden.json; made by tests/golden/make_golden.py) ----
def test_oracle_reproduces_golden_fixtures(pcq):
    import hashlib

    from tests.golden.make_golden import CASES, build_case, oracle_answer

    golden = json.loads((Path(__file__).parent / "golden" / "scan_golden.json").read_text())
    assert set(golden) == set(CASES)
    for name in CASES:
        files, exts, kw = build_case(pcq, name)
        assert hashlib.sha256(b"".join(f.tobytes() for f in files)).hexdigest() == golden[name]["input_sha256"], "generator drifted"
        got = oracle_answer(files, exts, kw)
        assert got["counts"] == golden[name]["counts"] and got["buffer_sha256"] == golden[name]["buffer_sha256"]
        assert got["chunk_headers_sha256"] == golden[name]["chunk_headers_sha256"]
        hh = hashlib.sha256()
        for f, e in zip(files, exts):
            hh.update(npo.chunk_headers(f, e).tobytes())  # the numpy restatement produces the same bytes
        assert hh.hexdigest() == golden[name]["chunk_headers_sha256"]
        # and the independent numpy restatement agrees with the frozen counts
        for f, e, want in zip(files, exts, golden[name]["counts"]):
            p = npo.search_bounds(f, e, *kw["bounds"]) if "bounds" in kw else npo.search_class(f, e, kw["cls"])
            assert len(p) == want


def test_chunk_headers_restatement_is_a_sound_filter():
    """improvements.md:3-10: a chunk whose header excludes the query holds no match of the per-point search, and the
    matches found chunk by chunk over the surviving chunks are the matches of the whole file (both layouts)."""
    rng = np.random.default_rng(3)
    n, ch = 10_000, 1024
    xyz = np.stack([np.arange(n) * 7 + rng.integers(-50, 50, n), rng.integers(0, 1000, n), rng.integers(0, 100, n)], axis=1).astype(np.int32)
    cls = rng.choice(np.array([1, 2, 5], np.uint8), size=n)
    cls[3 * ch + 5] = 6
    cls[3 * ch + 9] = 6 | 0x40
    for layout in ("las", "last"):
        f = make_file(xyz, cls, fmt=1, layout=layout)
        h = npo.chunk_headers(f, layout, chunk_points=ch)
        assert h.shape[0] == 10 and h["n_points"].sum() == n and h["n_points"][-1] == n - 9 * ch
        assert np.array_equal(h["lo"][2], xyz[2 * ch: 3 * ch].min(axis=0)) and np.array_equal(h["hi"][9], xyz[9 * ch:].max(axis=0))
        has6 = [(int(r["cls_bits"][0]) >> 6) & 1 for r in h]
        assert has6 == [0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
        assert (int(h["cls_bits"][3][(6 | 0x40) >> 5]) >> ((6 | 0x40) & 31)) & 1 == 1
        qmin, qmax = (200.0, -1.0, -1.0), (300.0, 100.0, 100.0)
        hd = npo.parse_header(f)
        lo, hi = npo.local_bounds(hd, qmin, qmax)
        keep = np.all((h["hi"] >= np.array(lo)) & (h["lo"] <= np.array(hi)), axis=1)
        assert 0 < keep.sum() < 5
        want = npo.search_bounds(f, layout, qmin, qmax)
        x = xyz.astype(np.int64)
        m = np.all((x >= np.array(lo)) & (x <= np.array(hi)), axis=1)
        assert m.sum() == want.shape[0] > 0
        assert not m[~np.repeat(keep, ch)[:n]].any()


@pytest.mark.parametrize("layout,fmt", [("las", 0), ("las", 3), ("las", 6), ("last", 1), ("last", 3), ("last", 6)])
def test_chunk_headers_c_and_numpy_restatements_agree(layout, fmt):
    rng = np.random.default_rng(50 + fmt)
    n, ch = 5000, 512
    xyz = rng.integers(-(1 << 31), (1 << 31) - 1, size=(n, 3), dtype=np.int64).astype(np.int32)
    cls = rng.integers(0, 256, size=n).astype(np.uint8)
    f = make_file(xyz, cls, fmt=fmt, layout=layout, record_len={0: 23, 3: None, 6: None, 1: None}.get(fmt),
                  version=(1, 4) if fmt >= 6 else (1, 2))
    a = orc.chunk_headers(f, layout, chunk_points=ch)
    b = npo.chunk_headers(f, layout, chunk_points=ch)
    assert a.shape == b.shape == (10,)
    for k in ("lo", "hi", "cls_bits", "n_points"):
        assert np.array_equal(a[k], b[k]), k
    # a point range, as a rank of a sharded search would index it
    a = orc.chunk_headers(f, layout, chunk_points=ch, first=700, count=1500)
    b = npo.chunk_headers(f, layout, chunk_points=ch, first=700, count=1500)
    for k in ("lo", "hi", "cls_bits", "n_points"):
        assert np.array_equal(a[k], b[k]), k
    assert a["n_points"].tolist() == [512, 512, 476]
