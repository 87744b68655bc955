"""GPU parity tests of the multi-GPU layer (pcq_group_*, csrc/group.cu) against the CPU oracle.

A group of n members is ONE process driving n GPUs.  On a box with fewer GPUs the members share a device (the
exchange then travels as peer copies instead of NCCL send / recv), so that the whole protocol — shard planning,
per-file count reduction, scan-ordered concatenation of selected records, the density all-to-all by cell owner and the
routing of the points of aliased SparseGrid keys to the key's owner — runs, and is compared bit for bit, on the
driver's single-GPU box as well.  The NCCL transport proper is covered by tests/dist_group_nccl.py (torchrun, one
process per GPU) when the box has two GPUs or more.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import make_file, same_point_seq, same_point_set
from tests.test_gpu_parity import ALIAS_BOX, ALIAS_CELL, HEADERS, _alias_files, box, oracle_run, random_file

pytestmark = pytest.mark.gpu

KINDS = {"count": orc.COLLECT_COUNT, "buffer": orc.COLLECT_BUFFER, "grid": orc.COLLECT_GRID}


def make_group(pcq, n):
    import torch

    have = torch.cuda.device_count()
    devices = list(range(n)) if have >= n else [i % have for i in range(n)]
    return pcq.Group.local(n, devices)


@pytest.fixture(scope="module")
def groups(pcq):
    made = {}

    def get(n):
        if n not in made:
            made[n] = make_group(pcq, n)
        return made[n]

    yield get
    for g in made.values():
        g.close()


def group_run(pcq, g, files, exts, kind, bounds=None, cls=None, grid=None, per_file=False, host=False, shard_mode=0):
    searcher = pcq.BoundsSearcher(*bounds) if bounds is not None else pcq.ClassSearcher(cls)
    images = list(zip(files, exts))
    if host:
        return g.search_host_files(images, [searcher], kind, per_file, grid=grid, shard_mode=shard_mode)[0]
    ds = g.stage_host_files(images, shard_mode)
    try:
        return g.search(ds, [searcher], kind, per_file, grid=grid)[0]
    finally:
        ds.release()


def assert_result(kind, res, want_cols):
    counts = res.counts()
    assert counts == [w.point_count() for w in want_cols]
    for lane, w in enumerate(want_cols):
        if kind == orc.COLLECT_COUNT:
            assert res.points(lane) is None
        elif kind == orc.COLLECT_BUFFER:
            assert same_point_seq(res.points(lane), w.points())
        else:
            assert same_point_set(res.points(lane), w.points())


def ragged_files(rng, layout, fmt, hdr):
    # sizes around the shard granularity (8192-point chunks): empty, one chunk, several chunks with a short tail
    return [random_file(rng, n, fmt, layout, hdr) for n in (30_000, 0, 5_000, 70_001, 8_192)]


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("layout,fmt", [("las", 1), ("las", 3), ("last", 2)])
def test_group_equals_oracle_all_collectors(pcq, groups, world, layout, fmt):
    rng = np.random.default_rng(100 + fmt + world)
    hdr = 1
    files = ragged_files(rng, layout, fmt, hdr)
    exts = [layout] * len(files)
    b = box(hdr)
    grid = (b[0], b[1], (b[1][0] - b[0][0]) / 37.0)
    g = groups(world)
    for name, kind in KINDS.items():
        for per_file in (False, True):
            for q in ({"bounds": b}, {"cls": 6}):
                want = oracle_run(files, exts, kind, grid=grid if kind == orc.COLLECT_GRID else None, per_file=per_file, **q)
                for host in (False, True):
                    res = group_run(pcq, g, files, exts, kind, grid=grid if kind == orc.COLLECT_GRID else None,
                                    per_file=per_file, host=host, **q)
                    assert_result(kind, res, want)
                    res.release()


@pytest.mark.parametrize("world", [2, 4])
def test_group_whole_file_sharding(pcq, groups, world):
    rng = np.random.default_rng(5)
    files = ragged_files(rng, "las", 1, 0)
    exts = ["las"] * len(files)
    b = box(0)
    grid = (b[0], b[1], (b[1][0] - b[0][0]) / 29.0)
    g = groups(world)
    for kind in KINDS.values():
        for per_file in (False, True):
            want = oracle_run(files, exts, kind, bounds=b, grid=grid if kind == orc.COLLECT_GRID else None, per_file=per_file)
            for host in (False, True):
                res = group_run(pcq, g, files, exts, kind, bounds=b, grid=grid if kind == orc.COLLECT_GRID else None,
                                per_file=per_file, host=host, shard_mode=pcq.binding.SHARD_FILES)
                assert_result(kind, res, want)


@pytest.mark.parametrize("world", [2, 3])
def test_group_density_with_aliased_keys(pcq, groups, world):
    """A tenth of the points sit on a max face of a power-of-two axis: their keys' results are sequential folds in
    scan order (grid_sampling.rs:62-102).  Sharded over GPUs, every point of such a key travels to the key's owner and
    is folded there in GLOBAL scan order — the result must be the reference's, not PCQ_ERR_ALIASED."""
    rng = np.random.default_rng(23 + world)
    files = _alias_files(rng, 3, 40_000)
    exts = ["las"] * 3
    grid = (ALIAS_BOX[0], ALIAS_BOX[1], ALIAS_CELL)
    g = groups(world)
    for per_file in (False, True):
        want = oracle_run(files, exts, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid, per_file=per_file)
        for host in (False, True):
            res = group_run(pcq, g, files, exts, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid, per_file=per_file, host=host)
            assert_result(orc.COLLECT_GRID, res, want)
    # class query over the same grid
    want = oracle_run(files, exts, orc.COLLECT_GRID, cls=2, grid=grid)
    res = group_run(pcq, g, files, exts, orc.COLLECT_GRID, cls=2, grid=grid)
    assert_result(orc.COLLECT_GRID, res, want)


def test_group_doc_s_density_25_with_points_on_the_z_face(pcq, groups):
    """The reference's own doc-S / --density 25 configuration (run_query_experiments.rs:117-120): 8 z-cells and an
    inclusive z = 200.00 face, so real data aliases."""
    S = pcq.synth
    rng = np.random.default_rng(77)
    files = []
    for k in range(3):
        n = 60_000
        xyz = np.empty((n, 3), np.int32)
        xyz[:, 0] = rng.integers(0, 60_000, n)           # x in [390000, 390600): partly outside the S box
        xyz[:, 1] = rng.integers(0, 1_000_000, n)        # y in [130000 + k * 3000, ...)
        xyz[:, 2] = rng.integers(-500, 20_500, n)
        xyz[rng.random(n) < 0.02, 2] = 20_000            # exactly z = 200.00
        files.append(make_file(xyz, rng.choice(np.array([2, 6], np.uint8), n), fmt=1, scale=(0.01,) * 3,
                               offset=(390000.0, 130000.0 + 3000.0 * k, 0.0), seed=k))
    exts = ["las"] * 3
    grid = (S.DOC_S[0], S.DOC_S[1], S.DOC_DENSITY)
    assert orc.Grid(*grid).dims_bits()[0][2] == 8
    for per_file in (False, True):
        want = oracle_run(files, exts, orc.COLLECT_GRID, bounds=S.DOC_S, grid=grid, per_file=per_file)
        for world in (2, 3):
            res = group_run(pcq, groups(world), files, exts, orc.COLLECT_GRID, bounds=S.DOC_S, grid=grid, per_file=per_file)
            assert_result(orc.COLLECT_GRID, res, want)


def test_group_batch_of_count_queries_is_asynchronous_and_exact(pcq, groups):
    """S, L and XL in one call; the results are read after a second batch has been queued."""
    S = pcq.synth
    specs = S.doc_specs(n_files=4, pts_per_file=50_000)
    files = [S.host_file(sp) for sp in specs]
    exts = ["las"] * 4
    g = groups(2)
    ds = g.stage_host_files(list(zip(files, exts)))
    searchers = [pcq.BoundsSearcher(*bx) for bx in (S.DOC_S, S.DOC_L, S.DOC_XL)]
    first = g.search(ds, searchers, orc.COLLECT_COUNT, True)
    second = g.search(ds, searchers, orc.COLLECT_COUNT, False)
    for r, bx in zip(first, (S.DOC_S, S.DOC_L, S.DOC_XL)):
        assert r.counts() == [int(v) for v in orc.count_parallel(files, exts, 4, bounds=bx)]
    for r, bx in zip(second, (S.DOC_S, S.DOC_L, S.DOC_XL)):
        assert r.counts() == [int(orc.count_parallel(files, exts, 4, bounds=bx).sum())]
    ds.release()


def test_group_wrapped_device_ranges(pcq, groups):
    """pcq_group_wrap_files: ranges generated straight into HBM (bench.py's path)."""
    import torch

    S, B = pcq.synth, pcq.binding
    g = groups(2)
    specs = S.doc_specs(n_files=4, pts_per_file=40_000, fma_sensitive=True)
    files = [S.host_file(sp) for sp in specs]
    ppf = [sp.n_points for sp in specs]
    plan = pcq.shard_plan(ppf, 2)
    pieces, keep = [], []
    for sh in plan:
        m = g.local_ranks.index(int(sh["rank"]))
        ctx = g.ctx(m)
        sp = specs[int(sh["file"])]
        n, first = int(sh["n_points"]), int(sh["first_point"])
        buf = torch.empty(n * sp.record_len + 256, dtype=torch.uint8, device=f"cuda:{ctx.device}")
        S.device_point_range(ctx.device, sp, first, n, buf.data_ptr())
        d = pcq.FileDesc()
        fimg = files[int(sh["file"])]
        B.check(pcq.lib.pcq_parse_header(C.c_void_p(fimg.ctypes.data), fimg.nbytes, 0, 1, C.byref(d)))
        d.n_points = n
        pieces.append((pcq.DeviceFile.wrap(ctx, d, buf.data_ptr(), first, keepalive=buf), m, int(sh["file"])))
    ds = g.wrap_files(ppf, pieces)
    exts = ["las"] * 4
    grid = (S.DOC_XL[0], S.DOC_XL[1], S.DOC_DENSITY)
    for kind in KINDS.values():
        for per_file in (False, True):
            want = oracle_run(files, exts, kind, bounds=S.DOC_L, grid=grid if kind == orc.COLLECT_GRID else None, per_file=per_file)
            res = g.search(ds, [pcq.BoundsSearcher(*S.DOC_L)], kind, per_file, grid=grid if kind == orc.COLLECT_GRID else None)[0]
            assert_result(kind, res, want)
    ds.release()


def test_group_errors_match_the_single_gpu_path(pcq, groups):
    g = groups(2)
    rng = np.random.default_rng(1)
    f = random_file(rng, 20_000, 1, "las", 0)
    with pytest.raises(pcq.PcqError) as e:  # AABB::from_min_max panics on min > max (main.rs:80)
        group_run(pcq, g, [f], ["las"], orc.COLLECT_COUNT, bounds=([1.0, 0.0, 0.0], [0.0, 1.0, 1.0]))
    assert e.value.code == pcq.binding.PCQ_ERR_PANIC
    with pytest.raises(pcq.PcqError) as e:
        group_run(pcq, g, [f], ["laz"], orc.COLLECT_COUNT, cls=2)
    assert e.value.code == pcq.binding.PCQ_ERR_FORMAT
    bad = f.copy()
    bad[104] = 11  # invalid point format for a bounds search (las.rs:59-60)
    with pytest.raises(pcq.PcqError) as e:
        group_run(pcq, g, [bad], ["las"], orc.COLLECT_COUNT, bounds=box(0))
    assert e.value.code == pcq.binding.PCQ_ERR_FORMAT
    # and the group still works afterwards
    want = oracle_run([f], ["las"], orc.COLLECT_COUNT, bounds=box(0))
    assert_result(orc.COLLECT_COUNT, group_run(pcq, g, [f], ["las"], orc.COLLECT_COUNT, bounds=box(0)), want)


def test_group_over_nccl_one_process_per_gpu(pcq):
    """torchrun, one process per GPU, the library's own NCCL communicator — needs >= 2 GPUs on the box."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the same protocol runs above with peer copies; planning on CPU in tests/test_group_plan.py)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(root, "tests", "dist_group_nccl.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout and "ALL OK" in r.stdout


def test_group_over_nccl_one_process(pcq):
    """one process, n distinct GPUs, ncclCommInitAll + grouped send / recv — needs >= 2 GPUs on the box."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    g = pcq.Group.local(2)
    rng = np.random.default_rng(31)
    files = _alias_files(rng, 2, 40_000)
    grid = (ALIAS_BOX[0], ALIAS_BOX[1], ALIAS_CELL)
    for per_file in (False, True):
        want = oracle_run(files, ["las"] * 2, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid, per_file=per_file)
        assert_result(orc.COLLECT_GRID, group_run(pcq, g, files, ["las"] * 2, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=grid, per_file=per_file), want)
    g.close()
