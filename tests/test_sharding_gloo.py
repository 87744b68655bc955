"""Host-side multi-rank logic on CPU: world_size-2 gloo runs of the sharding plan, the count sum and the
density candidate exchange (the device kernels are emulated here with the oracle / numpy; the same
exchange code moves device buffers over NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest

from oracle import np_oracle as npo
from oracle import oracle as orc
from tests.helpers import make_file, same_point_set


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_plan_point_ranges_properties(pcq):
    sh = pcq.sharding
    for sizes, world in (([1000, 513, 1, 40_000, 7], 2), ([56_200_000], 8), ([31_250_000] * 64, 8), ([5, 5, 5], 4), ([0, 10_000, 0], 3)):
        plan = sh.plan_point_ranges(sizes, world)
        assert len(plan) == world
        flat = [r for ranges in plan for r in ranges]
        starts = np.concatenate([[0], np.cumsum(sizes)])
        # ranges tile the global scan order exactly once, in order
        pos = 0
        for r in flat:
            assert r.scan_base == pos == starts[r.file] + r.first_point and r.n_points > 0
            assert r.first_point % sh.TILE == 0 or r.first_point == 0
            pos += r.n_points
        assert pos == sum(sizes)
        per_rank = [sum(r.n_points for r in ranges) for ranges in plan]
        if sum(sizes) > 100 * sh.TILE * world:
            assert max(per_rank) - min(per_rank) <= 2 * sh.TILE + max(0, 0)
    assert sh.plan_files(64, 8) == [list(range(8 * r, 8 * r + 8)) for r in range(8)]
    assert sorted(sum(sh.plan_files(10, 4), [])) == list(range(10))


def test_owner_hash_matches_device_constants(pcq):
    sh = pcq.sharding
    # murmur3 fmix64 known values
    assert int(sh.mix64(np.array([0], np.uint64))[0]) == 0
    assert int(sh.mix64(np.array([1], np.uint64))[0]) == 0xB456BCFC34C2CB2C
    o = sh.owner_of(np.arange(10_000, dtype=np.uint64), 8)
    assert o.min() == 0 and o.max() == 7 and abs(np.bincount(o).mean() - 1250) < 1e-9 and np.bincount(o).min() > 1100


def _matches_with_index(buf, qmin, qmax):
    h = npo.parse_header(buf)
    lo, hi = npo.local_bounds(h, qmin, qmax)
    xyz, cls, rgb = npo._columns(buf, h, "las")
    x = xyz.astype(np.int64)
    sel = np.ones(h["n"], bool)
    for a in range(3):
        sel &= (x[:, a] >= lo[a]) & (x[:, a] <= hi[a])
    return np.nonzero(sel)[0], npo.search_bounds(buf, "las", qmin, qmax)


def _local_candidates(pcq, g, idx, pts, scan_base):
    """what the insert + export kernels produce: one candidate per locally occupied cell"""
    best = {}
    for i, p in zip(idx, pts):
        pos = [float(v) for v in p["pos"]]
        key, aliased = g.key_of(pos)
        assert not aliased
        c = g._cell(pos)
        centre = [np.float64(float(c[a]) + 0.5) * np.float64(g.cell) + np.float64(g.bmin[a]) for a in range(3)]
        d = [centre[a] - np.float64(pos[a]) for a in range(3)]
        dist = float((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
        cand = (np.float64(dist).view(np.uint64), scan_base + int(i))
        if key not in best or cand < best[key][0]:
            best[key] = (cand, p)
    out = np.zeros(len(best), dtype=pcq.CANDIDATE_DTYPE)
    for j, (key, ((dbits, gidx), p)) in enumerate(best.items()):
        out[j]["key"], out[j]["dist_bits"], out[j]["scan_idx"], out[j]["point"] = key, dbits, gidx, p
    return out


def _worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist

    from pcq_import import pcq

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = pcq.sharding
    rng = np.random.default_rng(123)  # same dataset on every rank
    sizes = [3000, 1500, 2777]
    scale, offset = (0.001,) * 3, (-23.108, -21.261, -10.029)
    arrays = [(rng.integers(0, 60_000, size=(n, 3), dtype=np.int32), rng.integers(0, 5, size=n).astype(np.uint8)) for n in sizes]
    qmin, qmax = (-20.0, -18.0, -9.0), (30.0, 33.0, 45.0)
    cell = 3.7

    # ---- count: whole-file sharding + host-side sum ----
    my_files = sh.plan_files(len(sizes), world)[rank]
    local = 0
    for f in my_files:
        img = make_file(arrays[f][0], arrays[f][1], fmt=2, scale=scale, offset=offset, seed=f)
        c = orc.Collector(orc.COLLECT_COUNT)
        orc.search_file(img, "las", c, bounds=(qmin, qmax))
        local += c.point_count()
    t = torch.tensor([local], dtype=torch.int64)
    dist.all_reduce(t)
    total_count = int(t.item())

    # ---- density: point-range sharding + candidate all-to-all ----
    plan = sh.plan_point_ranges(sizes, world)[rank]
    g = npo.SparseGrid(qmin, qmax, cell)
    cands = []
    for r in plan:
        xyz, cls = arrays[r.file]
        full = make_file(xyz, cls, fmt=2, scale=scale, offset=offset, seed=r.file)
        h = npo.parse_header(full)
        # the range as its own file image: same header fields, records [first, first+n)
        R, off = h["record_len"], h["off"]
        img = np.concatenate([full[:off], full[off + r.first_point * R: off + (r.first_point + r.n_points) * R]]).copy()
        img[107:111] = np.frombuffer(np.uint32(r.n_points).tobytes(), np.uint8)
        idx, pts = _matches_with_index(img, qmin, qmax)
        cands.append(_local_candidates(pcq, g, idx, pts, r.scan_base))
    cands = np.concatenate(cands) if cands else np.zeros(0, pcq.CANDIDATE_DTYPE)
    # a rank may hold two candidates for one key (two ranges): the owner's merge handles that
    own = sh.owner_of(cands["key"], world)
    order = np.argsort(own, kind="stable")
    send_counts = [int((own == p).sum()) for p in range(world)]
    send = torch.from_numpy(np.ascontiguousarray(cands[order]).view(np.uint8).reshape(-1).copy())
    recv, recv_counts = sh.exchange_parts(send, send_counts, 64)
    got = np.frombuffer(recv.numpy().tobytes(), dtype=pcq.CANDIDATE_DTYPE)
    assert len(got) == sum(recv_counts) and (sh.owner_of(got["key"], world) == rank).all()
    merged = {}
    for c in got:
        k = int(c["key"])
        v = (int(c["dist_bits"]), int(c["scan_idx"]))
        if k not in merged or v < merged[k][0]:
            merged[k] = (v, c["point"].copy())
    mine = np.array([p for _, p in merged.values()], dtype=pcq.POINT_DTYPE) if merged else np.zeros(0, pcq.POINT_DTYPE)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tobytes())
    # ---- select: point-range sharding, per-range record streams concatenated in scan order (no collective) ----
    sel = []
    for r in plan:
        xyz, cls = arrays[r.file]
        full = make_file(xyz, cls, fmt=2, scale=scale, offset=offset, seed=r.file)
        h = npo.parse_header(full)
        R, off = h["record_len"], h["off"]
        img = np.concatenate([full[:off], full[off + r.first_point * R: off + (r.first_point + r.n_points) * R]]).copy()
        img[107:111] = np.frombuffer(np.uint32(r.n_points).tobytes(), np.uint8)
        c = orc.Collector(orc.COLLECT_BUFFER)
        orc.search_file(img, "las", c, bounds=(qmin, qmax))
        sel.append(c.points())
    seq = sh.gather_selected(plan, sel, len(sizes), per_file=False)
    per = sh.gather_selected(plan, sel, len(sizes), per_file=True)
    if rank == 0:
        allp = np.frombuffer(b"".join(gathered), dtype=pcq.POINT_DTYPE)
        np.save(os.path.join(tmpdir, "density.npy"), allp.view(np.uint8))
        np.save(os.path.join(tmpdir, "count.npy"), np.array([total_count]))
        np.save(os.path.join(tmpdir, "select_seq.npy"), seq.view(np.uint8))
        for f, a in enumerate(per):
            np.save(os.path.join(tmpdir, f"select_{f}.npy"), a.view(np.uint8))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_count_and_density_exchange(pcq, tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # the single-process reference answer: one sequential fold over all files (run_search_sequential)
    rng = np.random.default_rng(123)
    sizes = [3000, 1500, 2777]
    scale, offset = (0.001,) * 3, (-23.108, -21.261, -10.029)
    qmin, qmax = (-20.0, -18.0, -9.0), (30.0, 33.0, 45.0)
    og = orc.Collector(orc.COLLECT_GRID, qmin, qmax, 3.7)
    oc = orc.Collector(orc.COLLECT_COUNT)
    ob = orc.Collector(orc.COLLECT_BUFFER)
    per_file = []
    for f, n in enumerate(sizes):
        xyz, cls = rng.integers(0, 60_000, size=(n, 3), dtype=np.int32), rng.integers(0, 5, size=n).astype(np.uint8)
        img = make_file(xyz, cls, fmt=2, scale=scale, offset=offset, seed=f)
        orc.search_file(img, "las", og, bounds=(qmin, qmax))
        orc.search_file(img, "las", oc, bounds=(qmin, qmax))
        orc.search_file(img, "las", ob, bounds=(qmin, qmax))
        one = orc.Collector(orc.COLLECT_BUFFER)
        orc.search_file(img, "las", one, bounds=(qmin, qmax))
        per_file.append(one.points())
    # select over sharded point ranges == one BufferCollector over all files / one per file, byte for byte
    from tests.helpers import same_point_seq

    assert same_point_seq(np.load(tmp_path / "select_seq.npy").view(pcq.POINT_DTYPE), ob.points())
    for f, want in enumerate(per_file):
        assert same_point_seq(np.load(tmp_path / f"select_{f}.npy").view(pcq.POINT_DTYPE), want)
    got = np.load(tmp_path / "density.npy").view(pcq.POINT_DTYPE)
    assert int(np.load(tmp_path / "count.npy")[0]) == oc.point_count() > 1000
    assert og.point_count() > 300
    assert same_point_set(got, og.points())
