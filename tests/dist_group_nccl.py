"""torchrun worker: the multi-GPU layer with one process per GPU (pcq_group_create_rank; the library's own NCCL
communicator, torch.distributed only carries the id).  Every rank builds the same small dataset on the host, the
group shards it, and rank 0 compares every result with the CPU oracle.  Run by tests/test_gpu_group.py."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist

    from oracle import oracle as orc
    from pcq_import import pcq
    from tests.helpers import same_point_seq, same_point_set
    from tests.test_gpu_parity import ALIAS_BOX, ALIAS_CELL, _alias_files, box, oracle_run, random_file

    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # only the NCCL id travels through it
    g = pcq.Group.from_torch_distributed(local)
    ok = True

    def check(name, kind, res, want):
        nonlocal ok
        good = res.counts() == [w.point_count() for w in want]
        if rank == 0 and kind != orc.COLLECT_COUNT:
            for lane, w in enumerate(want):
                same = same_point_seq if kind == orc.COLLECT_BUFFER else same_point_set
                good = good and same(res.points(lane), w.points())
        if not good:
            ok = False
        print(f"rank {rank} {name}: {'OK' if good else 'MISMATCH'}", flush=True)

    rng = np.random.default_rng(4242)  # same seed on every rank: same files
    files = [random_file(rng, n, 1, "las", 1) for n in (50_000, 9_000, 120_001)]
    exts = ["las"] * 3
    images = list(zip(files, exts))
    b = box(1)
    grid = (b[0], b[1], (b[1][0] - b[0][0]) / 41.0)
    ds = g.stage_host_files(images)
    for kname, kind in (("count", orc.COLLECT_COUNT), ("buffer", orc.COLLECT_BUFFER), ("grid", orc.COLLECT_GRID)):
        for per_file in (False, True):
            gr = grid if kind == orc.COLLECT_GRID else None
            want = oracle_run(files, exts, kind, bounds=b, grid=gr, per_file=per_file)
            check(f"{kname} per_file={per_file} resident", kind, g.search(ds, [pcq.BoundsSearcher(*b)], kind, per_file, grid=gr)[0], want)
            check(f"{kname} per_file={per_file} host", kind, g.search_host_files(images, [pcq.BoundsSearcher(*b)], kind, per_file, grid=gr)[0], want)
    ds.release()
    # aliased keys: routed to their owners over NCCL
    rng = np.random.default_rng(99)
    files = _alias_files(rng, 3, 40_000)
    images = list(zip(files, exts))
    agrid = (ALIAS_BOX[0], ALIAS_BOX[1], ALIAS_CELL)
    ds = g.stage_host_files(images)
    for per_file in (False, True):
        want = oracle_run(files, exts, orc.COLLECT_GRID, bounds=ALIAS_BOX, grid=agrid, per_file=per_file)
        check(f"aliased grid per_file={per_file}", orc.COLLECT_GRID,
              g.search(ds, [pcq.BoundsSearcher(*ALIAS_BOX)], orc.COLLECT_GRID, per_file, grid=agrid)[0], want)
    ds.release()
    flags = [None] * dist.get_world_size()
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print("ALL OK" if all(flags) else "SOME MISMATCH", flush=True)
    g.close()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
