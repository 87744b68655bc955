"""Multi-GPU density query over NCCL (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_density_nccl.py
One navvis-shape file is cut into tile-aligned point ranges (one per rank); every rank scans its range into a
local cell table, the per-cell candidates travel by owner with one all-to-all, owners merge, rank 0 checks the
union against the oracle's single sequential fold.  Also checks the sharded count and the sharded select (per-range
record streams concatenated in range order == one BufferCollector over the whole file)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist

    from oracle import oracle as orc
    from pcq_import import pcq
    from tests.helpers import same_point_set

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    S, sh = pcq.synth, pcq.sharding
    ctx = pcq.Context(local)
    ok = True
    for fma in (False, True):
        img = S.host_file(S.navvis_spec(n_points=1_000_003, fma_sensitive=fma))
        for qmin, qmax in (S.NAVVIS_L, S.NAVVIS_XL):
            ranges = sh.plan_point_ranges([1_000_003], world)[rank]
            dfs = []
            for r in ranges:
                df = pcq.DeviceFile.stage(ctx, img, "las", first_point=r.first_point, n_points=r.n_points)
                df.set_scan_base(r.scan_base)
                dfs.append(df)
            searcher = pcq.BoundsSearcher(qmin, qmax)
            impl = pcq.SearchImplementation.Optimized
            # count: host-side sum, no data-path collective
            cc = pcq.CountCollector(ctx)
            if dfs:
                searcher.search_files(dfs, impl, [cc])
            t = torch.tensor([cc.point_count()], dtype=torch.int64, device=f"cuda:{local}")
            dist.all_reduce(t)
            # select: every rank compacts its ranges in scan order, the host concatenates them in range order
            sel = []
            for df in dfs:
                bc = pcq.BufferCollector(ctx)
                searcher.search_files([df], impl, [bc])
                p = bc.points()
                sel.append(p if p is not None else np.zeros(0, pcq.POINT_DTYPE))
            selected = sh.gather_selected(ranges, sel, 1, per_file=False)
            # density: local table -> candidates by owner -> all-to-all -> merge
            lg = pcq.GridSampledCollector(qmin, qmax, S.NAVVIS_DENSITY, ctx=ctx)
            if dfs:
                searcher.search_files(dfs, impl, [lg])
            merged = sh.density_exchange(lg, lambda: pcq.GridSampledCollector(qmin, qmax, S.NAVVIS_DENSITY, ctx=ctx))
            mine = merged.points()
            gathered = [None] * world
            dist.all_gather_object(gathered, mine.tobytes())
            if rank == 0:
                got = np.frombuffer(b"".join(gathered), dtype=pcq.POINT_DTYPE)
                og = orc.Collector(orc.COLLECT_GRID, qmin, qmax, S.NAVVIS_DENSITY)
                orc.search_file(img, "las", og, bounds=(qmin, qmax))
                oc = orc.Collector(orc.COLLECT_COUNT)
                orc.search_file(img, "las", oc, bounds=(qmin, qmax))
                ob = orc.Collector(orc.COLLECT_BUFFER)
                orc.search_file(img, "las", ob, bounds=(qmin, qmax))
                good = (same_point_set(got, og.points()) and int(t.item()) == oc.point_count()
                        and selected.tobytes() == ob.points().tobytes())
                print(f"world={world} fma={fma} box={qmax}: count {int(t.item())} cells {len(got)} -> {'OK' if good else 'MISMATCH'}", flush=True)
                ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
