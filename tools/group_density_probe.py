#!/usr/bin/env python
"""Phases of a sharded density query (pcq_group, one process, n members): C4 navvis XL, --density 0.1.

    python tools/group_density_probe.py [n_members] [iterations]
"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from pcq_import import pcq

    S, B = pcq.synth, pcq.binding
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    have = torch.cuda.device_count()
    g = pcq.Group.local(n, [i % have for i in range(n)])
    sp = S.navvis_spec()
    N, R = sp.n_points, sp.record_len
    plan = pcq.shard_plan([N], n)
    pieces, mms = [], []
    for sh in plan:
        m = g.local_ranks.index(int(sh["rank"]))
        ctx = g.ctx(m)
        first, cnt = int(sh["first_point"]), int(sh["n_points"])
        buf = torch.empty(cnt * R + 256, dtype=torch.uint8, device=f"cuda:{ctx.device}")
        mms.append(S.device_point_range(ctx.device, sp, first, cnt, buf.data_ptr()))
        pieces.append((buf, m, first, cnt))
    mm = [min(x[a] for x in mms) for a in range(3)] + [max(x[a] for x in mms) for a in range(3, 6)]
    files = []
    for buf, m, first, cnt in pieces:
        d = S.desc_of(sp, mm)
        d.n_points = cnt
        files.append((pcq.DeviceFile.wrap(g.ctx(m), d, buf.data_ptr(), first, keepalive=buf), m, 0))
    ds = g.wrap_files([N], files)
    box = S.NAVVIS_XL
    s = pcq.BoundsSearcher(*box)
    res = None
    for it in range(iters):
        if res is not None:
            res.release()
        g.synchronize()
        t0 = time.perf_counter()
        res = g.search(ds, [s], B.COLLECT_GRID, False, grid=(box[0], box[1], S.NAVVIS_DENSITY))[0]
        dt = (time.perf_counter() - t0) * 1e3
        st = g.last_stats
        print(f"iter {it}: {dt:.2f} ms, cells {res.counts()[0]}, " + ", ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in st.items()), flush=True)


if __name__ == "__main__":
    main()
