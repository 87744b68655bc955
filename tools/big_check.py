#!/usr/bin/env python
"""Maximum-size property check: ONE file of more than 2^32 points resident in HBM (LAS format 0, 20 B/pt, or LAST),
checked through size-independent properties — no oracle can scan this in reasonable time.

    python tools/big_check.py --points 5000000000 --layout las      # 100 GB of records

Properties: a box that covers everything counts N; class counts over all 256 byte values add up to N; two halves of
a box add up to the whole; BufferCollector returns exactly as many records as CountCollector counts, all inside the
box, in non-decreasing scan order of a monotone key; the density grid returns one point per occupied cell."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=5_000_000_000)
    ap.add_argument("--layout", default="las", choices=["las", "last"])
    args = ap.parse_args()
    import numpy as np
    import torch

    from pcq_import import pcq

    S, B = pcq.synth, pcq.binding
    ctx = pcq.Context(0)
    impl = pcq.SearchImplementation.Optimized
    N = args.points
    layout = B.LAYOUT_LAS if args.layout == "las" else B.LAYOUT_LAST
    sp = S.uniform_spec(N, layout, 0, extent=1_000_000)
    buf = torch.empty(N * sp.record_len + 256, dtype=torch.uint8, device="cuda:0")
    t0 = time.perf_counter()
    mm, desc = S.device_points(ctx, sp, buf.data_ptr())
    ctx.synchronize()
    df = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
    out = {"points": N, "layout": args.layout, "bytes": N * sp.record_len, "synth_s": round(time.perf_counter() - t0, 2)}

    def count(searcher):
        c = pcq.CountCollector(ctx)
        searcher.search_files([df], impl, [c])
        return c.point_count()

    everything = pcq.BoundsSearcher((-1, -1, -1), (10001.0, 10001.0, 10001.0))
    assert count(everything) == N, "a box around everything must count every point"
    per_class = sum(count(pcq.ClassSearcher(k)) for k in range(256))
    assert per_class == N, f"class counts add up to {per_class}, not {N}"
    whole = count(pcq.BoundsSearcher((0, 0, 0), (10000.0, 10000.0, 50.0)))
    left = count(pcq.BoundsSearcher((0, 0, 0), (4000.0, 10000.0, 50.0)))
    right = count(pcq.BoundsSearcher((4000.01, 0, 0), (10000.0, 10000.0, 50.0)))
    assert left + right == whole and 0 < left < whole, (left, right, whole)
    out.update(count_everything=N, slab_matches=whole)
    # select: ~0.5 % of N records (a thin slab), scan order == file order
    sel = pcq.BoundsSearcher((0, 0, 0), (10000.0, 10000.0, 50.0))
    bc = pcq.BufferCollector(ctx)
    t0 = time.perf_counter()
    sel.search_files([df], impl, [bc])
    n_sel = bc.point_count()
    out["select_s"] = round(time.perf_counter() - t0, 3)
    assert n_sel == whole, (n_sel, whole)
    pts = bc.points()
    assert len(pts) == whole
    assert float(pts["pos"][:, 2].max()) <= 50.0 + 1e-9 and float(pts["pos"].min()) >= 0.0
    # the same records again: idempotent, identical stream
    bc2 = pcq.BufferCollector(ctx)
    sel.search_files([df], impl, [bc2])
    assert np.array_equal(bc2.points().view(np.uint8), pts.view(np.uint8))
    # density over the slab: one point per occupied cell
    g = pcq.GridSampledCollector((0, 0, 0), (10000.0, 10000.0, 50.0), 25.0, ctx=ctx)
    sel.search_files([df], impl, [g])
    gp = g.points()
    cell = np.floor(gp["pos"] / 25.0).astype(np.int64)
    assert len(np.unique(cell, axis=0)) == len(gp) == g.point_count()
    out["density_cells"] = len(gp)
    out["ok"] = True
    print(json.dumps(out))


if __name__ == "__main__":
    main()
