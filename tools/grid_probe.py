#!/usr/bin/env python
"""One kernel family under ncu: the density insert (C4 navvis, --density 0.1) or the LAST class select (C3 shape).

    python tools/grid_probe.py grid XL        # k_scan_* in MODE_GRID over the navvis-XL box
    python tools/grid_probe.py lastsel 6      # k_select_bytes for class 6 over 8 ca13-shape LAST files
"""
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from pcq_import import pcq

    S, B = pcq.synth, pcq.binding
    what = sys.argv[1] if len(sys.argv) > 1 else "grid"
    arg = sys.argv[2] if len(sys.argv) > 2 else ("XL" if what == "grid" else "6")
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ctx = pcq.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    impl = pcq.SearchImplementation.Optimized

    def resident(sp):
        buf = torch.empty(sp.n_points * sp.record_len + 256, dtype=torch.uint8, device="cuda:0")
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        return pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)

    def timed(fn):
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.synchronize()
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts), min(ts)

    if what == "grid":
        sp = S.navvis_spec()
        df = resident(sp)
        box = {"S": S.NAVVIS_S, "L": S.NAVVIS_L, "XL": S.NAVVIS_XL}[arg]
        s = pcq.BoundsSearcher(*box)
        g = pcq.GridSampledCollector(box[0], box[1], S.NAVVIS_DENSITY, ctx=ctx)

        def run():
            g.reset()
            s.search_files([df], impl, [g])

        run()
        med, mn = timed(run)
        import time

        t0 = time.perf_counter()
        cells = g.point_count()
        fin = (time.perf_counter() - t0) * 1e3
        print(f"grid {arg}: insert median {med:.3f} ms (min {mn:.3f}), finalize {fin:.3f} ms host clock, {cells} cells")
    else:
        specs = S.ca13_specs()[:8]
        dfs = [resident(sp) for sp in specs]
        s = pcq.ClassSearcher(int(arg))
        cols = [pcq.BufferCollector(ctx) for _ in dfs]

        def run():
            for c in cols:
                c.reset()
            s.search_files(dfs, impl, cols)

        run()
        med, mn = timed(run)
        n = sum(sp.n_points for sp in specs)
        m = sum(c.point_count() for c in cols)
        print(f"LAST class {arg} select over {n} points: median {med:.3f} ms (min {mn:.3f}), {m} matches, "
              f"{(n + m * 43) / med / 1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
