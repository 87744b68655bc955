#!/usr/bin/env python
"""Kernel sweep: every kernel family of the scan path on device-resident synthetic data,
timed with CUDA events (median of --reps launches after warm-up), reported against the measured HBM peak.

    python tools/sweep.py --points 268435456 > gpurun_out/sweep.jsonl

One JSON line per (layout, format, query, collector, variant): achieved GB/s uses the algorithmic bytes of
SURVEY §8d (record_len or 12 or 1 byte read per point, + 31 bytes written per selected record)."""
import argparse
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 28)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--only", default="")
    args = ap.parse_args()

    import torch

    from pcq_import import pcq

    S, B = pcq.synth, pcq.binding
    peak = 6650.0
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text())["hbm_gbs"])
    ctx = pcq.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    impl = pcq.SearchImplementation.Optimized
    N = args.points

    def timed(fn, reps=args.reps):
        fn()
        fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.synchronize()
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    cases = [("las", 0), ("las", 1), ("las", 2), ("las", 3), ("last", 1), ("last", 3)]
    for ext, fmt in cases:
        layout = B.LAYOUT_LAS if ext == "las" else B.LAYOUT_LAST
        sp = S.uniform_spec(N, layout, fmt, extent=1_000_000)
        R = sp.record_len
        buf = torch.empty(N * R + 256, dtype=torch.uint8, device="cuda:0")
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        df = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
        read_bounds = R if ext == "las" else 12
        read_class = R if ext == "las" else 1
        # boxes with ~50 %, ~1 % and 100 % selectivity (uniform cube of 10 km at scale 0.01)
        boxes = {"50pct": ((0, 0, 0), (10000.0, 10000.0, 5000.0)), "1pct": ((0, 0, 0), (2154.0, 2154.0, 2154.0)), "100pct": ((-1, -1, -1), (10001.0, 10001.0, 10001.0))}
        for variant in (2, 1):
            ctx.set_scan_variant(variant)
            vname = {1: "direct", 2: "staged"}[variant]

            def emit(query, collector, ms, read_b, matches, extra=None):
                alg = N * read_b + (matches * 31 if collector == "buffer" else 0) + (matches * 8 if collector == "grid" else 0)
                line = {"layout": ext, "format": fmt, "record_len": R, "query": query, "collector": collector, "variant": vname,
                        "points": N, "matches": matches, "ms": ms, "gpoints_per_s": N / ms / 1e6, "algorithmic_gb": alg / 1e9,
                        "achieved_gbs": alg / ms / 1e6, "frac_of_measured_peak": alg / ms / 1e6 / peak}
                if extra:
                    line.update(extra)
                print(json.dumps(line), flush=True)

            if args.only and args.only not in ("count", "all"):
                pass
            # ---- count ----
            if not args.only or args.only in ("count", "all"):
                for name in ("50pct",):
                    c = pcq.CountCollector(ctx)
                    s = pcq.BoundsSearcher(*boxes[name])
                    ms = timed(lambda: s.search_files([df], impl, [c]))
                    c.reset()
                    s.search_files([df], impl, [c])
                    emit(f"bounds_{name}", "count", ms, read_bounds, c.point_count())
                c = pcq.CountCollector(ctx)
                s = pcq.ClassSearcher(2)
                ms = timed(lambda: s.search_files([df], impl, [c]))
                c.reset()
                s.search_files([df], impl, [c])
                emit("class_2", "count", ms, read_class, c.point_count())
            # ---- select (BufferCollector) ----
            if not args.only or args.only in ("select", "all"):
                for name in ("1pct", "50pct", "100pct"):
                    c = pcq.BufferCollector(ctx)
                    s = pcq.BoundsSearcher(*boxes[name])

                    def run():
                        c.reset()
                        s.search_files([df], impl, [c])

                    ms = timed(run)
                    emit(f"bounds_{name}", "buffer", ms, read_bounds, c.point_count())
                    c.close()
                c = pcq.BufferCollector(ctx)
                s = pcq.ClassSearcher(6)

                def run():
                    c.reset()
                    s.search_files([df], impl, [c])

                ms = timed(run)
                emit("class_6", "buffer", ms, read_class, c.point_count())
                c.close()
            # ---- density (GridSampledCollector): scan + insert, and finalisation separately ----
            if not args.only or args.only in ("grid", "all"):
                for name, cell in (("50pct", 25.0), ("100pct", 10.0)):
                    g = pcq.GridSampledCollector(boxes[name][0], boxes[name][1], cell, ctx=ctx)
                    s = pcq.BoundsSearcher(*boxes[name])

                    def run():
                        g.reset()
                        s.search_files([df], impl, [g])

                    ms = timed(run, reps=3)
                    run()
                    import time

                    t0 = time.perf_counter()
                    cells = g.point_count()
                    fin_ms = (time.perf_counter() - t0) * 1e3
                    cm = pcq.CountCollector(ctx)
                    s.search_files([df], impl, [cm])
                    emit(f"bounds_{name}_cell{cell:g}", "grid", ms, read_bounds, cm.point_count(), {"cells": cells, "finalize_ms_host_timed": fin_ms})
                    g.close()
        ctx.set_scan_variant(0)
        df.release()
        del buf
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
