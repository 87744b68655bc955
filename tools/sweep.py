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
    ap.add_argument("--cases", default="", help="comma list of layout:format, e.g. las:1,last:3")
    ap.add_argument("--select-queries", default="", help="comma list out of 1pct,50pct,100pct")
    ap.add_argument("--variants", default="2,1", help="scan variants to run: 2 = staged, 1 = direct")
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
        timed.last = ts
        return statistics.median(ts)

    if args.only in ("configs", "c3", "c4"):
        run_configs(pcq, ctx, stream, timed, peak, args)
        return
    if args.only == "index":
        run_index(pcq, ctx, timed, peak, args)
        return
    cases = [("las", 0), ("las", 1), ("las", 2), ("las", 3), ("last", 1), ("last", 3)]
    if args.cases:  # e.g. --cases las:1,last:3
        cases = [(c.split(":")[0], int(c.split(":")[1])) for c in args.cases.split(",")]
    sel_names = tuple(args.select_queries.split(",")) if args.select_queries else ("1pct", "50pct", "100pct")
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
        for variant in [int(v) for v in args.variants.split(",")]:
            ctx.set_scan_variant(variant)
            vname = {1: "direct", 2: "staged"}[variant]

            def emit(query, collector, ms, read_b, matches, extra=None):
                alg = N * read_b + (matches * 31 if collector == "buffer" else 0) + (matches * 8 if collector == "grid" else 0)
                line = {"layout": ext, "format": fmt, "record_len": R, "query": query, "collector": collector, "variant": vname,
                        "points": N, "matches": matches, "ms": ms, "gpoints_per_s": N / ms / 1e6, "algorithmic_gb": alg / 1e9,
                        "achieved_gbs": alg / ms / 1e6, "frac_of_measured_peak": alg / ms / 1e6 / peak}
                line["ms_min"], line["ms_max"] = min(timed.last), max(timed.last)
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
                for name in sel_names:
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


def run_index(pcq, ctx, timed, peak, args):
    """The on-the-fly chunk index (improvements.md:3-10) on one file in acquisition order: 64 flight
    strips stored one after the other.  Build cost, then every query with and without the index (same file)."""
    import time

    import torch

    S, B = pcq.synth, pcq.binding
    impl = pcq.SearchImplementation.Optimized
    cases = [("las", 1), ("last", 1)]
    if args.cases:
        cases = [(c.split(":")[0], int(c.split(":")[1])) for c in args.cases.split(",")]
    for ext, fmt in cases:
        layout = B.LAYOUT_LAS if ext == "las" else B.LAYOUT_LAST
        buf, desc, hdr = S.strips_device(ctx, args.points, 64, layout, fmt)
        N, R = int(desc.n_points), int(desc.record_len)
        plain = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
        indexed = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
        builds = []
        for _ in range(3):
            indexed.drop_index()
            ctx.synchronize()
            t0 = time.perf_counter()
            indexed.build_index()  # synchronises: kernel + copy of the headers to the host
            builds.append((time.perf_counter() - t0) * 1e3)
        build_bytes = N * (R if ext == "las" else 13)
        print(json.dumps({"layout": ext, "format": fmt, "points": N, "index": "build", "ms_host_timed": min(builds),
                          "chunks": int(indexed.index.shape[0]), "algorithmic_gb": build_bytes / 1e9,
                          "achieved_gbs": build_bytes / min(builds) / 1e6, "frac_of_measured_peak": build_bytes / min(builds) / 1e6 / peak}), flush=True)
        ox = 390000.0
        # x bands of the 10 km footprint: 3 % (two strips), 25 %, 100 %
        boxes = {"x3pct": ((ox + 4000.0, 0.0, -1e4), (ox + 4300.0, 1e7, 1e4)), "x25pct": ((ox + 2000.0, 0.0, -1e4), (ox + 4500.0, 1e7, 1e4)),
                 "all": ((0.0, 0.0, -1e4), (1e7, 1e7, 1e4))}
        queries = [(n, pcq.BoundsSearcher(*b), R if ext == "las" else 12) for n, b in boxes.items()]
        queries += [("class_6_in_every_4th_strip", pcq.ClassSearcher(6), R if ext == "las" else 1),
                    ("class_19_absent", pcq.ClassSearcher(19), R if ext == "las" else 1)]
        for name, s, read_b in queries:
            for cname, make in (("count", lambda: pcq.CountCollector(ctx)), ("buffer", lambda: pcq.BufferCollector(ctx))):
                if cname == "buffer" and name == "all":
                    continue
                res = {}
                for label, df in (("full", plain), ("indexed", indexed)):
                    c = make()

                    def run():
                        c.reset()
                        s.search_files([df], impl, [c])

                    ms = timed(run)
                    # the same call timed on the host clock: the chunk filter and the segment table are host work
                    ctx.synchronize()
                    t0 = time.perf_counter()
                    run()
                    ctx.synchronize()
                    wall = (time.perf_counter() - t0) * 1e3
                    st = ctx.last_scan_stats
                    res[label] = (ms, wall, c.point_count(), st.points_scanned, st.segments, st.chunks_skipped, st.chunks_total)
                    c.close()
                assert res["full"][2] == res["indexed"][2], (name, cname, res)
                fm, im = res["full"], res["indexed"]
                print(json.dumps({"layout": ext, "format": fmt, "points": N, "query": name, "collector": cname, "matches": fm[2],
                                  "full_ms": fm[0], "indexed_ms": im[0], "full_wall_ms": fm[1], "indexed_wall_ms": im[1],
                                  "speedup": fm[0] / im[0], "speedup_wall": fm[1] / im[1], "points_scanned": im[3], "segments": im[4],
                                  "chunks_skipped": im[5], "chunks_total": im[6],
                                  "full_frac_of_measured_peak": (N * read_b + (fm[2] * 31 if cname == "buffer" else 0)) / fm[0] / 1e6 / peak}), flush=True)
        plain.release()
        indexed.release()

        # ---- the same file as a pinned host image: passes through pcq_search_host_files_indexed ----
        n_host = min(N, 1 << 26)  # 64 M points (1.9 GB as LAS format 1) keep the pinned allocation modest
        if n_host != N:
            del buf
            torch.cuda.empty_cache()
            buf, desc, hdr = S.strips_device(ctx, n_host, 64, layout, fmt)
        body = int(desc.n_points) * R
        img = torch.empty(227 + body, dtype=torch.uint8).pin_memory()
        img[:227] = torch.from_numpy(hdr)
        img[227:].copy_(buf[:body])
        image = (img.numpy(), ext)
        per_pt = {"bounds": R if ext == "las" else 12, "class": R if ext == "las" else 1}
        for name, mk, kind in (("x3pct", lambda: pcq.BoundsSearcher(*boxes["x3pct"]), "bounds"), ("x25pct", lambda: pcq.BoundsSearcher(*boxes["x25pct"]), "bounds"),
                               ("class_6_in_every_4th_strip", lambda: pcq.ClassSearcher(6), "class")):
            ix = pcq.HostIndex(ctx)
            rows = {}
            # (the warm-up pass allocates the context's staging ring the first time round)
            for label, index in (("warm_up", None), ("no_index", None), ("first_pass_builds", ix), ("indexed", ix), ("indexed_again", ix)):
                c = pcq.CountCollector(ctx)
                ctx.synchronize()
                t0 = time.perf_counter()
                pcq.search_host_files_multi([image], [mk()], [[c]], index=index)
                cnt = c.point_count()
                wall = (time.perf_counter() - t0) * 1e3
                st = ctx.last_scan_stats
                rows[label] = {"ms_wall": wall, "matches": cnt, "points_over_pcie": int(st.points_scanned),
                               "h2d_gb": st.points_scanned * per_pt[kind] / 1e9, "chunks_skipped": int(st.chunks_skipped)}
                c.close()
            assert len({r["matches"] for r in rows.values()}) == 1, rows
            del rows["warm_up"]
            print(json.dumps({"layout": ext, "format": fmt, "points": int(desc.n_points), "host_staged": True, "query": name, "passes": rows,
                              "speedup_wall": rows["no_index"]["ms_wall"] / rows["indexed_again"]["ms_wall"]}), flush=True)
            ix.close()
        del buf, img
        torch.cuda.empty_cache()


def run_configs(pcq, ctx, stream, timed, peak, args):
    """BASELINE.json configs[2] (C3, ca13-shape LAST class queries with compacted output) and configs[3]
    (C4, navvis-shape dense LAS, bounds + max-density) at their full per-GPU sizes."""
    import time

    import torch

    S, B = pcq.synth, pcq.binding
    impl = pcq.SearchImplementation.Optimized

    def out(d):
        print(json.dumps(d), flush=True)

    # ---- C4: one navvis-shape file, 56.2 M points, format 3, S/L/XL + density 0.1 ----
    for fma in ((False, True) if args.only != "c3" else ()):
        sp = S.navvis_spec(n_points=56_200_000, fma_sensitive=fma)
        buf = torch.empty(sp.n_points * sp.record_len + 256, dtype=torch.uint8, device="cuda:0")
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        df = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
        for name, (qmin, qmax) in (("S", S.NAVVIS_S), ("L", S.NAVVIS_L), ("XL", S.NAVVIS_XL)):
            s = pcq.BoundsSearcher(qmin, qmax)
            cc = pcq.CountCollector(ctx)
            ms_count = timed(lambda: s.search_files([df], impl, [cc]))
            cc.reset()
            s.search_files([df], impl, [cc])
            matches = cc.point_count()
            g = pcq.GridSampledCollector(qmin, qmax, S.NAVVIS_DENSITY, ctx=ctx)

            def run():
                g.reset()
                s.search_files([df], impl, [g])

            ms_grid = timed(run, reps=5)
            fin = []
            for _ in range(3):  # first finalisation allocates its scratch; report the steady state
                run()
                ctx.synchronize()
                t0 = time.perf_counter()
                cells = g.point_count()
                fin.append((time.perf_counter() - t0) * 1e3)
            fin_ms = min(fin)
            bc = pcq.BufferCollector(ctx)

            def runb():
                bc.reset()
                s.search_files([df], impl, [bc])

            ms_sel = timed(runb, reps=5)
            nb = sp.n_points * sp.record_len
            out({"config": "C4 navvis", "fma_sensitive_header": fma, "box": name, "points": sp.n_points, "matches": matches, "cells": cells,
                 "count_ms": ms_count, "count_gbs": nb / ms_count / 1e6, "select_ms": ms_sel,
                 "select_gbs": (nb + 31 * matches) / ms_sel / 1e6, "density_insert_ms": ms_grid,
                 "density_insert_gbs": (nb + 8 * matches) / ms_grid / 1e6, "density_finalize_ms_host_timed": fin_ms,
                 "density_total_gpoints_per_s": sp.n_points / (ms_grid + fin_ms) / 1e6, "peak_gbs": peak})
            g.close()
            bc.close()
        df.release()
        del buf
        torch.cuda.empty_cache()

    if args.only == "c4":
        return
    # ---- C3: ca13-shape LAST, 64 files x 40.75 M points (only the three columns the path reads are resident) ----
    specs = S.ca13_specs()
    dfs, keep = [], []
    for sp in specs:
        # generate the full transposed record block, then keep only position / class columns (19 of 28 bytes are never read)
        buf = torch.empty(sp.n_points * sp.record_len + 256, dtype=torch.uint8, device="cuda:0")
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        host_needed = None
        dfs.append(pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf))
        keep.append(buf)
    total = sum(sp.n_points for sp in specs)
    for klass in (2, 6, 19):
        s = pcq.ClassSearcher(klass)
        cols = [pcq.CountCollector(ctx) for _ in dfs]
        ms_count = timed(lambda: s.search_files(dfs, impl, cols))
        for c in cols:
            c.reset()
        s.search_files(dfs, impl, cols)
        matches = sum(c.point_count() for c in cols)
        bcols = [pcq.BufferCollector(ctx) for _ in dfs]

        def runb():
            for c in bcols:
                c.reset()
            s.search_files(dfs, impl, bcols)

        ms_sel = timed(runb, reps=5)
        assert sum(c.point_count() for c in bcols) == matches
        out({"config": "C3 ca13 LAST", "class": klass, "points": total, "matches": matches, "count_ms": ms_count,
             "count_gbs": total / ms_count / 1e6, "count_gpoints_per_s": total / ms_count / 1e6, "select_ms": ms_sel,
             "select_gbs": (total + matches * (12 + 31)) / ms_sel / 1e6, "select_gpoints_per_s": total / ms_sel / 1e6, "peak_gbs": peak})
        for c in bcols:
            c.close()
    for (name, box) in (("S", S.CA13_S), ("L", S.CA13_L), ("XL", S.CA13_XL)):
        s = pcq.BoundsSearcher(*box)
        cols = [pcq.CountCollector(ctx) for _ in dfs]
        ms_count = timed(lambda: s.search_files(dfs, impl, cols))
        for c in cols:
            c.reset()
        s.search_files(dfs, impl, cols)
        per = [c.point_count() for c in cols]
        scanned = sum(sp.n_points for sp, df in zip(specs, dfs))  # upper bound; tiles skipped by the header test cost nothing
        out({"config": "C3 ca13 LAST", "box": name, "points": total, "matches": sum(per), "count_ms": ms_count,
             "count_gpoints_per_s_dataset": total / ms_count / 1e6, "peak_gbs": peak})


if __name__ == "__main__":
    main()
