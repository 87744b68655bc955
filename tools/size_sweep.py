#!/usr/bin/env python
"""the count scan over N in {1 M ... 4 B} x {LAS 0, 1, 3, LAST} x {device-resident, host-staged}.

    python tools/size_sweep.py > gpurun_out/size_sweep.jsonl

Uniform points, a box with ~50 % selectivity, CountCollector.  Device-resident: CUDA events around the call, median of
--reps after warm-up.  Host-staged (N <= --host-max): the same file as a pinned host image through
pcq_search_host_files, wall-clock including the read-back of the count.  Cells that do not fit the GPU's HBM are
reported as n/a (capacity).  One JSON line per cell."""
import argparse
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,4,16,64,256,1024,2048,4096", help="millions of points")
    ap.add_argument("--cases", default="las:0,las:1,las:3,last:1")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--host-max", type=int, default=256, help="largest N (millions) that is also run host-staged")
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
    box = ((0, 0, 0), (10000.0, 10000.0, 5000.0))

    for case in args.cases.split(","):
        ext, fmt = case.split(":")[0], int(case.split(":")[1])
        layout = B.LAYOUT_LAS if ext == "las" else B.LAYOUT_LAST
        for m in [int(v) for v in args.sizes.split(",")]:
            N = m * 1_000_000
            sp = S.uniform_spec(N, layout, fmt, extent=1_000_000)
            R = sp.record_len
            line = {"layout": ext, "format": fmt, "record_len": R, "points": N}
            free_b, _ = torch.cuda.mem_get_info()
            if N * R + (1 << 30) > free_b:
                line["resident"] = "n/a (capacity)"
                print(json.dumps(line), flush=True)
                continue
            buf = torch.empty(N * R + 256, dtype=torch.uint8, device="cuda:0")
            mm, desc = S.device_points(ctx, sp, buf.data_ptr())
            df = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
            s = pcq.BoundsSearcher(*box)
            c = pcq.CountCollector(ctx)

            def run():
                c.reset()
                s.search_files([df], impl, [c])

            run()
            run()
            ts = []
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ctx.synchronize()
                e0.record(stream)
                run()
                e1.record(stream)
                ctx.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            read_b = R if ext == "las" else 12
            line.update(matches=c.point_count(), resident_ms=ms, resident_gpoints_per_s=N / ms / 1e6,
                        resident_gbs=N * read_b / ms / 1e6, resident_frac_of_measured_peak=N * read_b / ms / 1e6 / peak)
            if m <= args.host_max:
                img = torch.empty(227 + N * R, dtype=torch.uint8).pin_memory()
                img[:227] = torch.from_numpy(S.header_bytes(sp, mm))
                img[227:].copy_(buf[: N * R])
                image = (img.numpy(), ext)
                hc = pcq.CountCollector(ctx)
                walls = []
                for _ in range(4):
                    hc.reset()
                    ctx.synchronize()
                    t0 = time.perf_counter()
                    s.search_host_files([image], [hc])
                    got = hc.point_count()
                    walls.append((time.perf_counter() - t0) * 1e3)
                assert got == line["matches"], (got, line["matches"])
                w = min(walls[1:])
                line.update(host_staged_ms=w, host_staged_gpoints_per_s=N / w / 1e6, host_staged_h2d_gbs=N * read_b / w / 1e6)
                del img
            print(json.dumps(line), flush=True)
            df.release()
            del df, buf
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
