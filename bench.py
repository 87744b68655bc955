#!/usr/bin/env python
"""bench.py — the headline benchmark of the full-scan query path (BASELINE.json).

Workload (configs[1]): doc-shape LAS dataset, 64 format-1 tiles x 31.25 M points
(2 B points, 56 GB) resident in ONE B200's HBM; a step = the S, L and XL bounding-box count queries of
run_query_experiments.rs:117-128 in `--optimized --parallel` form (one CountCollector per file, counts
summed on the host).  Files the query box does not touch are skipped by the header test exactly as the
reference does (las.rs:82-84), so a step scans 5 + 30 + 64 = 99 files = 3.09 G points.

  value  Gpoints/s scanned, inputs resident in HBM, timed with CUDA events on the launch stream
  e2e    same metric through pcq_search_host_files: pinned host file images -> H2D -> scan -> counts D2H
  roofline  the scan kernel's algorithmic bytes (28 B per scanned point) / its event-timed duration
  cpu_baseline  the oracle's run_search_parallel (one thread per file) on this box's host cores

N > 1 (torchrun, one rank per GPU): every rank holds its own 64-tile dataset (weak scaling); counts
need no collective (host sum); the barrier + max-over-ranks timing go through torch.distributed/NCCL.
`--impl reference` times the CPU path alone (the reference is Rust and cannot be built here, so this
is the oracle port), rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "bbox_count_scan_throughput"
UNIT = "Gpoints/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files", type=int, default=64, help="tiles per GPU (square number)")
    ap.add_argument("--pts-per-file", type=int, default=31_250_000)
    ap.add_argument("--variant", type=int, default=0, help="scan kernel variant: 0 auto, 1 direct, 2 staged")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-pts-per-file", type=int, default=2_000_000, help="CPU sample: points per tile")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_name(files, ppf):
    return (f"C2 doc-shape LAS fmt1: {files} tiles x {ppf / 1e6:.2f} M pts per GPU, S+L+XL bbox count, "
            f"--optimized --parallel semantics")


def queries(S):
    return [("S", S.DOC_S), ("L", S.DOC_L), ("XL", S.DOC_XL)]


def tiles_hit(specs, box) -> list:
    """indices of tiles whose (true) bounds can intersect the box — for bookkeeping only"""
    qmin, qmax = box
    out = []
    for k, sp in enumerate(specs):
        lo = [sp.offset[a] + sp.lo[a] * sp.scale[a] for a in range(3)]
        hi = [sp.offset[a] + sp.hi[a] * sp.scale[a] for a in range(3)]
        if all(lo[a] <= qmax[a] and hi[a] >= qmin[a] for a in range(3)):
            out.append(k)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median of the samples under load (the upper half: idle samples before/after the region drag it down)
        top = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(pcq, files_n, ppf, steps, warmup, threads):
    """The reference's CPU path on host cores: oracle run_search_parallel (count) on a bounded sample of the
    workload: the same 64-tile layout at `ppf` points per tile, S+L+XL.  -> (Gpts/s, ms/step, sample text)"""
    import numpy as np

    from oracle import oracle as orc

    S = pcq.synth
    specs = S.doc_specs(n_files=files_n, pts_per_file=ppf)
    images = [None] * len(specs)

    def gen(lo, hi):
        for k in range(lo, hi):
            images[k] = S.host_file(specs[k])

    nthr = max(1, min(threads, len(specs)))
    chunks = [(i * len(specs) // nthr, (i + 1) * len(specs) // nthr) for i in range(nthr)]
    ths = [threading.Thread(target=gen, args=c) for c in chunks]
    [t.start() for t in ths]
    [t.join() for t in ths]
    exts = ["las"] * len(specs)
    qs = queries(S)
    scanned = sum(len(tiles_hit(specs, box)) for _, box in qs) * ppf

    def step():
        tot = 0
        for _, box in qs:
            tot += int(orc.count_parallel(images, exts, threads, bounds=box).sum())
        return tot

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        matches = step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    sample = (f"{len(specs)} tiles x {ppf / 1e6:.2f} M pts (same tile layout, {100.0 * ppf / 31_250_000:.1f} % of each "
              f"file's points), S+L+XL, {threads} threads = min(files, cores), warm page cache")
    return scanned / dt / 1e9, dt * 1e3, sample, matches


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = host_cores()

    from pcq_import import pcq

    S, B = pcq.synth, pcq.binding
    config = {"workload": workload_name(args.files, args.pts_per_file), "queries": ["doc-S", "doc-L", "doc-XL"],
              "collector": "CountCollector per file (run_search_parallel)", "l2": "inputs (tens of GB per query) far larger than the 126 MB L2; no flush needed",
              "parallelism": f"{world} x 1 GPU, files sharded by rank, host-summed counts"}

    # ---------------- reference arm: the CPU path alone ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        val, ms, sample, _ = cpu_reference_run(pcq, args.files, args.cpu_pts_per_file, args.steps, args.warmup, cores)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "i32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- B200 arm ----------------
    import numpy as np
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = f"cuda:{local_rank}"
    ctx = pcq.Context(local_rank)
    if args.variant:
        ctx.set_scan_variant(args.variant)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # dataset: generated on the device by the counter-based generator (rank-specific seed)
    specs = S.doc_specs(n_files=args.files, pts_per_file=args.pts_per_file, seed=0x5EED0000 + 2000 + 100_000 * rank)
    R = specs[0].record_len
    bufs, dfiles, descs, minmaxes = [], [], [], []
    for sp in specs:
        buf = torch.empty(sp.n_points * R + 256, dtype=torch.uint8, device=dev)
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        bufs.append(buf)
        minmaxes.append(mm)
        descs.append(desc)
        dfiles.append(pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf))
    qs = queries(S)
    searchers = [pcq.BoundsSearcher(*box) for _, box in qs]
    impl = pcq.SearchImplementation.Optimized
    collectors = [[pcq.CountCollector(ctx) for _ in specs] for _ in qs]
    hit_files = [tiles_hit(specs, box) for _, box in qs]
    scanned_pts = sum(len(h) for h in hit_files) * args.pts_per_file
    alg_bytes_per_step = scanned_pts * R

    def reset_collectors():
        for cs in collectors:
            for c in cs:
                c.reset()

    def device_step(events=None):
        for qi, s in enumerate(searchers):
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            s.search_files(dfiles, impl, collectors[qi])
            if events is not None:
                e1.record(stream)
                events.append((e0, e1))

    def read_counts():
        return [sum(c.point_count() for c in cs) for cs in collectors]

    # ---- value: device-resident, K steps between barriers, CUDA events on the launch stream ----
    for _ in range(args.warmup):
        device_step()
    ctx.synchronize()
    reset_collectors()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    kernel_events = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        device_step(kernel_events)
    ev1.record(stream)
    ctx.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    dt_ms = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = dt_ms / args.steps
    total_scanned = sum_over_ranks(float(scanned_pts))
    value = total_scanned / (ms_per_step * 1e-3) / 1e9
    counts = read_counts()
    per_step_counts = [c // args.steps for c in counts]
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_events) / len(kernel_events)
    bytes_per_launch = alg_bytes_per_step / len(qs)
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tr = ROOT / "profiles" / "traffic.json"
    if tr.exists():
        try:
            # DRAM bytes per launch = this launch size x the dram/algorithmic ratio of the committed `ncu --set full`
            # capture of the same kernel (profiles/traffic.json; captured on a 16 x 8 M-point dataset)
            ratio = json.loads(tr.read_text()).get("scan_count_dram_to_algorithmic_ratio")
            traffic = None if ratio is None else float(ratio) * bytes_per_launch
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "k_scan_staged<28,COUNT> (k_scan_direct when --variant 1)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": kernel_ms,
                "note": "average over the S, L and XL launches of the timed region (5 / 30 / 64 tiles of 875 MB); traffic = "
                        "algorithmic bytes x the dram/algorithmic ratio of the committed ncu capture (profiles/traffic.json)"}

    # ---- e2e: pinned host file images -> H2D -> scan -> counts to the host, every step ----
    e2e = None
    host_imgs = None
    if not args.no_e2e:
        import psutil

        file_bytes = 227 + args.pts_per_file * R
        avail = psutil.virtual_memory().available
        n_host = min(len(specs), max(1, int(avail * 0.45 / world) // file_bytes))
        host_imgs, host_keep = [], []
        for k in range(n_host):
            # pinned by torch's host allocator: the driver knows the pages, so libpcq's H2D copies run at link speed
            img = torch.empty(file_bytes, dtype=torch.uint8, pin_memory=True)
            img[:227] = torch.from_numpy(S.header_bytes(specs[k], minmaxes[k]))
            img[227:].copy_(bufs[k][: file_bytes - 227])  # D2H of the device-generated points
            host_keep.append(img)
            host_imgs.append((img.numpy(), "las"))
        torch.cuda.synchronize()
        sub_specs = specs[:n_host]
        e_hit = [[k for k in h if k < n_host] for h in hit_files]
        e_pts = sum(len(h) for h in e_hit) * args.pts_per_file
        e_cols = [[pcq.CountCollector(ctx) for _ in range(n_host)] for _ in qs]
        e_steps, e_warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))

        def e2e_step():
            # ONE plugin call per step: the step's three queries share one host-staged pass, so every file image
            # crosses PCIe once per step; then the per-file counts come back to the host
            for cs in e_cols:
                for c in cs:
                    c.reset()
            pcq.search_host_files_multi(host_imgs, searchers, e_cols)
            return [sum(c.point_count() for c in cs) for cs in e_cols]

        def e2e_step_per_query():
            out = []
            for qi, s in enumerate(searchers):
                for c in e_cols[qi]:
                    c.reset()
                s.search_host_files(host_imgs, e_cols[qi])
                out.append(sum(c.point_count() for c in e_cols[qi]))
            return out

        for _ in range(e_warm):
            e_counts = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e_counts = e2e_step()
        torch.cuda.synchronize()
        e_dt = max_over_ranks((time.perf_counter() - t0) / e_steps)
        barrier()
        e2e_step_per_query()
        t0 = time.perf_counter()
        pq_counts = e2e_step_per_query()
        torch.cuda.synchronize()
        pq_dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert pq_counts == e_counts
        files_any = sorted(set(k for h in e_hit for k in h))
        if n_host == len(specs):
            assert e_counts == per_step_counts, f"e2e counts {e_counts} != resident counts {per_step_counts}"
        e2e = {"value": sum_over_ranks(float(e_pts)) / e_dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": len(files_any) * args.pts_per_file * R, "d2h_bytes_per_step": 8 * n_host * len(qs),
               "ms_per_step": e_dt * 1e3, "steps": e_steps, "warmup": e_warm, "files_in_host_memory": n_host,
               "api": "pcq_search_host_files_multi: one call per step; pinned file images stream through HBM chunk buffers once "
                      "(H2D overlapped with the scans), each chunk is scanned by S, L and XL, per-file counts D2H",
               "per_query_staging": {"value": sum_over_ranks(float(e_pts)) / pq_dt / 1e9, "unit": UNIT,
                                     "h2d_bytes_per_step": e_pts * R, "ms_per_step": pq_dt * 1e3,
                                     "api": "pcq_search_host_files once per query (every query re-stages the files it touches)"}}

    # ---- cpu_baseline: rank 0, bounded sample, same run, same box ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        v, ms, sample, _ = cpu_reference_run(pcq, args.files, min(args.cpu_pts_per_file, args.pts_per_file), 3, 1, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms}

    if rank == 0:
        config.update({"points_scanned_per_step_per_gpu": scanned_pts, "dataset_points_per_gpu": args.files * args.pts_per_file,
                       "tiles_scanned": {n: len(h) for (n, _), h in zip(qs, hit_files)},
                       "matches_per_step": dict(zip([n for n, _ in qs], per_step_counts)), "record_len": R,
                       "scan_variant": {0: "auto (staged)", 1: "direct", 2: "staged"}[args.variant]})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32",
                "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
