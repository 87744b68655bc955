#!/usr/bin/env python
"""bench.py — the headline benchmark of the full-scan query path (BASELINE.json).

Workload (configs[1]): the doc-shape LAS dataset — 64 format-1 tiles x 31.25 M points, 2 B points,
56 GB — and the S, L and XL bounding-box count queries of run_query_experiments.rs:117-128 in `--optimized --parallel`
form (one CountCollector per file, counts summed).  Files the query box does not touch are skipped by the header test
exactly as the reference does (las.rs:82-84), so a step scans 5 + 30 + 64 = 99 tile-scans = 3.09 G points.

STRONG scaling: the dataset is the same 2 B points at every N.  Every tile is cut into N chunk-aligned point ranges
and GPU r holds range r of every tile (pcq_shard_plan, PCQ_SHARD_RANGES), so that also the S box — 5 of 64 tiles —
spreads over all GPUs.  One process per GPU (torchrun); the multi-GPU layer is the library's own (pcq_group_*,
csrc/group.cu: its own NCCL communicator, per-file counts reduced with one ncclAllReduce per step); torch.distributed
only carries the barrier and the max-over-ranks of the timings.

  value     Gpoints/s scanned by the whole job, inputs resident in HBM, CUDA events on the launch stream, max over ranks
  e2e       same metric through pcq_group_search_host_files: pinned host file images -> H2D -> scan -> counts D2H
  roofline  the scan kernel's algorithmic bytes (28 B per scanned point) / its event-timed duration (rank 0)
  cpu_baseline   (N = 1) the oracle's run_search_parallel on this box's host cores over the SAME file images the e2e leg
            streams; its per-file counts must equal the GPU's before anything is printed
  density   BASELINE configs[3] (C4): one navvis-shape file, range-sharded over the N GPUs, bounds + --density 0.1 with
            the candidate-cell all-to-all; phases timed, result compared with the oracle in the run
  extra     (N = 1) driver-visible numbers for C1 / C3 / C4 / C5, each with a roofline recomputable from bytes and ms

`--impl reference` times the CPU path alone (the reference is Rust and cannot be built here, so this is the oracle
port `oracle/liboracle.so`), rank 0 only, on the same 64-tile dataset generated on the host by libpcq_synth.so (the
product library is not mapped), at full size when the box's RAM allows.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import mmap
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
SEED = 0x5EED0000 + 2000
GOLDEN_COUNTS = ROOT / "tests" / "golden" / "c2_counts.json"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files", type=int, default=64, help="tiles of the dataset (square number)")
    ap.add_argument("--pts-per-file", type=int, default=31_250_000)
    ap.add_argument("--variant", type=int, default=0, help="scan kernel variant: 0 auto, 1 direct, 2 staged")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-density", action="store_true")
    ap.add_argument("--write-golden", action="store_true", help="record the oracle-checked per-query counts (N = 1)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_name(files, ppf):
    return (f"C2 doc-shape LAS fmt1: {files} tiles x {ppf / 1e6:.2f} M pts (one dataset, range-sharded over the GPUs), "
            f"S+L+XL bbox count, --optimized --parallel semantics")


def queries(S):
    return [("S", S.DOC_S), ("L", S.DOC_L), ("XL", S.DOC_XL)]


def spec_box(sp):
    """what the header of a generated tile will say, to within the generator's true min / max"""
    lo = [sp.offset[a] + sp.lo[a] * sp.scale[a] for a in range(3)]
    hi = [sp.offset[a] + sp.hi[a] * sp.scale[a] for a in range(3)]
    return lo, hi


def tiles_hit(specs, box) -> list:
    """indices of tiles whose bounds can intersect the box — for bookkeeping only"""
    qmin, qmax = box
    out = []
    for k, sp in enumerate(specs):
        lo, hi = spec_box(sp)
        if all(lo[a] <= qmax[a] and hi[a] >= qmin[a] for a in range(3)):
            out.append(k)
    return out


class ClockSampler:
    """SM clock and throttle reasons of one GPU while it is under load (B200_PROFILING.md's clocks line), sampled through
    NVML every 5 ms from a thread (nvidia-smi's shortest period, 100 ms, misses the timed region of an 8-GPU run, which
    lasts tens of milliseconds); falls back to `nvidia-smi -lms 100`."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []  # (sm MHz, reasons bit mask)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.lines = []

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].strip().isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            # the first call of each query refreshes NVML's cached state and takes 7-26 ms (measured, tools/nvml_cost.py;
            # later calls ~1 us): made here, before the thread and the timed region start
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons"):
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            else:
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons")
                                             else int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
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
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            sm = sorted(s for s, _ in self.samples)
            mask = 0
            for _, r in self.samples:
                mask |= r
            reasons = [name for name, bit in self.REASONS if mask & bit]
            top = sm[len(sm) // 2:]  # the upper half: idle samples at the edges of the window drag the median down
            return {"sm_mhz": statistics.median(top), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm), "via": "nvml, 5 ms"}
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
        top = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "via": "nvidia-smi, 100 ms"}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def golden_key(files, ppf):
    return f"files={files},pts_per_file={ppf},seed={SEED}"


# ---------------------------------------------------------------------------------------------------
# the reference arm: the CPU path alone
# ---------------------------------------------------------------------------------------------------
def reference_arm(args, cores, config):
    import numpy as np
    import psutil

    from oracle import oracle as orc

    # the generator is a library of its own: this process maps oracle/liboracle.so and libpcq_synth.so, never libpcq.so
    S = load_synth_only()
    R = 28
    full = args.pts_per_file
    avail = psutil.virtual_memory().available
    ppf = full
    while args.files * (227 + ppf * R) > avail * 0.6 and ppf > 1_000_000:
        ppf //= 2
    specs = S.doc_specs(n_files=args.files, pts_per_file=ppf, seed=SEED)
    images = [None] * len(specs)

    def gen(lo, hi):
        for k in range(lo, hi):
            img = np.empty(227 + ppf * R, dtype=np.uint8)
            S.host_file_into(specs[k], img)
            images[k] = img

    nthr = max(1, min(cores, len(specs)))
    chunks = [(i * len(specs) // nthr, (i + 1) * len(specs) // nthr) for i in range(nthr)]
    t_gen = time.perf_counter()
    ths = [threading.Thread(target=gen, args=c) for c in chunks]
    [t.start() for t in ths]
    [t.join() for t in ths]
    t_gen = time.perf_counter() - t_gen
    exts = ["las"] * len(specs)
    qs = queries(S)
    scanned = sum(len(tiles_hit(specs, box)) for _, box in qs) * ppf

    def step():
        return [orc.count_parallel(images, exts, cores, bounds=box) for _, box in qs]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        per_file = step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = scanned / dt / 1e9
    sample = (f"{len(specs)} tiles x {ppf / 1e6:.2f} M pts = {100.0 * ppf / full:.0f} % of the workload (generated on the host in "
              f"{t_gen:.1f} s by libpcq_synth.so), S+L+XL per step, {cores} threads = min(files, cores), images resident in RAM")
    totals = {n: int(c.sum()) for (n, _), c in zip(qs, per_file)}
    config = dict(config, matches_per_step=totals, same_dataset_as_b200_arm=(ppf == full))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "i32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


_SYNTH_ONLY = None


def load_synth_only():
    """The synthetic-data half of the package without its product half: binding.py dlopens libpcq.so on import, which the
    reference arm must not map, so the few definitions it needs are rebuilt here on libpcq_synth.so alone."""
    import types

    global _SYNTH_ONLY
    if _SYNTH_ONLY is not None:
        return _SYNTH_ONLY
    so = C.CDLL(str(ROOT / "adhoc-queries-pointclouds_b200" / "libpcq_synth.so"))
    _D3 = C.c_double * 3

    class SynthSpec(C.Structure):
        _fields_ = [("seed", C.c_uint64), ("n_points", C.c_uint64), ("layout", C.c_uint8), ("format", C.c_uint8), ("shape", C.c_uint8),
                    ("n_classes", C.c_uint8), ("record_len", C.c_uint16), ("flag_per_64k", C.c_uint16), ("lo", C.c_int32 * 3),
                    ("hi", C.c_int32 * 3), ("scale", _D3), ("offset", _D3), ("class_val", C.c_uint8 * 8), ("class_cum", C.c_uint16 * 8)]

    so.pcq_synth_host.restype = C.c_int
    so.pcq_synth_host.argtypes = [C.POINTER(SynthSpec), C.c_void_p, C.c_size_t]
    so.pcq_synth_last_error.restype = C.c_char_p
    DOC_CLASSES = ((1, 0.25), (2, 0.45), (5, 0.15), (6, 0.10), (7, 0.02), (9, 0.03))
    m = types.SimpleNamespace()
    m.DOC_S = ((390000.0, 130000.0, 0.0), (390500.0, 140000.0, 200.0))
    m.DOC_L = ((390000.0, 130000.0, 0.0), (400000.0, 140000.0, 200.0))
    m.DOC_XL = ((389400.0, 124200.0, -94.88), (406200.0, 148200.0, 760.03))

    def doc_specs(n_files, pts_per_file, seed):
        # the same arithmetic as synth.doc_specs (kept in step by tests/test_bench_contract.py)
        g = max(1, int(round(n_files ** 0.5)))
        assert g * g == n_files
        (x0, y0, _), (x1, y1, _) = m.DOC_XL
        w, h = (x1 - x0) / g, (y1 - y0) / g
        out = []
        for j in range(g):
            for i in range(g):
                sp = SynthSpec()
                sp.seed = (seed + j * g + i) & 0xFFFFFFFFFFFFFFFF
                sp.n_points = int(pts_per_file)
                sp.layout, sp.format, sp.shape, sp.n_classes, sp.record_len, sp.flag_per_64k = 0, 1, 1, len(DOC_CLASSES), 28, 655
                hi_x, hi_y = int(round(w / 0.01)) - 1, int(round(h / 0.01)) - 1
                lo_z, hi_z = int(round(-94.88 / 0.01)), int(round(194.88 / 0.01))
                for a, (lo, hi, off) in enumerate(((0, hi_x, x0 + i * w), (0, hi_y, y0 + j * h), (lo_z, hi_z, 0.0))):
                    sp.lo[a], sp.hi[a], sp.scale[a], sp.offset[a] = lo, hi, 0.01, off
                cum = 0.0
                for k, (val, p) in enumerate(DOC_CLASSES):
                    cum += p
                    sp.class_val[k] = val
                    sp.class_cum[k] = 65535 if k == len(DOC_CLASSES) - 1 else min(65535, int(round(cum * 65536)) - 1)
                out.append(sp)
        return out

    def host_file_into(spec, out):
        rc = so.pcq_synth_host(C.byref(spec), C.c_void_p(out.ctypes.data), out.nbytes)
        if rc != 0:
            raise RuntimeError(so.pcq_synth_last_error().decode())

    m.doc_specs, m.host_file_into, m.SynthSpec = doc_specs, host_file_into, SynthSpec
    _SYNTH_ONLY = m
    return m


# ---------------------------------------------------------------------------------------------------
# B200 arm helpers
# ---------------------------------------------------------------------------------------------------
PAGE = 4096


class HostImage:
    """A whole-file image in host memory of which only the header and ONE point range are populated and pinned
    (pcq_host_register): what a rank of a sharded scan has to hold of a file."""

    def __init__(self, pcq, file_bytes: int, header, first_byte: int, n_bytes: int, device_src):
        import numpy as np
        import torch

        self.pcq = pcq
        self.map = mmap.mmap(-1, file_bytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)  # untouched pages cost nothing
        try:  # 2 MB pages where the kernel hands them out: fewer IOMMU / TLB entries under the copy engines
            self.map.madvise(mmap.MADV_HUGEPAGE)
        except (AttributeError, OSError):
            pass
        self.np = np.frombuffer(self.map, dtype=np.uint8)
        self.addr = self.np.ctypes.data
        self.nbytes = file_bytes
        self.np[:227] = header
        lo = (227 + first_byte) // PAGE * PAGE
        hi = min(file_bytes, (227 + first_byte + n_bytes + PAGE - 1) // PAGE * PAGE)
        self.reg = None
        if n_bytes:
            self.np[lo:hi:PAGE] = 0  # fault the pages in (first touch: on this rank's NUMA node)
            self.np[:227] = header
            pcq.binding.check(pcq.lib.pcq_host_register(C.c_void_p(self.addr + lo), hi - lo))
            self.reg = self.addr + lo
            t = torch.from_numpy(self.np)
            t[227 + first_byte: 227 + first_byte + n_bytes].copy_(device_src[:n_bytes])
            self.np[:227] = header

    def close(self):
        if self.reg is not None:
            self.pcq.lib.pcq_host_unregister(C.c_void_p(self.reg))
            self.reg = None
        self.np = None
        try:
            self.map.close()
        except BufferError:
            pass


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = host_cores()
    config = {"workload": workload_name(args.files, args.pts_per_file), "queries": ["doc-S", "doc-L", "doc-XL"],
              "collector": "CountCollector per file (run_search_parallel)",
              "l2": "inputs (GB per query and GPU) far larger than the 126 MB L2; no flush needed",
              "parallelism": f"{world} x 1 GPU; every tile cut into {world} chunk-aligned point ranges, GPU r holds range r of every tile "
                             f"(pcq_shard_plan); per-file counts reduced by the library (ncclAllReduce of {args.files} integers per query)"}

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, cores, config)
        return

    # ---------------- B200 arm ----------------
    import numpy as np
    import torch
    import torch.distributed as dist

    from pcq_import import pcq

    S, B = pcq.synth, pcq.binding
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    group = pcq.Group.from_torch_distributed(local_rank) if world > 1 else pcq.Group.local(1, [local_rank])
    ctx = group.ctx(0)
    node = C.c_int(-1)
    B.check(pcq.lib.pcq_ctx_bind_host_thread(ctx.handle, C.byref(node)))  # pinned staging memory next to the GPU's PCIe root
    if args.variant:
        ctx.set_scan_variant(args.variant)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    peak, peak_src = hbm_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM) if world > 1 else x

    # ---- the dataset: this rank's point ranges, generated straight into HBM by the counter-based generator ----
    specs = S.doc_specs(n_files=args.files, pts_per_file=args.pts_per_file, seed=SEED)
    R = specs[0].record_len
    ppf = [sp.n_points for sp in specs]
    plan = pcq.shard_plan(ppf, world)
    mine = plan[plan["rank"] == rank]
    bufs, firsts, counts_n = {}, {}, {}
    mm_all = torch.empty((args.files, 6), dtype=torch.int64)
    mm_all[:, :3] = 2**31 - 1
    mm_all[:, 3:] = -(2**31)
    for sh in mine:
        f, first, n = int(sh["file"]), int(sh["first_point"]), int(sh["n_points"])
        buf = torch.empty(n * R + 256, dtype=torch.uint8, device=dev)
        mm = S.device_point_range(local_rank, specs[f], first, n, buf.data_ptr())
        bufs[f], firsts[f], counts_n[f] = buf, first, n
        mm_all[f] = torch.tensor(mm, dtype=torch.int64)
    if world > 1:  # a tile's header holds the min / max of ALL its points
        lo, hi = mm_all[:, :3].to(dev), mm_all[:, 3:].to(dev)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        mm_all = torch.cat([lo, hi], dim=1).cpu()
    minmax = [[int(v) for v in mm_all[f]] for f in range(args.files)]
    descs = [S.desc_of(specs[f], minmax[f]) for f in range(args.files)]
    pieces = []
    for f in sorted(bufs):
        d = B.FileDesc.from_buffer_copy(descs[f])
        d.n_points = counts_n[f]
        pieces.append((pcq.DeviceFile.wrap(ctx, d, bufs[f].data_ptr(), firsts[f], keepalive=bufs[f]), 0, f))
    ds = group.wrap_files(ppf, pieces)
    qs = queries(S)
    searchers = [pcq.BoundsSearcher(*box) for _, box in qs]
    hit_files = [tiles_hit(specs, box) for _, box in qs]
    scanned_pts = sum(len(h) for h in hit_files) * args.pts_per_file            # whole job, per step
    my_scanned = [sum(counts_n.get(f, 0) for f in h) for h in hit_files]        # this rank, per query

    def device_step():
        return group.search(ds, searchers, B.COLLECT_COUNT, True)

    # ---- value: K steps between barriers, CUDA events on the launch stream, max over ranks ----
    for _ in range(args.warmup):
        res = device_step()
    per_file_counts = [r.counts() for r in res]
    sampler = ClockSampler(local_rank)
    sampler.start()  # (primes NVML, then samples every 5 ms; the idle samples before the barrier fall out of the median)
    barrier()
    launches0 = group.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        res = device_step()
    last_counts = [r.counts() for r in res]  # (waits for the last step's counts)
    ev1.record(stream)
    ctx.synchronize()
    barrier()
    launches = int(sum_over_ranks(float(group.launch_count - launches0)))
    dt_ms = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = dt_ms / args.steps
    value = scanned_pts / (ms_per_step * 1e-3) / 1e9
    assert last_counts == per_file_counts, "counts changed between steps"
    totals = {n: sum(c) for (n, _), c in zip(qs, last_counts)}

    # ---- roofline: every launch of a second pass bracketed by events (rank 0's kernels; every rank holds the same share) ----
    kernel_ms = [[] for _ in qs]
    for _ in range(max(3, min(args.steps, 10))):
        for qi, s in enumerate(searchers):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.synchronize()
            e0.record(stream)
            r = group.search(ds, [s], B.COLLECT_COUNT, True)
            e1.record(stream)
            r[0].counts()
            ctx.synchronize()
            kernel_ms[qi].append(e0.elapsed_time(e1))
    clocks = sampler.stop()  # sampled from the start of the timed region to the end of this pass (the GPU is busy throughout)
    avg_ms = [statistics.mean(v) for v in kernel_ms]
    bytes_q = [p * R for p in my_scanned]
    achieved = sum(bytes_q) / (sum(avg_ms) * 1e-3) / 1e9
    traffic = None
    tr = ROOT / "profiles" / "traffic.json"
    if tr.exists():
        try:
            ratio = json.loads(tr.read_text()).get("scan_count_dram_to_algorithmic_ratio")
            traffic = None if ratio is None else float(ratio) * sum(bytes_q) / len(qs)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "k_scan_staged<28,COUNT> (k_scan_direct when --variant 1)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sum(bytes_q) / len(qs), "avg_launch_ms": sum(avg_ms) / len(qs),
                "per_query": {n: {"bytes": b, "ms": m, "gbs": b / m / 1e6} for (n, _), b, m in zip(qs, bytes_q, avg_ms)},
                "note": "rank 0's S, L and XL launches of a second pass, each bracketed by CUDA events on the launch stream (the bracket "
                        "also holds the per-file count gather kernel and, for N > 1, the all-reduce of 64 integers); traffic = algorithmic "
                        "bytes x the dram/algorithmic ratio of the committed ncu capture (profiles/traffic.json)"}

    # ---- in-run parity: the oracle-checked counts of this very configuration ----
    key = golden_key(args.files, args.pts_per_file)
    golden = json.loads(GOLDEN_COUNTS.read_text()) if GOLDEN_COUNTS.exists() else {}
    parity = {"checked_against": None}
    if key in golden:
        assert totals == golden[key]["matches_per_step"], f"counts {totals} != oracle-checked golden counts {golden[key]['matches_per_step']}"
        parity = {"checked_against": "tests/golden/c2_counts.json (written by an N = 1 run whose per-file counts equalled the oracle's)"}

    # ---- e2e: pinned host images -> H2D over every GPU's own PCIe link -> scan -> counts to the host, every step ----
    e2e, images, host_imgs = None, None, []
    if not args.no_e2e:
        import psutil

        need = sum(counts_n.values()) * R
        avail = psutil.virtual_memory().available
        assert need * world < avail * 0.7, f"host images of {need * world / 1e9:.0f} GB do not fit {avail / 1e9:.0f} GB of free RAM"
        images = []
        for f in range(args.files):
            fb = 227 + ppf[f] * R
            hdr = S.header_bytes(specs[f], minmax[f])
            if f in bufs:
                hi = HostImage(pcq, fb, hdr, firsts[f] * R, counts_n[f] * R, bufs[f])
            else:
                hi = HostImage(pcq, fb, hdr, 0, 0, None)
            host_imgs.append(hi)
            images.append(((hi.addr, hi.nbytes), "las"))
        torch.cuda.synchronize()
        e_steps, e_warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))

        def e2e_step():
            # ONE plugin call per step: the step's three queries share one host-staged pass over this rank's ranges, so
            # every byte crosses PCIe once per step; the per-file counts come back reduced over the group
            rs = group.search_host_files(images, searchers, B.COLLECT_COUNT, True)
            return [r.counts() for r in rs]

        for _ in range(e_warm):
            e_counts = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e_counts = e2e_step()
        torch.cuda.synchronize()
        e_dt = max_over_ranks((time.perf_counter() - t0) / e_steps)
        barrier()
        assert e_counts == per_file_counts, "host-staged counts differ from the resident counts"
        files_any = sorted(set(k for h in hit_files for k in h))
        e2e = {"value": scanned_pts / e_dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(sum(ppf[f] for f in files_any)) * R, "d2h_bytes_per_step": 8 * args.files * len(qs) * world,
               "ms_per_step": e_dt * 1e3, "steps": e_steps, "warmup": e_warm, "h2d_gbs_per_gpu": sum(counts_n.get(f, 0) for f in files_any) * R / e_dt / 1e9,
               "numa_node_of_rank0": node.value,
               "api": "pcq_group_search_host_files: one call per step; every rank streams its point ranges of the pinned file images "
                      "through HBM chunk buffers once (H2D overlapped with the scans), each chunk is scanned by S, L and XL, per-file "
                      "counts are reduced over the group and copied to the host"}

    # ---- cpu_baseline (N = 1): the oracle on the SAME images, its per-file counts must equal the GPU's ----
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu and images is not None:
        from oracle import oracle as orc

        views = [hi.np for hi in host_imgs]
        exts = ["las"] * len(views)

        def cpu_step():
            return [orc.count_parallel(views, exts, cores, bounds=box) for _, box in qs]

        want = cpu_step()
        for (n, _), got, w in zip(qs, per_file_counts, want):
            assert got == [int(v) for v in w], f"query {n}: GPU per-file counts differ from the oracle's"
        t0 = time.perf_counter()
        c_steps = 2
        for _ in range(c_steps):
            cpu_step()
        c_dt = (time.perf_counter() - t0) / c_steps
        cpu = {"value": scanned_pts / c_dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": c_dt * 1e3,
               "sample": f"the whole workload: the {len(views)} file images of the e2e leg ({args.pts_per_file / 1e6:.2f} M pts each), S+L+XL, "
                         f"{cores} threads = min(files, cores), {c_steps} steps after 1 warm-up; per-file counts asserted equal to the GPU's"}
        del views
        parity = {"checked_against": "oracle (oracle/liboracle.so run_search_parallel) on the same file images, per file and query, in this run"}
        if args.write_golden:
            golden[key] = {"matches_per_step": totals, "per_file_sha": None}
            GOLDEN_COUNTS.write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")
            (ROOT / "gpurun_out").mkdir(exist_ok=True)
            (ROOT / "gpurun_out" / "c2_counts.json").write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")

    # ---- e2e sub-value (N = 1): the same tiles as LAST files — only the 12-byte position column crosses PCIe ----
    if e2e is not None and world == 1 and not args.no_extra:
        e2e["as_last_columns"] = e2e_last(pcq, group, specs, minmax, bufs, searchers, scanned_pts, hit_files, per_file_counts)

    for hi in host_imgs:
        hi.close()
    host_imgs, images = [], None
    ds.release()
    del pieces, bufs
    torch.cuda.empty_cache()

    # ---- e2e sub-value (N = 1): repeated queries over acquisition-ordered tiles with the on-the-fly chunk index ----
    if e2e is not None and world == 1 and not args.no_extra:
        e2e["strip_ordered_indexed_repeat"] = e2e_indexed_repeat(pcq, ctx, specs, qs, scanned_pts, R)

    # ---- density (C4): one navvis-shape file, range-sharded, bounds + --density 0.1, cell all-to-all ----
    density = None
    if not args.no_density:
        density = density_record(pcq, group, ctx, stream, rank, world, local_rank, dev, max_over_ranks, sum_over_ranks, barrier, peak)

    extra = None
    if world == 1 and not args.no_extra:
        extra = extra_records(pcq, ctx, stream, peak)

    if rank == 0:
        config.update({"points_scanned_per_step": scanned_pts, "dataset_points": sum(ppf),
                       "tiles_scanned": {n: len(h) for (n, _), h in zip(qs, hit_files)}, "matches_per_step": totals, "record_len": R,
                       "scan_variant": {0: "auto (staged)", 1: "direct", 2: "staged"}[args.variant], "parity": parity})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32",
                "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "clocks": clocks, "density": density, "extra": extra}
        print(json.dumps(line))
    group.close()
    if world > 1:
        dist.destroy_process_group()


def e2e_last(pcq, group, specs, minmax, bufs, searchers, scanned_pts, hit_files, want_counts):
    """The C2 tiles as LAST files in host memory (N = 1): header + position column populated, 12 of 28 bytes per point
    cross PCIe."""
    import torch

    S, B = pcq.synth, pcq.binding
    imgs, keep = [], []
    for f, sp in enumerate(specs):
        n, Rl = sp.n_points, sp.record_len
        pos = bufs[f][: n * Rl].view(n, Rl)[:, :12].contiguous().view(-1)  # the positions column of the same points
        hi = HostImage(pcq, 227 + n * Rl, S.header_bytes(sp, minmax[f]), 0, n * 12, pos)
        del pos
        keep.append(hi)
        imgs.append(((hi.addr, hi.nbytes), "last"))
    torch.cuda.synchronize()

    def step():
        return [r.counts() for r in group.search_host_files(imgs, searchers, B.COLLECT_COUNT, True)]

    got = step()
    t0 = time.perf_counter()
    n_steps = 3
    for _ in range(n_steps):
        got = step()
    dt = (time.perf_counter() - t0) / n_steps
    assert got == want_counts, "LAST host-staged counts differ from the LAS resident counts"
    files_any = sorted(set(k for h in hit_files for k in h))
    out = {"value": scanned_pts / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": n_steps,
           "h2d_bytes_per_step": sum(specs[f].n_points for f in files_any) * 12,
           "note": "same points, same queries, same counts; the files are LAST (columnar): a bounds count needs the position column only"}
    for hi in keep:
        hi.close()
    return out


def e2e_indexed_repeat(pcq, ctx, specs, qs, scanned_pts, R):
    """The C2 tiles with their points in ACQUISITION order (every tile = 16 flight strips stored one after the other,
    each covering its own band of x with 10 % overlap), host-staged through pcq_search_host_files_indexed: the first
    pass copies everything and builds the chunk headers as a by-product (improvements.md:3-10), every later pass copies
    only the runs of chunks in which some query of the step can find a match.  Same shape, sizes and queries as the
    headline workload; the points differ (they are generated strip by strip), so the counts are checked against the
    unindexed pass over the same images."""
    import copy

    import torch

    S, B = pcq.synth, pcq.binding
    n_strips = 16
    searchers = [pcq.BoundsSearcher(*box) for _, box in qs]
    dev = f"cuda:{ctx.device}"
    keep, images = [], []
    tmp = torch.empty(max(sp.n_points for sp in specs) * R + 256, dtype=torch.uint8, device=dev)
    for f, sp in enumerate(specs):
        per = sp.n_points // n_strips
        w = (sp.hi[0] - sp.lo[0] + 1) // n_strips
        lo_all, hi_all = [2**31 - 1] * 3, [-(2**31)] * 3
        for k in range(n_strips):
            sk = copy.copy(sp)
            sk.seed = (sp.seed + 7919 * (k + 1)) & 0xFFFFFFFFFFFFFFFF
            sk.n_points = per if k < n_strips - 1 else sp.n_points - per * (n_strips - 1)
            sk.lo[0] = max(sp.lo[0], sp.lo[0] + k * w - w // 10)
            sk.hi[0] = min(sp.hi[0], sp.lo[0] + (k + 1) * w + w // 10)
            mm, _ = S.device_points(ctx, sk, tmp.data_ptr() + k * per * R)
            lo_all = [min(a, b) for a, b in zip(lo_all, mm[:3])]
            hi_all = [max(a, b) for a, b in zip(hi_all, mm[3:])]
        hi = HostImage(pcq, 227 + sp.n_points * R, S.header_bytes(sp, lo_all + hi_all), 0, sp.n_points * R, tmp)
        keep.append(hi)
        images.append(((hi.addr, hi.nbytes), "las"))
    torch.cuda.synchronize()
    del tmp
    torch.cuda.empty_cache()

    def run(searcher, index):
        cols = [[pcq.CountCollector(ctx) for _ in specs]]
        t0 = time.perf_counter()
        pcq.search_host_files_multi(images, [searcher], cols, index=index)
        got = [c.point_count() for c in cols[0]]
        dt = time.perf_counter() - t0
        st = ctx.last_scan_stats
        for c in cols[0]:
            c.close()
        return got, dt, st

    # One query per pass (an index cannot skip anything for a batch that contains the XL box: every chunk may match it).
    hix = pcq.HostIndex(ctx)
    run(searchers[-1], None)  # warm-up
    _, dt_first, _ = run(searchers[-1], hix)  # first indexed pass (XL: copies everything) builds the headers of every tile
    per_query, t_plain, t_indexed, h2d = {}, 0.0, 0.0, 0
    for (name, _), s_ in zip(qs, searchers):
        want, dt_plain, st0 = run(s_, None)
        dts, st = [], None
        for _ in range(3):
            got, dt, st = run(s_, hix)
            assert got == want, f"query {name}: the indexed pass differs from the unindexed pass over the same images"
            dts.append(dt)
        dt = statistics.median(dts)
        per_query[name] = {"unindexed_ms": dt_plain * 1e3, "indexed_ms": dt * 1e3, "matches": int(sum(got)),
                           "h2d_bytes_unindexed": int(st0.points_scanned) * R, "h2d_bytes_indexed": int(st.points_scanned) * R,
                           "chunks_skipped": int(st.chunks_skipped), "chunks_total": int(st.chunks_total)}
        t_plain += dt_plain
        t_indexed += dt
        h2d += int(st.points_scanned) * R
    out = {"value": scanned_pts / t_indexed / 1e9, "unit": UNIT, "ms_per_step": t_indexed * 1e3, "steps": 3,
           "unindexed_ms_per_step": t_plain * 1e3, "index_building_pass_ms": dt_first * 1e3, "h2d_bytes_per_step": h2d,
           "per_query": per_query,
           "note": "C2 shape with the points of every tile in acquisition order (16 flight strips per tile); S, L and XL as "
                   "three separate passes through pcq_search_host_files_indexed after the index-building pass (unindexed: the same "
                   "three passes through pcq_search_host_files_multi); counts equal the unindexed passes over the same images"}
    hix.close()
    for hi in keep:
        hi.close()
    return out


def density_record(pcq, group, ctx, stream, rank, world, local_rank, dev, max_over_ranks, sum_over_ranks, barrier, peak):
    """BASELINE configs[3]: navvis-shape dense indoor LAS (format 3, 56.2 M points), S / L / XL boxes + --density 0.1."""
    import numpy as np
    import torch

    from tests.helpers import same_point_set

    S, B = pcq.synth, pcq.binding
    sp = S.navvis_spec()
    N, R = sp.n_points, sp.record_len
    plan = pcq.shard_plan([N], world)
    mine = plan[plan["rank"] == rank]
    first, n = (int(mine[0]["first_point"]), int(mine[0]["n_points"])) if len(mine) else (0, 0)
    buf = torch.empty(n * R + 256, dtype=torch.uint8, device=dev)
    mm = S.device_point_range(local_rank, sp, first, n, buf.data_ptr())
    lo, hi = torch.tensor(mm[:3], device=dev), torch.tensor(mm[3:], device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    mm = [int(v) for v in lo.tolist()] + [int(v) for v in hi.tolist()]
    d = S.desc_of(sp, mm)
    d.n_points = n
    pieces = [(pcq.DeviceFile.wrap(ctx, d, buf.data_ptr(), first, keepalive=buf), 0, 0)] if n else []
    ds = group.wrap_files([N], pieces)
    # the oracle's answer, on rank 0: the same file generated on the host
    want = {}
    if rank == 0:
        from oracle import oracle as orc

        img = np.empty(227 + N * R, dtype=np.uint8)
        nthr = max(1, min(host_cores(), 32))
        cuts = [N * i // nthr // 8192 * 8192 for i in range(nthr)] + [N]
        ths = [threading.Thread(target=S.host_point_range, args=(sp, cuts[i], cuts[i + 1] - cuts[i], img[227 + cuts[i] * R: 227 + cuts[i + 1] * R]))
               for i in range(nthr)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        img[:227] = S.header_bytes(sp, mm)
    out = {"workload": f"C4 navvis-shape LAS fmt3, {N / 1e6:.1f} M points in one file, range-sharded over {world} GPU(s); --bounds + --density 0.1",
           "queries": {}}
    for name, box in (("S", S.NAVVIS_S), ("L", S.NAVVIS_L), ("XL", S.NAVVIS_XL)):
        s = pcq.BoundsSearcher(*box)
        grid = (box[0], box[1], S.NAVVIS_DENSITY)
        # matches (for the algorithmic bytes of the insert)
        matches = group.search(ds, [s], B.COLLECT_COUNT, False)[0].counts()[0]
        times, stats = [], None
        res = None
        for it in range(5):
            if res is not None:
                res.release()  # (hands its pinned buffer back to the group before the next search needs one)
            barrier()
            t0 = time.perf_counter()
            res = group.search(ds, [s], B.COLLECT_GRID, False, grid=grid)[0]
            dt = max_over_ranks(time.perf_counter() - t0)
            if it:
                times.append(dt * 1e3)
                stats = group.last_stats
        cells = res.counts()[0]
        ok = None
        if rank == 0:
            from oracle import oracle as orc

            oc = orc.Collector(orc.COLLECT_GRID, grid[0], grid[1], grid[2])
            orc.search_file(img, "las", oc, bounds=box)
            ok = bool(oc.point_count() == cells and same_point_set(res.points(0), oc.points()))
            assert ok, f"C4 {name}: sharded density result differs from the oracle's"
        st = {k: (max_over_ranks(float(v)) if k.endswith("_ms") else int(sum_over_ranks(float(v)))) for k, v in stats.items()}
        ms = statistics.median(times)
        alg = N * R + 8 * matches
        out["queries"][name] = {"matches": int(matches), "cells": int(cells), "ms_end_to_end_host_clock": ms, "phases": st,
                                "nvlink_bytes": st["bytes_sent"], "equals_oracle": ok,
                                "scan_insert_gbs_per_gpu": (alg / world) / max(st["scan_ms"], 1e-6) / 1e6,
                                "scan_insert_frac_of_peak": (alg / world) / max(st["scan_ms"], 1e-6) / 1e6 / peak,
                                "gpoints_per_s": N / ms / 1e6}
    ds.release()
    del pieces, buf
    torch.cuda.empty_cache()
    return out


def extra_records(pcq, ctx, stream, peak):
    """Driver-visible numbers for the other BASELINE configs on ONE GPU (device-resident, CUDA events, median of 5 after 2
    warm-ups).  `gbs` = algorithmic bytes (record_len / 12 / 1 byte read per point, + 31 bytes written per
    selected record, + 8 bytes per density match) / ms; `frac` = gbs / the measured HBM peak."""
    import torch

    S, B = pcq.synth, pcq.binding
    impl = pcq.SearchImplementation.Optimized
    out = []

    def timed(fn, reps=5):
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

    def rec(config, query, collector, points, matches, alg_bytes, ms, **kw):
        d = {"config": config, "query": query, "collector": collector, "points": int(points), "matches": int(matches),
             "algorithmic_bytes": int(alg_bytes), "ms": ms, "gbs": alg_bytes / ms / 1e6, "frac": alg_bytes / ms / 1e6 / peak,
             "gpoints_per_s": points / ms / 1e6}
        d.update(kw)
        out.append(d)

    def resident(sp):
        buf = torch.empty(sp.n_points * sp.record_len + 256, dtype=torch.uint8, device=f"cuda:{ctx.device}")
        mm, desc = S.device_points(ctx, sp, buf.data_ptr())
        return pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)

    def count_of(s, dfs):
        cols = [pcq.CountCollector(ctx) for _ in dfs]
        s.search_files(dfs, impl, cols)
        return sum(c.point_count() for c in cols)

    # ---- C1: one 10 M-point format-1 file, doc-S box, count ----
    sp = S.make_spec(0x5EED0000 + 1000, 10_000_000, B.LAYOUT_LAS, 1, B.SHAPE_TERRAIN, (0, 0, -9500), (99_999, 99_999, 76_000),
                     (0.01,) * 3, (390000.0, 130000.0, 0.0), S.DOC_CLASSES)
    df = resident(sp)
    s = pcq.BoundsSearcher(*S.DOC_S)
    c = pcq.CountCollector(ctx)
    ms = timed(lambda: s.search_files([df], impl, [c]))
    rec("C1 10M LAS fmt1", "bounds doc-S", "count", sp.n_points, count_of(s, [df]), sp.n_points * 28, ms)
    df.release()

    # ---- C5: 256 M uniform points, formats 0 / 1 / 3 and LAST, 50 % box, count ----
    for ext, fmt in (("las", 0), ("las", 1), ("las", 3), ("last", 1)):
        layout = B.LAYOUT_LAS if ext == "las" else B.LAYOUT_LAST
        sp = S.uniform_spec(1 << 28, layout, fmt)
        df = resident(sp)
        s = pcq.BoundsSearcher((0, 0, 0), (10000.0, 10000.0, 5000.0))
        c = pcq.CountCollector(ctx)
        ms = timed(lambda: s.search_files([df], impl, [c]))
        rec(f"C5 256M {ext} fmt{fmt}", "bounds 50 %", "count", sp.n_points, count_of(s, [df]),
            sp.n_points * (sp.record_len if ext == "las" else 12), ms)
        if ext == "last":
            s2 = pcq.ClassSearcher(2)
            ms = timed(lambda: s2.search_files([df], impl, [c]))
            rec(f"C5 256M {ext} fmt{fmt}", "class 2", "count", sp.n_points, count_of(s2, [df]), sp.n_points, ms)
        df.release()
        torch.cuda.empty_cache()

    # ---- C4: navvis, one GPU: count / select / density insert + finalise per box ----
    sp = S.navvis_spec()
    df = resident(sp)
    nb = sp.n_points * sp.record_len
    for name, box in (("S", S.NAVVIS_S), ("L", S.NAVVIS_L), ("XL", S.NAVVIS_XL)):
        s = pcq.BoundsSearcher(*box)
        matches = count_of(s, [df])
        g = pcq.GridSampledCollector(box[0], box[1], S.NAVVIS_DENSITY, ctx=ctx)

        def run():
            g.reset()
            s.search_files([df], impl, [g])

        ms_ins = timed(run)
        fin = []
        for _ in range(3):
            run()
            ctx.synchronize()
            t0 = time.perf_counter()
            cells = g.point_count()
            fin.append((time.perf_counter() - t0) * 1e3)
        rec("C4 navvis 56.2M LAS fmt3", f"bounds navvis-{name} + density 0.1", "grid (insert)", sp.n_points, matches, nb + 8 * matches, ms_ins,
            cells=int(cells), finalize_ms_host_clock=min(fin), insert_plus_finalize_frac=(nb + 8 * matches) / (ms_ins + min(fin)) / 1e6 / peak)
        g.close()
    df.release()
    torch.cuda.empty_cache()

    # ---- C3: ca13-shape LAST, 64 files x 40.75 M points, class 2 / 6 / 19, count and compacted output ----
    specs = S.ca13_specs()
    dfs = [resident(sp) for sp in specs]
    total = sum(sp.n_points for sp in specs)
    for klass in (2, 6, 19):
        s = pcq.ClassSearcher(klass)
        cols = [pcq.CountCollector(ctx) for _ in dfs]
        ms = timed(lambda: s.search_files(dfs, impl, cols))
        matches = count_of(s, dfs)
        rec("C3 ca13 LAST 64 x 40.75M", f"class {klass}", "count", total, matches, total, ms)
        bcols = [pcq.BufferCollector(ctx) for _ in dfs]

        def runb():
            pcq.reset_collectors(bcols)
            s.search_files(dfs, impl, bcols)

        ms = timed(runb, reps=3)
        # SURVEY §8d counts the gathered positions of a sparse select "at 32-byte sector granularity": a 12-byte position
        # at a 4-byte phase lies in 1.25 sectors on average (40 bytes), and no more than the whole column can be read
        sector_bytes = total + min(matches * 40, total * 12) + matches * 31
        rec("C3 ca13 LAST 64 x 40.75M", f"class {klass}", "buffer (compacted 31-byte records)", total, matches, total + matches * (12 + 31), ms,
            algorithmic_bytes_sector_granularity=int(sector_bytes), frac_sector_granularity=sector_bytes / ms / 1e6 / peak)
        for b in bcols:
            b.close()
        torch.cuda.empty_cache()
    for d_ in dfs:
        d_.release()
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
