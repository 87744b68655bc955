#!/usr/bin/env python
"""run_query_experiments — the reference's experiment runner (query/src/bin/run_query_experiments.rs)
re-expressed for the B200 scan path.

Like the reference it spawns the `query` binary as a subprocess per run (run_query_experiments.rs:29-75),
`--optimized --parallel`, no `-o` (so bbox runs print a match count and density runs print none), 5 runs per
experiment (:412-413), and prints one `name;mean;median;stddev` line per experiment (:287-304).  Experiments:
S / L / XL boxes x {full, lod (max-density)} for navvis / doc / ca13 (:109-266) and class 6 / class 19 (:320-343),
on the las and last flavours of each dataset (laz / lazer stay on the reference).

The reference's datasets are private; `--generate DIR` writes seeded synthetic datasets of the same shapes
(`--scale` shrinks them: 1.0 = 56.2 M / 64 x 31.25 M / 64 x 40.75 M points).  The page-cache purge of the
reference (:8-27, macOS `purge`) is replaced by `--drop-caches` (needs root; off by default).

Every experiment is run twice: with the GPU `query` and — the CPU column, `<name>_cpu` — with `oracle/query_ref`, the
same command line on the CPU oracle (the reference binary itself cannot be built here).  The two must print the same
"Found N matching points" line or the run fails.  `--arms gpu` / `--arms cpu` run one of them only (the CPU arm needs no
GPU); `--gpus N` shards every query over N GPUs of the box (`query --gpus N`).

    python tools/run_query_experiments.py --generate /dev/shm/pcq --scale 0.01
    python tools/run_query_experiments.py --input /dev/shm/pcq --experiment 2
"""
import argparse
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
QUERY = ROOT / "adhoc-queries-pointclouds_b200" / "query"
QUERY_CPU = ROOT / "oracle" / "query_ref"
EXTS = ["las", "last"]
ARMS = ["gpu", "cpu"]
GPUS = 1


def generate(root: Path, scale: float):
    from pcq_import import pcq

    S, B = pcq.synth, pcq.binding
    for ext in EXTS:
        layout = B.LAYOUT_LAS if ext == "las" else B.LAYOUT_LAST
        d = root / "navvis3" / ext
        d.mkdir(parents=True, exist_ok=True)
        S.host_file(S.navvis_spec(n_points=max(1000, int(56_200_000 * scale)), layout=layout)).tofile(d / f"navvis_m6_HQ3rdFloor.{ext}")
        n_files = 64 if scale >= 0.05 else 16
        d = root / "doc" / ext
        d.mkdir(parents=True, exist_ok=True)
        for k, sp in enumerate(S.doc_specs(n_files=n_files, pts_per_file=max(1000, int(2_000_000_000 * scale / n_files)), layout=layout)):
            S.host_file(sp).tofile(d / f"doc_{k:02d}.{ext}")
        d = root / "ca13" / ext
        d.mkdir(parents=True, exist_ok=True)
        for k, sp in enumerate(S.ca13_specs(n_files=n_files, pts_per_file=max(1000, int(2_608_000_000 * scale / n_files)), layout=layout)):
            S.host_file(sp).tofile(d / f"ca13_{k:02d}.{ext}")
    print(f"datasets written under {root}")


def purge(drop: bool):
    subprocess.run(["sync"], check=False)
    if drop:
        try:
            Path("/proc/sys/vm/drop_caches").write_text("3\n")
        except OSError:
            pass


def run_query(binary, args_list, drop):
    purge(drop)
    t0 = time.perf_counter()
    r = subprocess.run([str(binary), *args_list], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"Could not execute query. Process exited with {r.returncode}: {r.stderr.strip()}")
    found = [ln for ln in r.stdout.splitlines() if ln.startswith("Found ")]
    return dt, (found[0] if found else None)


def report(name, times):
    mean = statistics.mean(times)
    median = statistics.median(times)
    std = statistics.stdev(times) if len(times) > 1 else 0.0
    print(f"{name};{mean:.6f};{median:.6f};{std:.6f}", flush=True)


def experiment(name, args_list, runs, drop, gpu):
    """one experiment on both arms; the arms must agree on what they found"""
    found = {}
    for arm in ARMS:
        if arm == "gpu":
            extra = ["--gpu", str(gpu)] + (["--gpus", str(GPUS)] if GPUS > 1 else [])
            res = [run_query(QUERY, args_list + extra, drop) for _ in range(runs)]
        else:
            res = [run_query(QUERY_CPU, args_list, drop) for _ in range(runs)]
        assert len({f for _, f in res}) == 1, f"{name}: runs of the {arm} arm disagree"
        found[arm] = res[0][1]
        report(name if arm == "gpu" else f"{name}_cpu", [t for t, _ in res])
    if len(found) == 2 and found["gpu"] != found["cpu"]:
        raise RuntimeError(f"{name}: GPU printed {found['gpu']!r}, CPU oracle printed {found['cpu']!r}")


def aabb_experiments(root: Path, runs: int, which: int, drop: bool, gpu: int):
    from pcq_import import pcq

    S = pcq.synth
    sets = {
        1: ("navvis3", [("s", S.NAVVIS_S), ("l", S.NAVVIS_L), ("xl", S.NAVVIS_XL)], S.NAVVIS_DENSITY),
        2: ("doc", [("s", S.DOC_S), ("l", S.DOC_L), ("xl", S.DOC_XL)], S.DOC_DENSITY),
        3: ("ca13", [("s", S.CA13_S), ("l", S.CA13_L), ("xl", S.CA13_XL)], S.CA13_DENSITY),
    }
    dataset, boxes, density = sets[which]
    for name, (qmin, qmax) in boxes:
        for dens in (None, density):
            for ext in EXTS:
                d = root / dataset / ext
                if not d.exists():
                    continue
                b = ";".join(repr(float(v)) for v in (*qmin, *qmax))
                a = ["-i", str(d), "--bounds", b, "--optimized", "--parallel"]
                if dens is not None:
                    a += ["--density", repr(float(dens))]
                experiment(f"{dataset}_{name}_{'lod' if dens is not None else 'full'}_{ext}", a, runs, drop, gpu)


def class_experiments(root: Path, runs: int, drop: bool, gpu: int):
    for dataset in ("doc", "ca13"):
        for cname, klass in (("building", 6), ("noclass", 19)):
            for ext in EXTS:
                d = root / dataset / ext
                if not d.exists():
                    continue
                a = ["-i", str(d), "--class", str(klass), "--optimized", "--parallel"]
                experiment(f"{dataset}_class_{cname}_{ext}", a, runs, drop, gpu)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", type=Path, help="root with <dataset>/<ext>/ directories")
    ap.add_argument("--experiment", type=int, default=0, help="1 navvis, 2 doc, 3 ca13, 4 classes, 0 all")
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--generate", type=Path)
    ap.add_argument("--scale", type=float, default=0.01)
    ap.add_argument("--drop-caches", action="store_true")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1, help="shard every query over N GPUs (query --gpus N)")
    ap.add_argument("--arms", default="gpu,cpu", help="gpu,cpu (default), gpu or cpu")
    a = ap.parse_args()
    global ARMS, GPUS
    ARMS = [x for x in a.arms.split(",") if x]
    GPUS = a.gpus
    if a.generate:
        generate(a.generate, a.scale)
        if not a.input:
            return
    if not a.input:
        ap.error("--input or --generate is required")
    which = [1, 2, 3, 4] if a.experiment == 0 else [a.experiment]
    for w in which:
        if w in (1, 2, 3):
            aabb_experiments(a.input, a.runs, w, a.drop_caches, a.gpu)
        else:
            class_experiments(a.input, a.runs, a.drop_caches, a.gpu)


if __name__ == "__main__":
    main()
