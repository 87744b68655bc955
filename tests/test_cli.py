"""The `query` CLI (query/src/main.rs surface) against the oracle: stdout lines, counts, LAS output."""
import re
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = Path(__file__).resolve().parent.parent
QUERY = ROOT / "adhoc-queries-pointclouds_b200" / "query"


def test_cli_binary_is_built():
    assert QUERY.exists(), "run __graft_entry__.build()"
    out = subprocess.run([str(QUERY), "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--bounds" in out.stderr and "--density" in out.stderr


def test_cli_argument_errors_need_no_gpu(tmp_path):
    r = subprocess.run([str(QUERY), "--bounds", "0;0;0;1;1;1"], capture_output=True, text=True)
    assert r.returncode == 1 and "--input" in r.stderr
    (tmp_path / "a.las").write_bytes(b"")
    r = subprocess.run([str(QUERY), "-i", str(tmp_path), "--bounds", "0;0;0;1;1", "--optimized"], capture_output=True, text=True)
    assert r.returncode == 101 and "Could not prase argument BOUNDS" in r.stderr  # .expect() panic, main.rs:235
    r = subprocess.run([str(QUERY), "-i", str(tmp_path), "--bounds", "5;0;0;1;1;1", "--optimized"], capture_output=True, text=True)
    assert r.returncode == 101  # AABB::from_min_max panics on min > max
    r = subprocess.run([str(QUERY), "-i", str(tmp_path), "--class", "300", "--optimized"], capture_output=True, text=True)
    assert r.returncode == 101 and "CLASS" in r.stderr
    r = subprocess.run([str(QUERY), "-i", str(tmp_path), "--class", "3", "--bounds", "0;0;0;1;1;1"], capture_output=True, text=True)
    assert r.returncode == 1 and "at the same time" in r.stderr
    r = subprocess.run([str(QUERY), "-i", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "neither BOUNDS nor CLASS" in r.stderr
    r = subprocess.run([str(QUERY), "-i", str(tmp_path / "nope"), "--class", "3"], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stderr


def _dataset(pcq, d: Path, layout_ext: str):
    S, B = pcq.synth, pcq.binding
    layout = B.LAYOUT_LAS if layout_ext == "las" else B.LAYOUT_LAST
    specs = S.doc_specs(n_files=4, pts_per_file=30_000, layout=layout, fmt=3 if layout_ext == "las" else 2)
    files = []
    for k, sp in enumerate(specs):
        img = S.host_file(sp)
        (d / f"tile_{k}.{layout_ext}").write_bytes(img.tobytes())
        files.append(img)
    (d / "notes.txt").write_text("not a point cloud")  # is_valid_file filters it out (main.rs:185-189)
    return files


def _read_las_fmt2(path: Path):
    b = path.read_bytes()
    assert b[:4] == b"LASF" and b[24:26] == b"\x01\x02" and b[104] == 2
    n = struct.unpack_from("<I", b, 107)[0]
    off = struct.unpack_from("<I", b, 96)[0]
    scale = struct.unpack_from("<3d", b, 131)
    offset = struct.unpack_from("<3d", b, 155)
    rec = np.frombuffer(b, dtype=np.uint8, count=n * 26, offset=off).reshape(n, 26)
    xyz = np.ascontiguousarray(rec[:, :12]).view("<i4").reshape(n, 3)
    return n, np.array(scale), np.array(offset), xyz, rec[:, 15].copy(), np.ascontiguousarray(rec[:, 20:26]).view("<u2").reshape(n, 3)


@pytest.mark.parametrize("ext", ["las", "last"])
def test_cli_box_that_touches_no_file_needs_no_gpu(pcq, tmp_path, ext):
    """Every search returns before its per-point loop when the file's header box does not intersect the query box
    (las.rs:82-84, last.rs:92-94): the CLI answers such a query from the headers, with the reference's lines."""
    d = tmp_path / "data"
    d.mkdir()
    files = _dataset(pcq, d, ext)
    far = "0;0;0;10;10;10"
    assert orc.count_parallel(files, [ext] * 4, 4, bounds=((0, 0, 0), (10, 10, 10))).sum() == 0
    for flags in (["--parallel"], []):
        r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", far, "--optimized", *flags], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        assert lines[0] == "Searching 4 files..." and lines[1] == "Found 0 matching points"
        assert re.match(r"Searched \d+\.\d\d MiB in \d+\.\d\ds \(throughput: \d+\.\d\dMiB/s\)", lines[-1])
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", far, "--optimized", "--parallel", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and "Found" not in r.stdout and "Writing" not in r.stdout and not list(out.iterdir())
    r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", far, "--optimized", "--density", "1.0"], capture_output=True, text=True)
    assert r.returncode == 0 and "Found" not in r.stdout
    # the refusal of the Regular implementation comes first
    r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", far], capture_output=True, text=True)
    assert r.returncode == 1 and ("Regular" in r.stderr or "CUDA" in r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("ext", ["las", "last"])
def test_cli_matches_oracle(pcq, tmp_path, ext):
    S = pcq.synth
    d = tmp_path / "data"
    d.mkdir()
    files = _dataset(pcq, d, ext)
    box = S.DOC_L
    bstr = ";".join(str(v) for v in (*box[0], *box[1]))
    want = orc.count_parallel(files, [ext] * 4, 4, bounds=box)

    # count mode, parallel and sequential
    for flags in (["--parallel"], []):
        r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", bstr, "--optimized", *flags], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        assert lines[0] == "Searching 4 files..."
        assert f"Found {int(want.sum())} matching points" in lines
        assert re.match(r"Searched \d+\.\d\d MiB in \d+\.\d\ds \(throughput: \d+\.\d\dMiB/s\)", lines[-1])

    # class 6 with output: one LAS 1.2 format-2 file per non-empty per-file buffer, no "Found" line
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([str(QUERY), "-i", str(d), "--class", "6", "--optimized", "--parallel", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Found" not in r.stdout
    per_file = orc.count_parallel(files, [ext] * 4, 4, cls=6)
    written = sorted(out.glob("matching_points_*.las"), key=lambda p: int(p.stem.split("_")[-1]))
    assert len(written) == int((per_file > 0).sum())
    assert sorted(int(x) for x in re.findall(r"Writing (\d+) points", r.stdout)) == sorted(int(v) for v in per_file if v > 0)
    total = 0
    for p in written:
        n, scale, offset, xyz, cls, rgb = _read_las_fmt2(p)
        total += n
        assert (cls == 6).all() and xyz.min() >= 0 and scale[0] == scale[1] == scale[2] >= 0.001
    assert total == int(per_file.sum())

    # the records of EVERY output file are the oracle's records as the oracle's restatement of FileDumper
    # (dump_points.rs:63-116, oracle/np_oracle.py dump_points_plan) writes them: offset = min position, one scale for all
    # axes, raw coordinates, classification and colour of every record in scan order; `k` counts non-empty buffers only
    from oracle import np_oracle

    names = sorted(p.name for p in d.iterdir() if p.suffix == f".{ext}")
    listed = [p.name for p in d.iterdir() if p.suffix == f".{ext}"]  # read_dir order == the CLI's file order
    buffers = []
    for name in listed:
        oc = orc.Collector(orc.COLLECT_BUFFER)
        orc.search_file(files[names.index(name)], ext, oc, cls=6)
        buffers.append(oc.points())
    plan = np_oracle.dump_points_plan(buffers)
    assert [f"matching_points_{e['index']}.las" for e in plan] == [p.name for p in written]
    for e, p in zip(plan, written):
        n, scale, offset, xyz, cls, rgb = _read_las_fmt2(p)
        assert n == len(e["raw"]) and np.array_equal(scale, [e["scale"]] * 3) and np.array_equal(offset, e["offset"])
        assert np.array_equal(xyz, e["raw"]) and np.array_equal(cls, e["cls"]) and np.array_equal(rgb, e["rgb"])
    # an empty first buffer does not consume a file index (dump_points.rs:65-71): class 7 is absent from no tile but rare;
    # a box that misses tile 0 entirely gives an empty first buffer
    out3 = tmp_path / "out3"
    out3.mkdir()
    r = subprocess.run([str(QUERY), "-i", str(d / names[0]), "--class", "6", "--optimized", "-o", str(out3)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    oc = orc.Collector(orc.COLLECT_BUFFER)
    orc.search_file(files[0], ext, oc, cls=6)
    (e,) = np_oracle.dump_points_plan([oc.points()])
    n, scale, offset, xyz, cls, rgb = _read_las_fmt2(out3 / "matching_points_0.las")
    assert np.array_equal(xyz, e["raw"]) and np.array_equal(rgb, e["rgb"]) and np.array_equal(offset, e["offset"])

    # density, no output: no count line at all (main.rs:137-141 with GridSampledCollector::points() = Some)
    r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", bstr, "--optimized", "--density", "25"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Found" not in r.stdout and "Writing" not in r.stdout

    # density with output, sequential: one grid over all files -> one output file with one point per cell
    out2 = tmp_path / "out2"
    out2.mkdir()
    xl = ";".join(str(v) for v in (*S.DOC_XL[0], *S.DOC_XL[1]))
    r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", xl, "--optimized", "--density", "25", "-o", str(out2)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # oracle in the same file order the CLI used (read_dir order is OS dependent; ties depend on it)
    og = orc.Collector(orc.COLLECT_GRID, S.DOC_XL[0], S.DOC_XL[1], 25.0)
    for f in files:
        orc.search_file(f, ext, og, bounds=S.DOC_XL)
    n, *_ = _read_las_fmt2(out2 / "matching_points_0.las")
    assert n == og.point_count()

    # without --optimized the request is refused loudly (Regular stays on the reference)
    r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", bstr], capture_output=True, text=True)
    assert r.returncode == 1 and "Regular" in r.stderr


# ---- the experiment harness (tools/run_query_experiments.py <- run_query_experiments.rs:106-380) -----------------------
def _harness(args, timeout=600):
    import sys

    return subprocess.run([sys.executable, str(ROOT / "tools" / "run_query_experiments.py"), *args], capture_output=True, text=True, timeout=timeout)


LINE = re.compile(r"^[a-z0-9_]+;\d+\.\d{6};\d+\.\d{6};\d+\.\d{6}$")


def test_experiment_harness_cpu_arm(tmp_path):
    """the CPU column alone needs no GPU: datasets, every experiment name of the reference's runner, its line format"""
    assert (ROOT / "oracle" / "query_ref").exists(), "run __graft_entry__.build()"
    r = _harness(["--generate", str(tmp_path), "--scale", "0.0001", "--input", str(tmp_path), "--runs", "2", "--arms", "cpu"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ";" in ln]
    assert all(LINE.match(ln) for ln in lines), lines
    names = [ln.split(";")[0] for ln in lines]
    # S / L / XL x full / lod x las / last for three datasets, class building / noclass x las / last for two
    assert len(names) == 3 * 3 * 2 * 2 + 2 * 2 * 2 and all(n.endswith("_cpu") for n in names)
    for want in ("navvis3_s_full_las_cpu", "doc_xl_lod_last_cpu", "ca13_l_full_last_cpu", "doc_class_building_las_cpu", "ca13_class_noclass_last_cpu"):
        assert want in names


@pytest.mark.gpu
def test_experiment_harness_gpu_and_cpu_columns_agree(tmp_path):
    """both arms, one line each per experiment; the harness itself fails when the two `Found N` lines differ"""
    r = _harness(["--generate", str(tmp_path), "--scale", "0.0005", "--input", str(tmp_path), "--runs", "2", "--experiment", "2"], timeout=1200)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ";" in ln]
    names = [ln.split(";")[0] for ln in lines]
    assert all(LINE.match(ln) for ln in lines) and len(names) == 2 * 12
    assert all(n + "_cpu" in names for n in names if not n.endswith("_cpu"))


@pytest.mark.gpu
@pytest.mark.parametrize("ext", ["las", "last"])
def test_cli_on_a_group_of_gpus(pcq, tmp_path, ext):
    """`query --gpus N`: files and point ranges shard over N GPUs (pcq_group, one process); stdout lines, counts and
    written records are those of one GPU and of the oracle — including doc-S --density 25, whose grid aliases."""
    import torch

    from oracle import np_oracle

    S = pcq.synth
    d = tmp_path / "data"
    d.mkdir()
    files = _dataset(pcq, d, ext)
    env = dict(**__import__("os").environ)
    n = int(env.get("PCQ_CLI_GPUS", "3"))  # (PCQ_CLI_GPUS=8 on an 8-GPU box: `query --gpus 8`)
    if torch.cuda.device_count() < n:
        env["PCQ_GROUP_DEVICES"] = ",".join(str(i % torch.cuda.device_count()) for i in range(n))
    names = sorted(p.name for p in d.iterdir() if p.suffix == f".{ext}")
    listed = [p.name for p in d.iterdir() if p.suffix == f".{ext}"]
    order = [names.index(x) for x in listed]  # the CLI's file order (read_dir)
    box = S.DOC_L
    bstr = ";".join(str(v) for v in (*box[0], *box[1]))
    want = orc.count_parallel(files, [ext] * 4, 4, bounds=box)
    for flags in (["--parallel"], []):
        r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", bstr, "--optimized", "--gpus", str(n), *flags], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        assert f"Found {int(want.sum())} matching points" in r.stdout.splitlines()
    # select with output, per file: the same files FileDumper's restatement writes
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([str(QUERY), "-i", str(d), "--class", "6", "--optimized", "--parallel", "--gpus", str(n), "-o", str(out)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    buffers = []
    for k in order:
        oc = orc.Collector(orc.COLLECT_BUFFER)
        orc.search_file(files[k], ext, oc, cls=6)
        buffers.append(oc.points())
    plan = np_oracle.dump_points_plan(buffers)
    written = sorted(out.glob("matching_points_*.las"), key=lambda p: int(p.stem.split("_")[-1]))
    assert len(written) == len(plan)
    for e, p in zip(plan, written):
        nrec, scale, offset, xyz, cls, rgb = _read_las_fmt2(p)
        assert np.array_equal(xyz, e["raw"]) and np.array_equal(cls, e["cls"]) and np.array_equal(rgb, e["rgb"]) and np.array_equal(offset, e["offset"])
    # doc-S + --density 25 (8 z-cells, inclusive z = 200.00 face: aliased keys), sequential: one grid over all files
    out2 = tmp_path / "out2"
    out2.mkdir()
    sb = ";".join(str(v) for v in (*S.DOC_S[0], *S.DOC_S[1]))
    for gpus, odir in ((1, tmp_path / "o1"), (n, out2)):
        odir.mkdir(exist_ok=True)
        r = subprocess.run([str(QUERY), "-i", str(d), "--bounds", sb, "--optimized", "--density", "25", "--gpus", str(gpus), "-o", str(odir)],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
    og = orc.Collector(orc.COLLECT_GRID, S.DOC_S[0], S.DOC_S[1], 25.0)
    for k in order:
        orc.search_file(files[k], ext, og, bounds=S.DOC_S)
    (e,) = np_oracle.dump_points_plan([og.points()]) if og.point_count() else (None,)
    if e is not None:
        for odir in (tmp_path / "o1", out2):
            nrec, scale, offset, xyz, cls, rgb = _read_las_fmt2(odir / "matching_points_0.las")
            assert nrec == og.point_count() and np.array_equal(offset, e["offset"])
            got = np.concatenate([xyz, cls[:, None].astype(np.int32), rgb.astype(np.int32)], axis=1)
            exp = np.concatenate([e["raw"], e["cls"][:, None].astype(np.int32), e["rgb"].astype(np.int32)], axis=1)
            assert np.array_equal(got[np.lexsort(got.T[::-1])], exp[np.lexsort(exp.T[::-1])])  # HashMap order is arbitrary
