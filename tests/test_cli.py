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

    # the records in the first output file are the oracle's records (positions re-quantised by the writer)
    names = sorted(p.name for p in d.iterdir() if p.suffix == f".{ext}")
    r = subprocess.run([str(QUERY), "-i", str(d / names[0]), "--class", "6", "--optimized", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

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
