"""Build-time check that no multiply-add is contracted on the path (hygiene): Rust evaluates
`(v as f64 * scale) + offset` and the squared distance with one rounding per operation, so the SASS must not hold
DFMA / FFMA where those run.  Two angles: (1) the shared device functions, compiled alone with the library's flags;
(2) every kernel of the built libpcq.so that does no division — counts, selects, index, finalisation — holds no FMA at
all (the density insert and the alias fold divide, and an IEEE division is a sequence of FMAs by construction)."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "adhoc-queries-pointclouds_b200" / "csrc"
LIB = ROOT / "adhoc-queries-pointclouds_b200" / "libpcq.so"
# (HFMA2 R, -RZ, RZ, 0, 0 is ptxas' way of zeroing a register on the half pipe, not arithmetic)
FMA = re.compile(r"\b(DFMA|FFMA)\b")

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit")


def _fma_per_function(sass: str):
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = 0
        elif cur is not None and FMA.search(line):
            out[cur] += 1
    return out


def _makefile_flags():
    text = (CSRC / "Makefile").read_text()
    assert "--fmad=false" in text, "the library must be compiled with --fmad=false"
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-std=c++17"]


def test_shared_arithmetic_has_no_fma(tmp_path):
    cubin = tmp_path / "probe.cubin"
    subprocess.check_call(["nvcc", *_makefile_flags(), "-I", str(CSRC), "-cubin", "-o", str(cubin), str(ROOT / "tests" / "probes" / "fma_probe.cu")])
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True, check=True).stdout
    per = _fma_per_function(sass)
    assert set(per) == {"probe_reconstruct", "probe_dist2", "probe_cells"}
    assert per["probe_reconstruct"] == 0 and per["probe_dist2"] == 0
    assert per["probe_cells"] > 0  # the division sequences: the probe does see FMAs where they belong


def test_library_kernels_without_division_have_no_fma():
    assert LIB.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = _fma_per_function(sass)
    assert len(per) > 40
    # kernels that evaluate SparseGrid::insert_point (three divisions per point) or the LAS writer's quantisation
    divides = ("k_grid_scan", "k_alias_fold", "k_points_to_las2")
    checked = 0
    for name, n in per.items():
        if any(d in name for d in divides):
            continue
        if "k_scan_direct" in name or "k_scan_staged" in name:
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout
            m = re.search(r"k_scan_direct<(\d)>|k_scan_staged<\d+, (\d),", demangled)
            mode = int(m.group(1) or m.group(2))
            if mode in (2, 3):
                continue
        assert n == 0, f"{name}: {n} fused multiply-adds"
        checked += 1
    assert checked > 25
