"""CPU tests of bench.py's plumbing: the reference arm runs without the product library and on the same bytes as the
B200 arm; the JSON line carries what the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_specs_equal_the_package_specs(pcq):
    sys.path.insert(0, str(ROOT))
    import bench

    S = pcq.synth
    mine = bench.load_synth_only().doc_specs(16, 12345, bench.SEED)
    theirs = S.doc_specs(n_files=16, pts_per_file=12345, seed=bench.SEED)
    assert len(mine) == len(theirs)
    for a, b in zip(mine, theirs):
        assert bytes(a) == bytes(b)
    img = np.empty(S.file_size(theirs[3]), dtype=np.uint8)
    bench.load_synth_only().host_file_into(mine[3], img)
    assert np.array_equal(img, S.host_file(theirs[3]))


def test_reference_arm_line_and_libraries():
    code = ("import bench, sys, json; sys.argv=['bench.py','--impl','reference','--files','4','--pts-per-file','40000','--steps','2','--warmup','1'];"
            "bench.main(); maps=open('/proc/self/maps').read(); print('MAPS', 'libpcq.so' in maps, 'libpcq_synth.so' in maps, 'liboracle.so' in maps)")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Gpoints/s" and line["higher_is_better"] is True
    assert line["scaling"] == "strong" and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["same_dataset_as_b200_arm"] is True
    # the reference arm maps the oracle and the generator, never the product library
    assert lines[-1] == "MAPS False True True"


def test_golden_counts_file_is_well_formed():
    p = ROOT / "tests" / "golden" / "c2_counts.json"
    if p.exists():
        g = json.loads(p.read_text())
        for k, v in g.items():
            assert k.startswith("files=") and set(v["matches_per_step"]) == {"S", "L", "XL"}
