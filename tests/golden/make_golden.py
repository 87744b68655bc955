"""Generates tests/golden/scan_golden.json: per-file match counts and a SHA-256 of the scan-ordered
31-byte record streams that the CPU ORACLE produces on small seeded synthetic datasets.

The reference is Rust and cannot be built or imported in this environment, so these vectors are
oracle outputs (not outputs of the reference binary); they freeze the oracle + generator behaviour
so that both the oracle (CPU test) and the CUDA path (GPU test) are checked against bytes that
travel with the repo.  Run:  python -m tests.golden.make_golden
"""
import hashlib
import json
from pathlib import Path

import numpy as np

CASES = ["c1_doc_s_fmt1", "doc_tiles_l_fma", "doc_tiles_xl_fmt3", "ca13_last_class2", "ca13_last_class6", "ca13_last_class19",
         "navvis_xl_fma", "navvis_l_last"]


def build_case(pcq, name):
    S, B = pcq.synth, pcq.binding
    if name == "c1_doc_s_fmt1":  # C1: one format-1 tile whose offset is the doc-S corner (KAV-1 local bounds)
        sp = S.make_spec(0x5EED0000 + 1000, 60_000, B.LAYOUT_LAS, 1, B.SHAPE_TERRAIN, (0, 0, -9488), (99_999, 99_999, 76_003),
                         (0.01,) * 3, (390000.0, 130000.0, 0.0))
        return [S.host_file(sp)], ["las"], {"bounds": S.DOC_S}
    if name == "doc_tiles_l_fma":
        specs = S.doc_specs(n_files=16, pts_per_file=6_000, fma_sensitive=True)
        return [S.host_file(s) for s in specs], ["las"] * 16, {"bounds": S.DOC_L}
    if name == "doc_tiles_xl_fmt3":
        specs = S.doc_specs(n_files=4, pts_per_file=7_001, fmt=3)
        return [S.host_file(s) for s in specs], ["las"] * 4, {"bounds": S.DOC_XL}
    if name.startswith("ca13_last_class"):
        specs = S.ca13_specs(n_files=4, pts_per_file=25_013)
        return [S.host_file(s) for s in specs], ["last"] * 4, {"cls": int(name[len("ca13_last_class"):])}
    if name == "navvis_xl_fma":
        return [S.host_file(S.navvis_spec(n_points=50_021, fma_sensitive=True))], ["las"], {"bounds": S.NAVVIS_XL}
    if name == "navvis_l_last":
        return [S.host_file(S.navvis_spec(n_points=50_021, layout=B.LAYOUT_LAST))], ["last"], {"bounds": S.NAVVIS_L}
    raise KeyError(name)


def oracle_answer(files, exts, kw):
    from oracle import oracle as orc

    counts, h = [], hashlib.sha256()
    for f, e in zip(files, exts):
        c = orc.Collector(orc.COLLECT_BUFFER)
        orc.search_file(f, e, c, **kw)
        counts.append(c.point_count())
        h.update(np.ascontiguousarray(c.points()).view(np.uint8).tobytes())
    # chunk headers of the on-the-fly index (8192-point chunks; fields lo, hi, cls_bits, n_points, pad = 0), file by file
    hh = hashlib.sha256()
    for f, e in zip(files, exts):
        hh.update(orc.chunk_headers(f, e).tobytes())
    return {"counts": counts, "buffer_sha256": h.hexdigest(), "chunk_headers_sha256": hh.hexdigest()}


def main():
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
    from pcq_import import pcq

    out = {}
    for name in CASES:
        files, exts, kw = build_case(pcq, name)
        out[name] = oracle_answer(files, exts, kw)
        out[name]["input_sha256"] = hashlib.sha256(b"".join(f.tobytes() for f in files)).hexdigest()
        print(name, out[name]["counts"])
    (Path(__file__).parent / "scan_golden.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
