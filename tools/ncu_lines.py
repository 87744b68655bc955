#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: stall samples per CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > mix.csv
    python tools/ncu_lines.py mix.csv [top]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path)))
    hdr = None
    per_line = {}
    src_text = {}
    total = 0
    for row in rows:
        if len(row) > 6 and row[0] == "Line No":
            hdr = row
            i_samp = hdr.index("# Samples")
            i_inst = hdr.index("Instructions Executed")
            stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(row) <= i_samp:
            continue
        if row[2] != "-":  # a SASS row; source-line rows carry the aggregate already
            continue
        try:
            s = int(float(row[i_samp]))
        except ValueError:
            continue
        key = row[0]
        d = per_line.setdefault(key, {"samples": 0, "inst": 0, "stalls": {}})
        d["samples"] += s
        d["inst"] += int(float(row[i_inst] or 0))
        for j, h in stall_cols:
            try:
                v = int(float(row[j]))
            except ValueError:
                v = 0
            if v:
                d["stalls"][h] = d["stalls"].get(h, 0) + v
        src_text[key] = row[1]
        total += s
    print(f"total samples {total}")
    for key, d in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print(f"{d['samples']:7d} {100.0 * d['samples'] / max(total, 1):5.1f}%  inst={d['inst']:9d}  L{key:>5}: {src_text[key].strip()[:90]}   [{st}]")


if __name__ == "__main__":
    main()
