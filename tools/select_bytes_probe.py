#!/usr/bin/env python
"""C3-shape LAST class select on a few files (for `ncu -k regex:k_select_bytes`): python tools/select_bytes_probe.py [files] [class]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from pcq_import import pcq  # noqa: E402

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 8
klass = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = pcq.synth
ctx = pcq.Context(0)
dfs, keep = [], []
for sp in S.ca13_specs()[:n_files]:
    buf = torch.empty(sp.n_points * sp.record_len + 256, dtype=torch.uint8, device="cuda:0")
    mm, desc = S.device_points(ctx, sp, buf.data_ptr())
    dfs.append(pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf))
s = pcq.ClassSearcher(klass)
cols = [pcq.BufferCollector(ctx) for _ in dfs]
for rep in range(3):
    for c in cols:
        c.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.synchronize()
    e0.record()
    s.search_files(dfs, pcq.SearchImplementation.Optimized, cols)
    ctx.synchronize()
    e1.record()
    torch.cuda.synchronize()
    print(rep, "ms", e0.elapsed_time(e1), "matches", sum(c.point_count() for c in cols), flush=True)
