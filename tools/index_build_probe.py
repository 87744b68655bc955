#!/usr/bin/env python
"""Launches k_chunk_index a few times on an acquisition-order file (for `ncu -k regex:k_chunk_index`)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pcq_import import pcq  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
ctx = pcq.Context(0)
for layout in (pcq.binding.LAYOUT_LAS, pcq.binding.LAYOUT_LAST):
    buf, desc, _ = pcq.synth.strips_device(ctx, n, 64, layout, 1)
    df = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
    for _ in range(3):
        df.drop_index()
        df.build_index()
    print(layout, df.index.shape[0], flush=True)
    df.release()
