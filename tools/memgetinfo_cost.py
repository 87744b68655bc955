#!/usr/bin/env python
"""How long cudaMemGetInfo takes on this box (it sat on the select / density launch path until it was made lazy)."""
import time

import torch

torch.cuda.init()
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for phase in ("idle", "after_work", "after_free"):
    if phase == "after_work":
        for _ in range(20):
            x.add_(1)
    if phase == "after_free":
        y = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
        del y
        torch.cuda.empty_cache()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        torch.cuda.mem_get_info()
        ts.append((time.perf_counter() - t0) * 1e6)
    print(phase, "us: min %.0f median %.0f max %.0f" % (min(ts), sorted(ts)[10], max(ts)), flush=True)
