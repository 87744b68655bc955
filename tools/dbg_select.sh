#!/bin/bash
# (needs a library built with the measurement hooks: make -C adhoc-queries-pointclouds_b200/csrc clean all EXTRA=-DPCQ_DEBUG_HOOKS)
# timing probe for the select kernels: PCQ_SELECT_DEBUG modes x scan variants (2 = k_select_ring where it applies, 1 = k_select)
for v in ${DBG_VARIANTS:-2}; do for d in ${DBG_MODES:-0 1 2 3}; do echo "== variant $v debug $d"; PCQ_SELECT_DEBUG=$d timeout 200 python tools/sweep.py --points ${DBG_POINTS:-67108864} --only select --variants $v --cases ${DBG_CASES:-las:0,las:1,las:3,last:1} 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d['layout'],d['format'],d['query'],'ms=%.3f frac=%.2f'%(d['ms'],d['frac_of_measured_peak']))
"; done; done
