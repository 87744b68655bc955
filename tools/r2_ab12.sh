#!/bin/bash
mkdir -p gpurun_out
{
for g in 32 64 128 0; do
  echo "granularity=$g"
  for c in 6 2; do PCQ_VERBOSE=1 PCQ_L2_FETCH_GRANULARITY=$g python tools/grid_probe.py lastsel $c 7 2>&1 | grep -v "^k_"; done
  PCQ_L2_FETCH_GRANULARITY=$g python tools/grid_probe.py grid XL 7
done
} > gpurun_out/ab12.txt 2>&1
cat gpurun_out/ab12.txt
for g in 32 128; do
PCQ_L2_FETCH_GRANULARITY=$g ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_lastsel_6_g$g.csv python tools/grid_probe.py lastsel 6 3 > /dev/null 2>&1
grep k_select_bytes gpurun_out/launches_lastsel_6_g$g.csv | tail -3 | awk -F'","' '{print $(NF-2), $NF}'
done
PCQ_L2_FETCH_GRANULARITY=32 python tools/sweep.py --points 134217728 --only count --variants 2 --cases las:1,last:1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['layout'],d['format'],d['query'],d['collector'],round(d['ms'],3),round(d['frac_of_measured_peak'],3))"
