#!/bin/bash
# full GPU suite + A/B density 3 vs 4 CTAs per SM + select sweep after the dispatcher change (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
{
for b in XL L S; do echo "minb=3"; python tools/grid_probe.py grid $b 7; done
for c in 19 6 2; do python tools/grid_probe.py lastsel $c 7; done
cp adhoc-queries-pointclouds_b200/libpcq.so /tmp/libpcq_main.so
cp adhoc-queries-pointclouds_b200/csrc/build/abm4/libpcq.so adhoc-queries-pointclouds_b200/libpcq.so
for b in XL L; do echo "minb=4"; PCQ_VERBOSE=1 python tools/grid_probe.py grid $b 7 2>&1 | sort | uniq; done
cp /tmp/libpcq_main.so adhoc-queries-pointclouds_b200/libpcq.so
} > gpurun_out/ab5.txt 2>&1
cat gpurun_out/ab5.txt
python tools/sweep.py --points 134217728 --only select --variants 2 > gpurun_out/sweep_select_128M.jsonl 2> gpurun_out/sweep_select.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_select_128M.jsonl'):
    d=json.loads(l); print(d['layout'],d['format'],d['query'],round(d['ms'],3),round(d['frac_of_measured_peak'],3))
PY
for c in 19 6; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lastsel_$c.csv python tools/grid_probe.py lastsel $c 3 > /dev/null 2>&1
grep k_select_bytes gpurun_out/launches_lastsel_$c.csv | tail -2
done
