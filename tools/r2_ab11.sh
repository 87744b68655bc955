#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_hypothesis.py -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
{
echo "owner path"; for c in 6 19 2; do python tools/grid_probe.py lastsel $c 7; done
for c in 6; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lastsel_$c.csv python tools/grid_probe.py lastsel $c 3 > /dev/null 2>&1
grep k_select_bytes gpurun_out/launches_lastsel_$c.csv | tail -2 | awk -F'","' '{print "kernel ns", $NF}'
done
cp adhoc-queries-pointclouds_b200/libpcq.so /tmp/libpcq_main.so
cp adhoc-queries-pointclouds_b200/csrc/build/abno/libpcq.so adhoc-queries-pointclouds_b200/libpcq.so
echo "binary search path"; for c in 6; do python tools/grid_probe.py lastsel $c 7; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lastsel_6_old.csv python tools/grid_probe.py lastsel 6 3 > /dev/null 2>&1
grep k_select_bytes gpurun_out/launches_lastsel_6_old.csv | tail -2 | awk -F'","' '{print "kernel ns", $NF}'
cp /tmp/libpcq_main.so adhoc-queries-pointclouds_b200/libpcq.so
} > gpurun_out/ab11.txt 2>&1
cat gpurun_out/ab11.txt
python tools/sweep.py --points 134217728 --only select --variants 2 --cases last:1,last:3 > gpurun_out/sweep_select_last.jsonl 2> gpurun_out/sweep_select.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_select_last.jsonl'):
    d=json.loads(l); print(d['layout'],d['format'],d['query'],round(d['ms'],3),round(d['frac_of_measured_peak'],3))
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_select_bytes -s 1 -c 1 -f -o gpurun_out/selb6 python tools/grid_probe.py lastsel 6 2 > gpurun_out/ncu_selb6.log 2>&1
tail -2 gpurun_out/ncu_selb6.log
