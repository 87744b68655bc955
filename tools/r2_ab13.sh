#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_hypothesis.py tests/test_gpu_index.py -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
{
for c in 6 19 2; do python tools/grid_probe.py lastsel $c 7; done
} > gpurun_out/ab13.txt 2>&1
cat gpurun_out/ab13.txt
for c in 6 2; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_lastsel_$c.csv python tools/grid_probe.py lastsel $c 3 > /dev/null 2>&1
grep k_select_bytes gpurun_out/launches_lastsel_$c.csv | tail -3 | awk -F'","' '{print $(NF-2), $NF}'
done
python tools/sweep.py --points 134217728 --only select --variants 2 --cases last:1,last:3 > gpurun_out/sweep_select_last.jsonl 2> gpurun_out/sweep_select.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_select_last.jsonl'):
    d=json.loads(l); print(d['layout'],d['format'],d['query'],round(d['ms'],3),round(d['frac_of_measured_peak'],3))
PY
