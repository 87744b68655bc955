#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "last or class or select" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
{
for c in 6 19 2; do python tools/grid_probe.py lastsel $c 7; done
} > gpurun_out/ab14.txt 2>&1
cat gpurun_out/ab14.txt
for c in 6 19 2; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_lastsel_$c.csv python tools/grid_probe.py lastsel $c 3 > /dev/null 2>&1
grep k_select_bytes gpurun_out/launches_lastsel_$c.csv | tail -3 | awk -F'","' '{print $(NF-2), $NF}'
done
