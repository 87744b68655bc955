#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_gpu_index.py tests/test_gpu_group.py tests/test_gpu_fullsize.py tests/test_cli.py -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
{
for b in XL L S; do python tools/grid_probe.py grid $b 7 2>&1; done
} > gpurun_out/ab8.txt 2>&1
cat gpurun_out/ab8.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_grid_xl_finlist.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
grep -E "k_grid_fin|k_grid_emit" gpurun_out/launches_grid_xl_finlist.csv | grep gpu__time | awk -F'","' '{print $5, $(NF)}'
