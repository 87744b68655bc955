#!/bin/bash
# density subset of the GPU suite + the C4 probes (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_gpu_index.py tests/test_gpu_group.py tests/test_gpu_fullsize.py -m gpu -x -q -k "density or grid or group or fullsize" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
{ for b in XL L S; do python tools/grid_probe.py grid $b 7; done; } | tee gpurun_out/grid_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_grid_xl.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
grep -E "k_grid_scan" gpurun_out/launches_grid_xl.csv | awk -F'","' '{print $NF}' | tail -3
