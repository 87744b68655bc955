#!/bin/bash
# A/B: density pre-filter on/off x records per lane; kernel-only times of the LAST class select (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_gpu_index.py tests/test_cli.py tests/test_gpu_group.py -m gpu -x -q -k "density or grid or las_records or cli or group" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
{
for flt in 1 0; do for ppt in 2 4; do
  for b in XL L; do
    echo "filter=$flt ppt=$ppt"; PCQ_GRID_FILTER=$flt PCQ_GRID_PPT=$ppt python tools/grid_probe.py grid $b 7 2>&1
  done
done; done
echo "filter=1 ppt=2"; python tools/grid_probe.py grid S 7
} > gpurun_out/ab3.txt 2>&1
cat gpurun_out/ab3.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_grid_xl.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
for c in 19 6; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lastsel_$c.csv python tools/grid_probe.py lastsel $c 3 > /dev/null 2>&1
done
grep -c . gpurun_out/launches_*.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_grid_scan -s 2 -c 1 -f -o gpurun_out/gsf_xl python tools/grid_probe.py grid XL 2 > gpurun_out/ncu_gsf.log 2>&1
tail -2 gpurun_out/ncu_gsf.log
