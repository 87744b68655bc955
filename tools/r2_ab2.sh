#!/bin/bash
# A/B: density insert with 4 vs 2 records per lane; LAST class select with pipelined class-byte loads (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_gpu_index.py -m gpu -x -q -k "density or grid or class or select or last" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
{
for ppt in 4 2; do
  for b in XL L S; do
    echo "ppt=$ppt"; PCQ_VERBOSE=1 PCQ_GRID_PPT=$ppt python tools/grid_probe.py grid $b 7 2>&1 | grep -v "^k_sel"
  done
done
for c in 19 6 2; do echo "search_max=512"; python tools/grid_probe.py lastsel $c 7; done
cp adhoc-queries-pointclouds_b200/libpcq.so /tmp/libpcq_main.so
cp adhoc-queries-pointclouds_b200/csrc/build/ab64/libpcq.so adhoc-queries-pointclouds_b200/libpcq.so
for c in 19 6 2; do echo "search_max=64"; python tools/grid_probe.py lastsel $c 7; done
cp /tmp/libpcq_main.so adhoc-queries-pointclouds_b200/libpcq.so
} > gpurun_out/ab2.txt 2>&1
cat gpurun_out/ab2.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_grid_scan -s 2 -c 1 -f -o gpurun_out/gs4_xl python tools/grid_probe.py grid XL 2 > gpurun_out/ncu_gs4.log 2>&1
tail -2 gpurun_out/ncu_gs4.log
