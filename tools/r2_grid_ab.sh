#!/bin/bash
# A/B of the density insert kernels + one ncu capture (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
{
for k in warp ring; do
  for b in XL L S; do
    echo "kernel=$k"; PCQ_GRID_KERNEL=$k python tools/grid_probe.py grid $b 7
  done
done
echo "kernel=warp fast_div=0"; PCQ_GRID_FAST_DIV=0 python tools/grid_probe.py grid XL 7
} > gpurun_out/grid_ab.txt 2>&1
cat gpurun_out/grid_ab.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_grid_warp -c 1 -f -o gpurun_out/gw_xl python tools/grid_probe.py grid XL 1 > gpurun_out/ncu_gw.log 2>&1
tail -2 gpurun_out/ncu_gw.log
