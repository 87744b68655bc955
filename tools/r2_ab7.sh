#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_gpu_index.py tests/test_gpu_group.py -m gpu -x -q -k "density or grid or group" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
{
for blk in 1 0; do
  for b in XL L S; do
    echo "blocked=$blk"; PCQ_GRID_BLOCKED=$blk python tools/grid_probe.py grid $b 7 2>&1
  done
done
} > gpurun_out/ab7.txt 2>&1
cat gpurun_out/ab7.txt
for blk in 1 0; do
PCQ_GRID_BLOCKED=$blk ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/launches_grid_xl_blk$blk.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
grep -E "k_grid_scan|k_grid_final|k_grid_emit" gpurun_out/launches_grid_xl_blk$blk.csv | tail -16 | cut -d, -f5,13-
done
