#!/bin/bash
mkdir -p gpurun_out
{
for mb in 0 1000; do for flt in 1 0; do
  for b in XL L; do
    echo "persist_mb=$mb filter=$flt"; PCQ_VERBOSE=1 PCQ_L2_PERSIST_MB=$mb PCQ_GRID_FILTER=$flt python tools/grid_probe.py grid $b 7 2>&1 | grep -v "^k_"
  done
done; done
} > gpurun_out/ab4.txt 2>&1
cat gpurun_out/ab4.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_grid_xl_persist.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
PCQ_L2_PERSIST_MB=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_grid_xl_nopersist.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
grep k_grid_scan gpurun_out/launches_grid_xl_persist.csv | tail -4
grep k_grid_scan gpurun_out/launches_grid_xl_nopersist.csv | tail -4
