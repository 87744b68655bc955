#!/bin/bash
# the reference's experiment runner at 1/8 of the real sizes, GPU and CPU columns, 5 runs (run under gpurun)
mkdir -p gpurun_out
df -h /dev/shm | tail -1; free -g | head -2; nproc
python tools/run_query_experiments.py --generate /dev/shm/pcq --scale 0.125 2>&1 | tail -1
du -sh /dev/shm/pcq
{
echo "# tools/run_query_experiments.py --scale 0.125 (navvis 7.0 M points; doc 64 x 3.9 M; ca13 64 x 5.1 M), --runs 5, one B200 box, $(nproc) host cores"
echo "# name;mean;median;stddev [s] — the reference's line format (run_query_experiments.rs:287-304); <name>_cpu = the same command line on oracle/query_ref"
timeout 1500 python tools/run_query_experiments.py --input /dev/shm/pcq --runs 5
echo "# exit status $?"
} > gpurun_out/query_experiments_scale0125.txt 2>&1
tail -5 gpurun_out/query_experiments_scale0125.txt; wc -l gpurun_out/query_experiments_scale0125.txt
rm -rf /dev/shm/pcq
