#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_gpu_index.py tests/test_gpu_group.py -m gpu -x -q -k "density or grid or group" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_grid_xl_finlist.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
grep -E "k_grid_fin|k_grid_emit" gpurun_out/launches_grid_xl_finlist.csv | grep gpu__time | awk -F'","' '{print substr($5,1,40), $(NF)}'
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'])
print('e2e', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='note'}) for k,v in d['e2e'].items() if k!='api'})
print('cpu', d['cpu_baseline'])
for r in d['density']['queries'] if 'queries' in d['density'] else []: print(r)
for r in d['extra']: print(r['config'], r['query'], r['collector'], round(r['ms'],3), round(r['frac'],3), r.get('finalize_ms_host_clock'))
PY
