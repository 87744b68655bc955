#!/bin/bash
mkdir -p gpurun_out
python bench.py --files 4 --pts-per-file 2000000 --steps 3 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; rc=$?; echo "small bench rc=$rc"
tail -c 1500 gpurun_out/bench_small.err
[ $rc -ne 0 ] && exit 1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_small.json').read().strip().splitlines()[-1])
print('small e2e', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='note'}) for k,v in d['e2e'].items() if k!='api'})
PY
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'])
print('e2e', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='note'}) for k,v in d['e2e'].items() if k!='api'})
print('cpu', d['cpu_baseline'])
print('density', json.dumps(d['density'])[:1500])
for r in d['extra']: print(r['config'], r['query'], r['collector'], round(r['ms'],3), round(r['frac'],3), r.get('finalize_ms_host_clock'))
PY
