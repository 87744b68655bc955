#!/bin/bash
# full GPU suite, select probes, full N = 1 bench (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
{ for c in 6 19 2; do python tools/grid_probe.py lastsel $c 7; done; } > gpurun_out/lastsel.txt 2>&1; cat gpurun_out/lastsel.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'])
e=d['e2e']; print('e2e', e['value'], 'last', e['as_last_columns']['value'])
print('indexed', json.dumps(e['strip_ordered_indexed_repeat'])[:1800])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k,q in d['density']['queries'].items(): print(k, q['ms_end_to_end_host_clock'], q['phases']['scan_ms'], q['phases']['finalize_ms'], q['equals_oracle'])
for r in d['extra']: print(r['config'], r['query'], r['collector'], round(r['ms'],3), round(r['frac'],3), r.get('finalize_ms_host_clock'))
PY
# ncu evidence of the same commands (a number printed under ncu is never a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-extra --no-density > gpurun_out/ncu_bench.log 2>&1
grep -c k_scan_staged gpurun_out/launches_bench_default.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_grid_scan -s 2 -c 1 -f -o gpurun_out/grid_scan_xl python tools/grid_probe.py grid XL 2 > gpurun_out/ncu_grid.log 2>&1
tail -1 gpurun_out/ncu_grid.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_grid_xl.csv python tools/grid_probe.py grid XL 3 > /dev/null 2>&1
grep -E "k_grid" gpurun_out/launches_grid_xl.csv | awk -F'","' '{print substr($5,1,44), $(NF-2), $NF}' | tail -16
