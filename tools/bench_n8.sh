#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['roofline']['frac'])
print('e2e', d['e2e']['value'], d['e2e'].get('h2d_gbs_per_gpu'))
for k,q in d['density']['queries'].items(): print(k, q['ms_end_to_end_host_clock'], json.dumps(q['phases']), q['equals_oracle'])
PY
