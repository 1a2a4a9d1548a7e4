#!/bin/bash
mkdir -p gpurun_out
s=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2_bench27.json 2> gpurun_out/r2_bench27.err; echo "bench rc $? in $(( $(date +%s) - s )) s"
s=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench27_ref.json 2> gpurun_out/r2_bench27_ref.err; echo "ref rc $? in $(( $(date +%s) - s )) s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench27.json').read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'cpu_baseline', 'clocks', 'gpu_launches'):
    print(k, d.get(k))
for k in ('train', 'sweep', 'train1m'):
    v = d.get(k)
    print(k, json.dumps(v)[:700] if v else None)
r = json.loads(open('gpurun_out/r2_bench27_ref.json').read().strip().splitlines()[-1])
print('ref', r.get('value'), r.get('cpu_baseline'), json.dumps(r.get('train'))[:300])
PY
