#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --legs forward,train,train1m --no-cpu-baseline > gpurun_out/r2_bench38.json 2> gpurun_out/r2_bench38.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench38.json').read().strip().splitlines()[-1])
print('value', d['value'], d['clocks'])
for k in ('train', 'train1m'):
    v = d.get(k) or {}
    print(k, {q: v.get(q) for q in ('value', 'ms_per_step', 'compute_ms')})
PY
