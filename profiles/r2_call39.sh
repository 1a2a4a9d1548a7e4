#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_backward.py tests/test_gpu_tensor_core.py tests/test_gpu_regression.py -m gpu -q -k "backward or gradient or autograd or fused or regression or concurrent" 2>&1 | tail -3
timeout 1500 python bench.py --legs forward,train,train1m --no-cpu-baseline > gpurun_out/r2_bench39.json 2> gpurun_out/r2_bench39.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench39.json').read().strip().splitlines()[-1])
print('value', d['value'], d['clocks'])
for k in ('train', 'train1m'):
    v = d.get(k) or {}
    print(k, {q: v.get(q) for q in ('value', 'ms_per_step', 'compute_ms')})
PY
