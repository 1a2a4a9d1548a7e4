#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_backward.py tests/test_gpu_regression.py -q -x > gpurun_out/r2_pytest22.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest22.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest22.log | cut -c1-300 | head
REPS=4 timeout 600 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | tail -2 | cut -c1-330
timeout 900 python bench.py --steps 3 --warmup 2 --legs forward,train,train1m --no-cpu-baseline > gpurun_out/r2_bench22.json 2> gpurun_out/r2_bench22.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench22.json').read().strip().splitlines()[-1])
print('value',d['value'],d['ms_per_step'])
t=d['train']; print('train',t['value'],t['ms_per_step'],t['e2e']['value'])
m=d['train1m']; print('1m',m['value'],m['ms_per_step'])
PY
