#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_backward.py -q -x > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest21.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest21.log | cut -c1-300 | head
timeout 900 python bench.py --steps 3 --warmup 2 --legs forward,train --no-cpu-baseline > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench21.json').read().strip().splitlines()[-1])
print('value',d['value'],d['ms_per_step'])
t=d['train']; print('train',t['value'],t['ms_per_step'],t['e2e']['value'],t['clocks'])
PY
