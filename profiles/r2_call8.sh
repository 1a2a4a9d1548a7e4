#!/bin/bash
mkdir -p gpurun_out
for sp in fp16x2 bf16x3; do TC_TIMING=1 TC_SPLIT=$sp timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -3 | cut -c1-420; done > gpurun_out/r2_timing8.log 2>&1; cat gpurun_out/r2_timing8.log
timeout 600 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | tail -1 | cut -c1-400 > gpurun_out/r2_bwd8.log; cat gpurun_out/r2_bwd8.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc $?"; tail -2 gpurun_out/r2_bench8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench8.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','executed_tensor_tflops','frac_executed')})
t=d['train']; print('train',t['value'],t['ms_per_step'],t['e2e']['value'],t['cpu_baseline']['value'])
print('sweep',d['sweep']['makespan_ms'],'1m',d['train1m']['value'],d['train1m']['ms_per_step'])
PY
