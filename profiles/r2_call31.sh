#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "concurrent" 2>&1 | tail -15
for C in 1 0; do
  timeout 900 python bench.py --steps 2 --warmup 3 --legs forward,sweep --no-cpu-baseline --sweep-concurrency $C > gpurun_out/r2_bench31_c$C.json 2> gpurun_out/r2_bench31_c$C.err; echo "bench C=$C rc $?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench31_c$C.json').read().strip().splitlines()[-1])
s = d['sweep']
print({k: s[k] for k in ('value', 'makespan_ms', 'ideal_ms', 'makespan_over_ideal', 'per_rank_ms', 'sum_of_fit_ms', 'longest_fit_ms')})
print([(e['arch'], e['dtype'], round(e['ms']), e.get('error', '')[:60], e.get('status_bad'), round(e.get('loss', 0), 3)) for e in s['fits']])
PY
done
tail -5 gpurun_out/r2_bench31_c0.err
