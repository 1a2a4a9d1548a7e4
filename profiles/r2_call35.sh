#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest35.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest35.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest35.log | cut -c1-300 | head -20
timeout 600 python profiles/adjoint_products_accuracy.py > gpurun_out/r2_adjoint_products_accuracy.log 2>&1; tail -6 gpurun_out/r2_adjoint_products_accuracy.log | cut -c1-200
timeout 1500 python bench.py > gpurun_out/r2_bench35.json 2> gpurun_out/r2_bench35.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench35.json').read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'e2e', 'clocks', 'gpu_launches'):
    print(k, d.get(k))
for k in ('train', 'sweep', 'train1m'):
    v = d.get(k) or {}
    print(k, {q: v.get(q) for q in ('value', 'ms_per_step', 'makespan_ms', 'ideal_ms', 'compute_ms', 'allreduce_ms', 'e2e', 'roofline')})
PY
