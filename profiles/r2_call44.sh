#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest44.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest44.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest44.log | cut -c1-300 | head -20
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_bench44.json 2> gpurun_out/r2_bench44.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench44.json').read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'e2e', 'clocks', 'gpu_launches'):
    print(k, d.get(k))
print(d['config'])
for k in ('train', 'sweep', 'train1m'):
    v = d.get(k) or {}
    print(k, {q: v.get(q) for q in ('value', 'ms_per_step', 'makespan_ms', 'ideal_ms', 'compute_ms')})
print(json.dumps(d['roofline'])[:600])
PY
