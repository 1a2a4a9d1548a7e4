#!/bin/bash
mkdir -p gpurun_out
s=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 > gpurun_out/r2_bench32_n2.json 2> gpurun_out/r2_bench32_n2.err; echo "bench rc $? in $(( $(date +%s) - s )) s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench32_n2.json').read().strip().splitlines()[-1])
for k in ('value', 'n_gpus', 'ms_per_step', 'e2e', 'clocks', 'gpu_launches'):
    print(k, d.get(k))
for k in ('train', 'sweep', 'train1m'):
    v = d.get(k) or {}
    print(k, {q: v.get(q) for q in ('value', 'ms_per_step', 'makespan_ms', 'ideal_ms', 'makespan_over_ideal', 'compute_ms', 'allreduce_ms', 'per_rank_ms', 'sum_of_fit_ms', 'longest_fit_ms')})
print([(e['arch'], e['dtype'], round(e['ms']), e.get('error', '')[:50]) for e in d['sweep']['fits'] if e.get('error') or e['ms'] > 3000])
PY
tail -3 gpurun_out/r2_bench32_n2.err
