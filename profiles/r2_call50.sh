#!/bin/bash
mkdir -p gpurun_out
s=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 > gpurun_out/r2_bench50_n4.json 2> gpurun_out/r2_bench50_n4.err; echo "bench rc $? in $(( $(date +%s) - s )) s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench50_n4.json').read().strip().splitlines()[-1])
for k in ('value', 'n_gpus', 'ms_per_step', 'e2e', 'clocks', 'gpu_launches'):
    print(k, d.get(k))
for k in ('train', 'sweep', 'train1m'):
    v = d.get(k) or {}
    print(k, {q: v.get(q) for q in ('value', 'ms_per_step', 'makespan_ms', 'ideal_ms', 'makespan_over_ideal', 'compute_ms', 'allreduce_ms', 'optimizer_ms', 'per_rank_ms')})
PY
tail -3 gpurun_out/r2_bench50_n4.err
