#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --cpu-seconds 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench rc $?"
tail -3 gpurun_out/r2_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'n',d['n_gpus'],'e2e',d['e2e']['value'])
t=d['train']; print('train',t['value'],t['ms_per_step'],t['e2e']['value'])
s=d['sweep']; print('sweep',s['makespan_ms'],s['ideal_ms'],s['makespan_over_ideal'],s['per_rank_ms'])
m=d['train1m']; print('1m',{k:m[k] for k in m if k!='workload'})
PY
