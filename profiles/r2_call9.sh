#!/bin/bash
mkdir -p gpurun_out
for sp in fp16x2 bf16x3; do TC_TIMING=1 TC_SPLIT=$sp timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -3 | cut -c1-330; done > gpurun_out/r2_timing9.log 2>&1; cat gpurun_out/r2_timing9.log
timeout 600 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | tail -1 | cut -c1-330 > gpurun_out/r2_bwd9.log; cat gpurun_out/r2_bwd9.log
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest9.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |d2 regression" gpurun_out/r2_pytest9.log | cut -c1-400 | head -20
timeout 900 python bench.py --steps 5 --warmup 3 --legs forward --no-cpu-baseline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc $?"; cut -c1-330 gpurun_out/r2_bench9.json
