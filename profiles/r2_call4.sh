#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/mlp_accuracy.py > gpurun_out/r2_mlp_accuracy.log 2>&1; echo "acc rc $?"
cat gpurun_out/r2_mlp_accuracy.log | cut -c1-600
for sp in fp16x2 bf16x3; do TC_SPLIT=$sp timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -1 | cut -c1-200; done > gpurun_out/r2_fwd_split.log 2>&1
cat gpurun_out/r2_fwd_split.log
timeout 1500 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_forward.py tests/test_gpu_parity_r2.py -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest4.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest4.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --legs forward --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc $?"
cut -c1-400 gpurun_out/r2_bench4.json
