#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -q -x -k "ping_pong" > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest10.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest10.log | cut -c1-300 | head -20
for pp in 0 1; do PP=$pp POOL=1 timeout 200 python profiles/prof_fwd.py 65536 pr4 f32 400 2>&1 | tail -1 | cut -c1-200; done > gpurun_out/r2_pp10.log 2>&1; cat gpurun_out/r2_pp10.log
