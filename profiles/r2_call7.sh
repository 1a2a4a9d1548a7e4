#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/diag_status.py fp16x2 2>&1 | cut -c1-300 > gpurun_out/r2_diag_status2.log; cat gpurun_out/r2_diag_status2.log
for i in 1 2; do for sp in fp16x2 bf16x3; do TC_SPLIT=$sp timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -1 | cut -c1-120; done; done > gpurun_out/r2_fwd_split2.log 2>&1
cat gpurun_out/r2_fwd_split2.log
timeout 2400 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_forward.py tests/test_gpu_backward.py tests/test_gpu_parity_r2.py -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest7.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2_pytest7.log | cut -c1-250 | head -40
