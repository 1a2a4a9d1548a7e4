#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/diag_status.py fp16x2 > gpurun_out/r2_diag_status.log 2>&1; cat gpurun_out/r2_diag_status.log | cut -c1-400
TC_TIMING=1 TC_SPLIT=fp16x2 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -3 | cut -c1-400 > gpurun_out/r2_timing_f16.log; cat gpurun_out/r2_timing_f16.log
TC_TIMING=1 TC_SPLIT=bf16x3 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -3 | cut -c1-400 > gpurun_out/r2_timing_bf16.log; cat gpurun_out/r2_timing_bf16.log
