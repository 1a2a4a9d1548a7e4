#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 250 > gpurun_out/r2_clocks20.csv &
SMI=$!
REPS=10 timeout 120 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep "rep:" | tr '\n' ' ' > gpurun_out/r2_reps20.log; echo >> gpurun_out/r2_reps20.log
REPS=10 TC_SPLIT=bf16x3 timeout 120 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep "rep:" | tr '\n' ' ' >> gpurun_out/r2_reps20.log; echo >> gpurun_out/r2_reps20.log
REPS=6 POOL=1 timeout 120 python profiles/prof_fwd.py 65536 pr4 f32 400 2>&1 | grep "rep:" | tr '\n' ' ' >> gpurun_out/r2_reps20.log; echo >> gpurun_out/r2_reps20.log
kill $SMI
cat gpurun_out/r2_reps20.log; sort gpurun_out/r2_clocks20.csv | uniq -c | sort -rn | head -8
