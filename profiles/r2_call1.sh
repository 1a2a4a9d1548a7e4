#!/bin/bash
# round 2, first GPU call: the GPU suite on the hygiene changes (shared cudart, descriptor bits) and
# the in-kernel phase clocks that size the ping-pong design
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest1.log
TC_TIMING=1 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 > gpurun_out/r2_timing_fwd_tile.log 2>&1
TC_TIMING=1 POOL=1 timeout 300 python profiles/prof_fwd.py 37888 pr4 f32 400 > gpurun_out/r2_timing_fwd_pool.log 2>&1
TC_TIMING=1 timeout 600 python profiles/prof_bwd.py 4096 staircase f32 1000 d2 > gpurun_out/r2_timing_bwd_4096.log 2>&1
tail -3 gpurun_out/r2_pytest1.log; tail -4 gpurun_out/r2_timing_fwd_tile.log
