#!/bin/bash
mkdir -p gpurun_out
TC_TIMING=1 PP=1 timeout 100 python profiles/prof_fwd.py 2048 pr4 f32 400 2>&1 | tail -8 | cut -c1-420 > gpurun_out/r2_pp14.log; cat gpurun_out/r2_pp14.log
