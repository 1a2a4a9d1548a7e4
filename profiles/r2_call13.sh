#!/bin/bash
mkdir -p gpurun_out
export IKR_B200_LIB=$PWD/build/libikr_dbg.so
PP=1 timeout 100 python profiles/prof_fwd.py 37888 pr4 f32 400 > gpurun_out/r2_pp13.log 2>&1; echo "rc $?" >> gpurun_out/r2_pp13.log
grep -E "mbar timeout|rc |Error|error" gpurun_out/r2_pp13.log | sort | uniq -c | sort -rn | head -40 | cut -c1-200
