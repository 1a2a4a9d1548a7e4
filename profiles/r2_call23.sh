#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest23.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest23.log | cut -c1-300 | head -20
REPS=5 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
timeout 900 python bench.py --steps 5 --warmup 3 --legs forward --no-cpu-baseline > gpurun_out/r2_bench23.json 2> gpurun_out/r2_bench23.err; echo "bench rc $?"; cut -c100-260 gpurun_out/r2_bench23.json
