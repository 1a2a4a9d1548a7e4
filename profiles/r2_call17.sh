#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -q -x -k "ping_pong" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest17.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest17.log | cut -c1-300 | head
: > gpurun_out/r2_pp17.log
TC_TIMING=1 PP=1 timeout 60 python profiles/prof_fwd.py 2048 pr4 f32 400 2>&1 | tail -4 | cut -c1-400 >> gpurun_out/r2_pp17.log
for cfg in "512 400" "37888 400" "65536 400"; do
  set -- $cfg
  for pp in 1 0; do
    echo "== B=$1 n_out=$2 PP=$pp" >> gpurun_out/r2_pp17.log
    PP=$pp POOL=1 timeout 45 python profiles/prof_fwd.py $1 pr4 f32 $2 2>&1 | tail -1 | cut -c1-130 >> gpurun_out/r2_pp17.log
  done
done
cat gpurun_out/r2_pp17.log
