#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_pp11.log
for B in 2048 37888 38036 40000 65536; do
  echo "== B=$B" >> gpurun_out/r2_pp11.log
  PP=1 timeout 60 python profiles/prof_fwd.py $B pr4 f32 40 2>&1 | tail -1 | cut -c1-160 >> gpurun_out/r2_pp11.log; echo "rc $?" >> gpurun_out/r2_pp11.log
done
cat gpurun_out/r2_pp11.log
