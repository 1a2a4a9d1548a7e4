#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_pp12.log
for cfg in "512 400" "2048 400" "37888 100" "37888 400" "65536 100" "65536 200"; do
  set -- $cfg
  echo "== B=$1 n_out=$2" >> gpurun_out/r2_pp12.log
  PP=1 timeout 45 python profiles/prof_fwd.py $1 pr4 f32 $2 2>&1 | tail -1 | cut -c1-160 >> gpurun_out/r2_pp12.log; echo "rc $?" >> gpurun_out/r2_pp12.log
done
cat gpurun_out/r2_pp12.log
