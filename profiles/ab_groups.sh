#!/bin/bash
# Same-box A/B of the number of column groups (IKR_TC_GROUPS) on the bench, the backward and the train leg.
mkdir -p gpurun_out
out=gpurun_out/ab_groups.log
: > $out
for rep in 1 2; do
  for g in 2 3; do
    echo "== bench groups=$g rep=$rep" >> $out
    IKR_TC_GROUPS=$g timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> $out
    echo "== bwd groups=$g rep=$rep" >> $out
    IKR_TC_GROUPS=$g timeout 300 python profiles/prof_bwd.py 18944 pr4 f32 400 2>&1 | tail -1 >> $out
  done
done
cat $out
