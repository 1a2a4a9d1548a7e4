#!/bin/bash
# Same-box A/B of the shipped library against another build (build/ikr_prev.so): tile forward + bench.
mkdir -p gpurun_out
out=gpurun_out/ab_lib.log
: > $out
for rep in 1 2; do
  for v in "" build/ikr_prev.so; do
    echo "== fwd lib=${v:-default} rep=$rep" >> $out
    IKR_B200_LIB=${v:+$PWD/$v} timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -1 | cut -c1-120 >> $out
    echo "== bench lib=${v:-default} rep=$rep" >> $out
    IKR_B200_LIB=${v:+$PWD/$v} timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --train-batch 0 2>&1 | tail -1 | cut -c1-300 >> $out
  done
done
cat $out
