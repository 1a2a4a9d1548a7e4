#!/bin/bash
# Same-box A/B of kernel variants built with csrc/build.py <out.so> -D... (selected via IKR_B200_LIB).
# Each variant runs twice, interleaved, so box-to-box and thermal drift cancel.
mkdir -p gpurun_out
out=gpurun_out/ab_variants.log
: > $out
for rep in 1 2; do
  for v in "" build/ikr_stag0.so build/ikr_stag100.so build/ikr_stag400.so; do
    echo "== fwd lib=${v:-default} rep=$rep" >> $out
    IKR_B200_LIB=${v:+$PWD/$v} timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -1 >> $out
  done
  for g in 2 3; do
    echo "== fwd groups=$g rep=$rep" >> $out
    IKR_TC_GROUPS=$g timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | tail -1 >> $out
  done
  for v in "" build/ikr_adjT.so; do
    echo "== bwd lib=${v:-default} rep=$rep" >> $out
    IKR_B200_LIB=${v:+$PWD/$v} timeout 300 python profiles/prof_bwd.py 18944 pr4 f32 400 2>&1 | tail -1 >> $out
  done
done
echo "== gradient accuracy with the truncated adjoint split" >> $out
IKR_B200_LIB=$PWD/build/ikr_adjT.so timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_tensor_core.py -q -x 2>&1 | tail -3 >> $out
cat $out
