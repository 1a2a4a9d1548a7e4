#!/bin/bash
mkdir -p gpurun_out
for lib in head v1 v2 cur head v1 v2 cur; do
  if [ $lib = cur ]; then unset IKR_B200_LIB; else export IKR_B200_LIB=/root/repo/build/libikr_$lib.so; fi
  echo "== $lib tile"; REPS=6 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
done
