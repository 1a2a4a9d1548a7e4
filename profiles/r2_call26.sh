#!/bin/bash
mkdir -p gpurun_out
for lib in head cur d1 head cur; do
  if [ $lib = cur ]; then unset IKR_B200_LIB; else export IKR_B200_LIB=/root/repo/build/libikr_$lib.so; fi
  echo "== $lib tile"; REPS=6 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
done
for lib in head cur; do
  if [ $lib = cur ]; then unset IKR_B200_LIB; else export IKR_B200_LIB=/root/repo/build/libikr_$lib.so; fi
  echo "== $lib pool"; REPS=5 timeout 300 python profiles/prof_fwd.py 65536 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
  echo "== $lib bwd"; REPS=4 timeout 300 python profiles/prof_bwd.py 2>&1 | tail -4 | tr '\n' ' '; echo
done
unset IKR_B200_LIB
timeout 1200 python -m pytest tests/test_gpu_tensor_core.py -m gpu -q 2>&1 | tail -3
