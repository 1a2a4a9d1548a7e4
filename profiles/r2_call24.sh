#!/bin/bash
mkdir -p gpurun_out
for lib in head cur head cur; do
  if [ $lib = head ]; then export IKR_B200_LIB=/root/repo/build/libikr_head.so; else unset IKR_B200_LIB; fi
  echo "== $lib tile"; REPS=6 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
done
for lib in head cur; do
  if [ $lib = head ]; then export IKR_B200_LIB=/root/repo/build/libikr_head.so; else unset IKR_B200_LIB; fi
  echo "== $lib pool"; REPS=5 timeout 300 python profiles/prof_fwd.py 65536 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
