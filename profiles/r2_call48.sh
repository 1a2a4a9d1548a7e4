#!/bin/bash
for lib in cur st0 cur st0; do
  if [ $lib = cur ]; then unset IKR_B200_LIB; else export IKR_B200_LIB=/root/repo/build/libikr_$lib.so; fi
  echo "== $lib tile"; REPS=5 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
  echo "== $lib d2 4096"; REPS=3 timeout 300 python profiles/prof_bwd.py 2>&1 | tail -1 | cut -c40-170
done
