#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/adjoint_products_accuracy.py 2>&1 | tail -12
for A in six dz3 all3; do
  echo "== ADJ=$A 18944 pr4"; ADJ=$A REPS=4 timeout 300 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | tail -2 | cut -c1-260
done
for A in six dz3 all3; do
  echo "== ADJ=$A d2 staircase 4096"; ADJ=$A REPS=3 timeout 300 python profiles/prof_bwd.py 2>&1 | tail -1 | cut -c1-260
done
