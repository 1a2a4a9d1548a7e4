#!/bin/bash
# same per-tile work on 37 / 74 / 148 SMs: if the time per tile falls with fewer SMs the kernel is power-bound
for B in 4736 9472 18944; do
  echo "== B=$B"; (while true; do nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits; sleep 0.05; done > /tmp/smi_$B.log) & SP=$!
  REPS=40 timeout 300 python profiles/prof_fwd.py $B pr4 f32 400 2>&1 | grep rep | tail -5 | tr '\n' ' '; echo
  kill $SP; sort -t, -k2 -n -r /tmp/smi_$B.log | head -3 | tr '\n' ';'; echo
done
