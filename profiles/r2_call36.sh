#!/bin/bash
for A in six three; do
echo "== $A"; ADJ=$A TC_TIMING=1 REPS=2 timeout 300 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | grep -E "timing|bwd " | tail -8 | cut -c1-420
done
