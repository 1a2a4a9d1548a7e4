#!/bin/bash
for G in 3 2 3 2; do
echo "== G=$G 65536 pr4"; TC_GROUPS=$G REPS=3 timeout 300 python profiles/prof_bwd.py 65536 pr4 f32 200 d1 2>&1 | tail -2 | cut -c40-175
done
