#!/bin/bash
for G in 3 2 3 2; do
echo "== G=$G d2 4096"; TC_GROUPS=$G REPS=3 timeout 300 python profiles/prof_bwd.py 2>&1 | tail -2 | cut -c40-200
done
