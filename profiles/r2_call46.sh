#!/bin/bash
for O in "" 1 "" 1; do
echo "== OVL=$O 65536 pr4"; OVL=$O REPS=3 timeout 300 python profiles/prof_bwd.py 65536 pr4 f32 200 d1 2>&1 | tail -2 | cut -c40-175
done
for O in "" 1; do
echo "== OVL=$O 18944 pr4"; OVL=$O REPS=3 timeout 300 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | tail -2 | cut -c40-175
done
