#!/bin/bash
for G in 2 3; do
  echo "== G=$G tile"; TC_GROUPS=$G REPS=5 timeout 300 python profiles/prof_fwd.py 18944 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
  echo "== G=$G pool"; TC_GROUPS=$G REPS=5 timeout 300 python profiles/prof_fwd.py 65536 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
done
echo "== PP pool"; PP=1 REPS=5 timeout 300 python profiles/prof_fwd.py 65536 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
echo "== PP 37888"; PP=1 REPS=5 timeout 300 python profiles/prof_fwd.py 37888 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
echo "== pool 37888"; REPS=5 timeout 300 python profiles/prof_fwd.py 37888 pr4 f32 400 2>&1 | grep rep | tr '\n' ' '; echo
