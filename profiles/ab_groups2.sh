#!/bin/bash
# Same-box A/B of 2 vs 3 column groups on the remaining kernels (regression, sweep widths, train leg)
# and the GPU test suite with 2 groups.
mkdir -p gpurun_out
out=gpurun_out/ab_groups2.log
: > $out
for rep in 1 2; do
  for g in 2 3; do
    echo "== regression groups=$g rep=$rep" >> $out
    IKR_TC_GROUPS=$g timeout 300 python profiles/prof_regression.py 214000 2>&1 | tail -2 >> $out
    echo "== fwd 4736 x pr4 (quarter wave) groups=$g rep=$rep" >> $out
    IKR_TC_GROUPS=$g timeout 300 python profiles/prof_fwd.py 4736 pr4 f32 400 2>&1 | tail -1 | cut -c1-140 >> $out
    echo "== fwd 28416 x pr4 (1.5 waves) groups=$g rep=$rep" >> $out
    IKR_TC_GROUPS=$g timeout 300 python profiles/prof_fwd.py 28416 pr4 f32 400 2>&1 | tail -1 | cut -c1-140 >> $out
  done
done
echo "== gpu tests with 2 groups" >> $out
IKR_TC_GROUPS=2 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> $out
cat $out
