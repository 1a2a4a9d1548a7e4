#!/bin/bash
for A in six three; do
echo "== $A"; ADJ=$A REPS=5 timeout 300 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | grep -E "bwd " | tail -3 | cut -c60-200
done
echo "== three timing"; ADJ=three TC_TIMING=1 REPS=2 timeout 300 python profiles/prof_bwd.py 18944 pr4 f32 200 d1 2>&1 | grep -E "timing" | head -2 | cut -c1-300
echo "== three d2 4096"; ADJ=three REPS=3 timeout 300 python profiles/prof_bwd.py 2>&1 | tail -1 | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_backward.py tests/test_gpu_tensor_core.py -m gpu -q -k "backward or gradient or autograd or fused" 2>&1 | tail -3
