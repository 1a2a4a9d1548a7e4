#!/bin/bash
# ncu captures of round 1 (run under gpurun, one GPU): forward, adjoint and weight-gradient kernels
set -x
python profiles/prof_fwd.py 23680 pr4 f32 400 > gpurun_out/plain_fwd_v3.log 2>&1 || exit 1
python profiles/prof_bwd.py 23680 pr4 f32 200 d1 > gpurun_out/plain_bwd_v1.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ikr_forward -s 1 -c 1 -f -o gpurun_out/fwd_r1_v3 python profiles/prof_fwd.py 23680 pr4 f32 400 > gpurun_out/ncu_fwd_v3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ikr_adjoint -s 3 -c 1 -f -o gpurun_out/adj_r1_v1 python profiles/prof_bwd.py 23680 pr4 f32 200 d1 > gpurun_out/ncu_adj_v1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ikr_wgrad -s 3 -c 1 -f -o gpurun_out/wgrad_r1_v1 python profiles/prof_bwd.py 23680 pr4 f32 200 d1 > gpurun_out/ncu_wgrad_v1.log 2>&1
tail -1 gpurun_out/plain_fwd_v3.log gpurun_out/plain_bwd_v1.log
