#!/bin/bash
# ncu captures of the tensor-core backward kernels (run under gpurun, one GPU)
set -x
python profiles/prof_bwd.py 18944 pr4 f32 200 d1 > gpurun_out/plain_bwd_tc.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_prof_bwd_tc.csv \
  python profiles/prof_bwd.py 18944 pr4 f32 200 d1 > gpurun_out/ncu_bwd_tc_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ikr_adjoint_tc -s 2 -c 1 -f -o gpurun_out/adj_r1_tc \
  python profiles/prof_bwd.py 18944 pr4 f32 200 d1 > gpurun_out/ncu_adj_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ikr_wgrad_tc -s 2 -c 1 -f -o gpurun_out/wgrad_r1_tc \
  python profiles/prof_bwd.py 18944 pr4 f32 200 d1 > gpurun_out/ncu_wgrad_tc.log 2>&1
ncu -i gpurun_out/adj_r1_tc.ncu-rep --page raw --csv > gpurun_out/adj_r1_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/wgrad_r1_tc.ncu-rep --page raw --csv > gpurun_out/wgrad_r1_tc_raw.csv 2>/dev/null
tail -1 gpurun_out/plain_bwd_tc.log
