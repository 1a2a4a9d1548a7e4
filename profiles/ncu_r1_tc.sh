#!/bin/bash
# ncu captures of the tcgen05 forward kernel (run under gpurun, one GPU): plain run first, then the
# launch list of bench.py and one --set full capture of the dominant kernel.
set -x
python profiles/prof_fwd.py 18944 pr4 f32 400 > gpurun_out/plain_fwd_tc.log 2>&1 || exit 1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-batch 0 > gpurun_out/bench_plain_tc.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench_tc.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-batch 0 > gpurun_out/ncu_bench_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ikr_forward_tc -s 1 -c 1 -f -o gpurun_out/fwd_r1_tc \
  python profiles/prof_fwd.py 18944 pr4 f32 400 > gpurun_out/ncu_fwd_tc.log 2>&1
ncu -i gpurun_out/fwd_r1_tc.ncu-rep --page raw --csv > gpurun_out/fwd_r1_tc_raw.csv 2>/dev/null
tail -1 gpurun_out/plain_fwd_tc.log
