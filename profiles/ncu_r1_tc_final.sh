#!/bin/bash
# ncu captures of the final (unit-ring) tcgen05 forward kernels, run under gpurun on one GPU:
# plain runs first, then the launch list of bench.py, one --set full capture of the lane-pool kernel
# inside the bench step and one of the tile kernel on the uniform 18,944 x pr4 workload.
set -x
python profiles/prof_fwd.py 18944 pr4 f32 400 > gpurun_out/plain_fwd_tc_final.log 2>&1 || exit 1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-batch 0 > gpurun_out/bench_plain_tc_final.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench_tc_final.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-batch 0 > gpurun_out/ncu_bench_tc_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ikr_forward_tc_pool -s 1 -c 1 -f -o gpurun_out/fwd_r1_tc_pool \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-batch 0 > gpurun_out/ncu_pool_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ikr_forward_tc_kernel -s 1 -c 1 -f -o gpurun_out/fwd_r1_tc_final \
  python profiles/prof_fwd.py 18944 pr4 f32 400 > gpurun_out/ncu_fwd_tc_final.log 2>&1
ncu -i gpurun_out/fwd_r1_tc_pool.ncu-rep --page raw --csv > gpurun_out/fwd_r1_tc_pool_raw.csv 2>/dev/null
ncu -i gpurun_out/fwd_r1_tc_final.ncu-rep --page raw --csv > gpurun_out/fwd_r1_tc_final_raw.csv 2>/dev/null
ncu -i gpurun_out/fwd_r1_tc_final.ncu-rep --page source --csv > gpurun_out/fwd_r1_tc_final_source.csv 2>/dev/null
# the .ncu-rep files (2 x ~45 MB) exceed what gpurun copies back: keep the CSV exports only
rm -f gpurun_out/fwd_r1_tc_pool.ncu-rep gpurun_out/fwd_r1_tc_final.ncu-rep
tail -1 gpurun_out/plain_fwd_tc_final.log
du -sh gpurun_out
