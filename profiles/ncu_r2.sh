#!/bin/bash
# Round-2 ncu captures (one gpurun call, one GPU): every command first runs plain (must exit 0), then
# under ncu.  Launch list of the bench's forward leg, one --set full capture of the lane-pool forward
# kernel (fp16x2 split) inside the bench step, one of the tile kernel on the uniform 18,944 x pr4
# workload, one each of the adjoint and weight-gradient kernels.  CSV exports only travel back.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --legs forward"
F="python profiles/prof_fwd.py 18944 pr4 f32 400"
W="python profiles/prof_bwd.py 18944 pr4 f32 200 d1"
$B > gpurun_out/r2_ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_bench_list.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ikr_forward_tc_pool -s 1 -c 1 -f -o gpurun_out/r2_fwd_pool $B > gpurun_out/r2_ncu_pool.log 2>&1
$F > gpurun_out/r2_ncu_plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ikr_forward_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_fwd_tile $F > gpurun_out/r2_ncu_tile.log 2>&1
$W > gpurun_out/r2_ncu_plain_bwd.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bwd.csv $W > gpurun_out/r2_ncu_bwd_list.log 2>&1
$W > /dev/null 2>&1 &&
ncu --set full --clock-control none -k regex:ikr_adjoint_tc -s 2 -c 1 -f -o gpurun_out/r2_adj $W > gpurun_out/r2_ncu_adj.log 2>&1
$W > /dev/null 2>&1 &&
ncu --set full --clock-control none -k regex:ikr_wgrad_tc -s 2 -c 1 -f -o gpurun_out/r2_wgrad $W > gpurun_out/r2_ncu_wgrad.log 2>&1
for r in r2_fwd_pool r2_fwd_tile r2_adj r2_wgrad; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_ncu_raw.csv 2>/dev/null
done
ncu -i gpurun_out/r2_fwd_tile.ncu-rep --page source --csv > gpurun_out/r2_fwd_tile_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
tail -1 gpurun_out/r2_ncu_plain_fwd.log | cut -c1-160; tail -1 gpurun_out/r2_ncu_plain_bwd.log | cut -c1-250
ls -la gpurun_out/r2_*ncu_raw.csv gpurun_out/r2_launches_*.csv; tail -3 gpurun_out/r2_ncu_pool.log | cut -c1-200
