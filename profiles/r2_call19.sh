#!/bin/bash
mkdir -p gpurun_out
# (1) every kernel family under the debug build with bounded mbarrier waits (compute-sanitizer is closed on this pool)
IKR_B200_LIB=$PWD/build/libikr_dbg.so timeout 300 python profiles/all_paths_small.py > gpurun_out/r2_all_paths_dbg.log 2>&1; echo "dbg rc $?" >> gpurun_out/r2_all_paths_dbg.log
tail -4 gpurun_out/r2_all_paths_dbg.log
# (2) A/B: column groups and stagger with the fp16x2 split
: > gpurun_out/r2_ab19.log
for rep in 1 2; do
 for v in default st0 st100 g2; do
  lib=""; grp=0
  case $v in st0) lib=$PWD/build/libikr_st0.so;; st100) lib=$PWD/build/libikr_st100.so;; g2) grp=2;; esac
  for cfg in "18944 0" "65536 1"; do
    set -- $cfg
    echo -n "$v B=$1: " >> gpurun_out/r2_ab19.log
    IKR_B200_LIB=$lib TC_GROUPS=$grp POOL=$2 timeout 100 python profiles/prof_fwd.py $1 pr4 f32 400 2>&1 | tail -1 | cut -c20-110 >> gpurun_out/r2_ab19.log
  done
 done
done
cat gpurun_out/r2_ab19.log
