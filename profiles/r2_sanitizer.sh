#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash profiles/r2_sanitizer.sh memcheck|racecheck|synccheck
tool=$1
mkdir -p gpurun_out
timeout 300 python profiles/all_paths_small.py > gpurun_out/r2_all_paths_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_all_paths_plain.log; exit 1; }
tail -3 gpurun_out/r2_all_paths_plain.log
timeout 2400 compute-sanitizer --tool $tool --print-limit 20 python profiles/all_paths_small.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "sanitizer rc $?" >> gpurun_out/r2_sanitizer_$tool.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer rc|done|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | sort | uniq -c | head -20
