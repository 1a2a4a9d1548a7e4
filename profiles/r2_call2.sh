#!/bin/bash
# round 2, call 2: the new parity tests (stand-ins, replay traces, TC backward, long inputs, 92-row KAT)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity_r2.py -q -s --durations=25 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest2.log
tail -40 gpurun_out/r2_pytest2.log
