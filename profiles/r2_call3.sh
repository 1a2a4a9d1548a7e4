#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -k "rk4_window or 92_logged" > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest3.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc $?"
tail -3 gpurun_out/r2_pytest3.log; tail -5 gpurun_out/r2_bench3.err
