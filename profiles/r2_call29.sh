#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest29.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest29.log
grep -E "^(FAILED|ERROR)|passed|failed|rc " gpurun_out/r2_pytest29.log | cut -c1-300 | head -20
bash profiles/ncu_r2.sh
