#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest40.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest40.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  |rc " gpurun_out/r2_pytest40.log | cut -c1-300 | head -20
