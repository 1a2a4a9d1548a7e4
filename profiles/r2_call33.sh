#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest33.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest33.log
grep -E "^(FAILED|ERROR)|passed|failed|rc " gpurun_out/r2_pytest33.log | cut -c1-300 | head -20
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
