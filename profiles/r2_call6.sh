#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/diag_lane.py > gpurun_out/r2_diag_lane.log 2>&1; cat gpurun_out/r2_diag_lane.log | cut -c1-1200
