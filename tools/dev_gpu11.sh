#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/exp/ts_mma 2>&1 | tail -10 | tee gpurun_out/ts_mma.log
timeout 600 python tools/exp_step.py 2>&1 | tail -8 | tee gpurun_out/exp_step2.log
