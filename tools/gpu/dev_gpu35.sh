#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_x3_ts" -s 150 -c 4 -o gpurun_out/prof_gemm_h3 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_gemm_h3.log 2>&1
tail -2 gpurun_out/ncu_gemm_h3.log
