#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc" -s 2 -c 2 -o gpurun_out/prof_knn_tc2 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_knn_tc2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_swin_attn_tc" -s 24 -c 2 -o gpurun_out/prof_attn_tc2 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_attn_tc2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tf32" -s 150 -c 4 -o gpurun_out/prof_gemm3 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_gemm3.log 2>&1
timeout 600 python tools/exp_step.py 2>&1 | tail -6 | tee gpurun_out/exp_step6.log
ls -la gpurun_out | tail -5
