#!/bin/bash
# round-2 final ncu evidence: launch list of one K16-mullevel frame, --set full of nn.Linear (cluster kernel), kNN, attention
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_32_launches.csv \
    python tools/prof_step.py 1 > gpurun_out/r2_32_ll.log 2>&1
tail -1 gpurun_out/r2_32_ll.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_x3_ts" -s 150 -c 4 -o gpurun_out/r2_32_gemm -f \
    python tools/prof_step.py 1 > gpurun_out/r2_32_gemm.log 2>&1
tail -1 gpurun_out/r2_32_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_knn_small|k_knn_rerank" -s 3 -c 3 -o gpurun_out/r2_32_knn -f \
    python tools/prof_step.py 1 > gpurun_out/r2_32_knn.log 2>&1
tail -1 gpurun_out/r2_32_knn.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_attn_prep|k_swin_attn_h" -s 24 -c 2 -o gpurun_out/r2_32_attn -f \
    python tools/prof_step.py 1 > gpurun_out/r2_32_attn.log 2>&1
tail -1 gpurun_out/r2_32_attn.log
ls -la gpurun_out/r2_32_*
