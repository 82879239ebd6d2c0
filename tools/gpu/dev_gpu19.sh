#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_all19.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke19.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01c.csv \
    python tools/prof_step.py 1 > gpurun_out/ncu_ll19.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_x3_ts" -s 150 -c 4 -o gpurun_out/prof_gemm_ts -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_gemm_ts.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_swin_attn_tc" -s 24 -c 2 -o gpurun_out/prof_attn_tc3 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_attn_tc3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_knn_small|k_knn_rerank" -s 3 -c 3 -o gpurun_out/prof_knn_tc4 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_knn_tc4.log 2>&1
ls -la gpurun_out | tail -6
