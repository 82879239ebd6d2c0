#!/bin/bash
# round-1 final evidence (after the fp16 attention / kNN pruning / GELU changes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_50.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke50.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01e.csv \
    python tools/prof_step.py 1 > gpurun_out/ncu_ll50.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_knn_small|k_knn_rerank" -s 3 -c 3 -o gpurun_out/prof_knn_final -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_knn_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_x3_ts" -s 150 -c 4 -o gpurun_out/prof_gemm_final -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_gemm_final.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_50.log | cut -c1-400
