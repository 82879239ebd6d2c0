#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc" -s 1 -c 1 -o gpurun_out/prof_knn_h3 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_knn_h3.log 2>&1
tail -2 gpurun_out/ncu_knn_h3.log
