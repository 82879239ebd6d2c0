#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k knn 2>&1 | tail -2 | tee gpurun_out/pytest_18.log
timeout 120 python tools/exp_knn.py 192 2>&1 | tail -1 | tee -a gpurun_out/exp_knn3.log
SCP_KNN_TRIG=0 timeout 600 python tools/exp_knn_trig.py 2>&1 | head -2 | tee -a gpurun_out/exp_knn3.log
