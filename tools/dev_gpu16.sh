#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2; do SCP_KNN_DBG=$dbg timeout 120 python tools/exp_knn.py 192 2>&1 | tail -1 | tee -a gpurun_out/exp_knn.log; done
SCP_KNN_DBG=0 timeout 120 python tools/exp_knn.py 144 2>&1 | tail -1 | tee -a gpurun_out/exp_knn.log
