#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_tc_gpu.py tests/test_models_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_17.log
timeout 120 python tools/exp_knn.py 192 2>&1 | tail -1 | tee -a gpurun_out/exp_knn2.log
timeout 120 python tools/exp_knn.py 144 2>&1 | tail -1 | tee -a gpurun_out/exp_knn2.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_17.log
