#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_tc_gpu.py tests/test_models_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_14.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_14.log
