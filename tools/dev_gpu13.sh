#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_tc_gpu.py tests/test_models_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_13.log
timeout 600 python tools/exp_step.py 2>&1 | tail -6 | tee gpurun_out/exp_step5.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_13.log
