#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -x -q -s 2>&1 | tail -32 | tee gpurun_out/pytest_tc.log
SCP_GEMM=tf32x3 timeout 600 python -m pytest tests/test_models_gpu.py tests/test_e2e_gpu.py -q -s 2>&1 | tail -30 | tee gpurun_out/pytest_models_x3.log
timeout 600 python -m pytest tests/test_dropin_gpu.py -q 2>&1 | tail -30 | tee gpurun_out/pytest_dropin.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
SCP_GEMM=tf32x3 timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_x3.log
