#!/bin/bash
mkdir -p gpurun_out
export SCP_AUTO_ENGINE=2
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q -k "f16x3" 2>&1 | tail -3 | tee gpurun_out/pytest_34a.log
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_e2e_gpu.py tests/test_roundtrip_gpu.py -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_34.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_34.log
