#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
for w in 0 1; do echo "== SCP_GEMM_WIDE=$w"; SCP_GEMM_WIDE=$w timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5; done | tee gpurun_out/r2_29_gemm_time.log
