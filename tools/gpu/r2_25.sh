#!/bin/bash
# cluster-multicast nn.Linear: bit-identity test, per-layer times at CL = 1 / 2 / 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
for cl in 1 2 4; do echo "== SCP_GEMM_CL=$cl"; SCP_GEMM_CL=$cl timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5; done | tee gpurun_out/r2_25_gemm_time.log
