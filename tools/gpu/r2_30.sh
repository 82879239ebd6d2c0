#!/bin/bash
mkdir -p gpurun_out
for r in 1 4 16; do echo "== SCP_GEMM_WREP=$r (CL=1)"; SCP_GEMM_CL=1 SCP_GEMM_WREP=$r timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5; done | tee gpurun_out/r2_30_gemm_time.log
