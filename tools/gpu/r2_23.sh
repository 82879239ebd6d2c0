#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3; do SCP_GEMM_DBG=$dbg timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5; done | tee gpurun_out/r2_23_gemm_time.log
SCP_GEMM_DBG=1 SCP_GEMM_TRACE=1 timeout 300 python tools/exp_gemm_trace.py 2>&1 | grep -A9 "N=768 K=256" | head -10
