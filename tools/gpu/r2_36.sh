#!/bin/bash
echo "== CL=2 multicast"; SCP_GEMM_CL=2 timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5
echo "== CL=2, each CTA receives only HALF of every weight tile (timing only)"; SCP_GEMM_CL=2 SCP_GEMM_HALFW=1 timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5
echo "== CL=4, each CTA receives a QUARTER"; SCP_GEMM_CL=4 SCP_GEMM_HALFW=1 timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5
