#!/bin/bash
mkdir -p gpurun_out
echo "== fp16 maps (CL=1)"; SCP_GEMM_CL=1 timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5
echo "== W as 32-bit words (CL=1)"; SCP_GEMM_CL=1 SCP_GEMM_W32=1 timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5
SCP_GEMM_CL=1 SCP_GEMM_W32=1 timeout 100 python tools/exp_gemm_cl.py 1 70016 768 256 | tail -1
SCP_GEMM_CL=1 SCP_GEMM_W32=1 timeout 100 python tools/exp_gemm_cl.py 1 7000 256 1024 | tail -1
