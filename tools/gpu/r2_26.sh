#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x -k cluster 2>&1 | grep -v "^$" | tail -40
