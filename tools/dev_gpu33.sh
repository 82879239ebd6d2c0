#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q -k "f16x3" -s 2>&1 | tail -30 | tee gpurun_out/pytest_33.log
