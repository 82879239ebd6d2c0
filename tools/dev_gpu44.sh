#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "swin_attention" 2>&1 | tail -15 | tee gpurun_out/pytest_44.log
