#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_gpu.py -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_31.log
