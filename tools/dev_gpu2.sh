#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_ops.log
timeout 1200 python -m pytest tests/test_models_gpu.py -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_models.log
