#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/exp_step.py 2>&1 | tail -8 | tee gpurun_out/exp_step.log
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "layernorm" 2>&1 | tail -3 | tee gpurun_out/pytest_ln.log
