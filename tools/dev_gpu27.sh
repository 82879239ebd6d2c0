#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_octree_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_27.log
timeout 300 python tools/exp_ctx.py 2>&1 | tee gpurun_out/exp_ctx.log
