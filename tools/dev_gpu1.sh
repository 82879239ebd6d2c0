#!/bin/bash
# first GPU contact: octree + coder parity, stage timings
export SCP_DEV_ALLOW_MISSING=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_octree_gpu.py tests/test_coder_gpu.py -x -q 2>&1 | tail -40 | tee gpurun_out/pytest1.log
timeout 600 python tools/bench_octree.py 2>&1 | tail -20 | tee gpurun_out/bench_octree.log
