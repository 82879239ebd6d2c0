#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_octree_gpu.py tests/test_e2e_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_20.log
timeout 300 python tools/bench_octree.py 2>&1 | tail -8 | tee gpurun_out/bench_octree3.log
