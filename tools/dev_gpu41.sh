#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_octree_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_41.log
