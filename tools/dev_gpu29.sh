#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_29.log
timeout 300 python tools/bench_octree.py 2>&1 | tee gpurun_out/bench_octree_29.log
