#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_octree_gpu.py tests/test_dropin_gpu.py tests/test_roundtrip_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_40.log
timeout 300 python tools/bench_octree.py 2>&1 | tee gpurun_out/bench_octree_40.log
