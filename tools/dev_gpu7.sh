#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "knn" 2>&1 | tail -15 | tee gpurun_out/pytest_knn.log
timeout 600 python -m pytest tests/test_octree_gpu.py tests/test_models_gpu.py tests/test_e2e_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_b.log
timeout 600 python tools/trace_step.py 2>&1 | tail -12 | tee gpurun_out/trace.log
timeout 600 python tools/bench_octree.py 2>&1 | tail -5 | tee gpurun_out/bench_octree2.log
