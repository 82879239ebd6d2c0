#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_21.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_21.log
