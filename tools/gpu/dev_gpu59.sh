#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_59.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke59.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_59.log | cut -c1-300
