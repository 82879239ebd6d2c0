#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_all.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1500 python bench.py --steps 2 --warmup 1 2>&1 | tail -5 | tee gpurun_out/bench1.log
