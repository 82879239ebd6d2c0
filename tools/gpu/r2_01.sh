#!/bin/bash
# round 2, call 1: new parity tests (OctAttention e2e, explained-rows criterion, pc_error PSNR pin), then the whole gpu suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests/test_octattn_e2e.py tests/test_models_gpu.py tests/test_metrics.py -q -m gpu -s 2>&1 | grep -v "^$" > gpurun_out/r2_01_new.log
tail -40 gpurun_out/r2_01_new.log | cut -c1-600
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | cut -c1-600
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_01_all.log
tail -15 gpurun_out/r2_01_all.log
