#!/bin/bash
# distortion report (row f-4): parity tests + timing of one full-size frame
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_metrics.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/metrics_tests.log
cat gpurun_out/metrics_tests.log
timeout 300 python tools/bench_metrics.py > gpurun_out/metrics_bench.log 2>&1
cat gpurun_out/metrics_bench.log
