#!/bin/bash
# end-of-round validation: full GPU suite, smoke, bench (un-profiled), then one ncu capture of the distortion kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_62.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke62.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_62.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_nn_dist2|k_dequantise_keys" -s 4 -c 3 -o gpurun_out/prof_metrics -f \
    python tools/bench_metrics.py > gpurun_out/ncu_metrics.log 2>&1
tail -2 gpurun_out/ncu_metrics.log
