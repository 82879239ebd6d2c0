#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_quantise_frames|k_filter_count|k_compact_hist|k_onesweep|k_head_hist|k_emit_nodes|k_occupancy|k_context" -s 15 -c 15 -o gpurun_out/prof_octree2 -f \
    python tools/prof_octree.py > gpurun_out/ncu_octree2.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_22.log
