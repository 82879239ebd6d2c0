#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_head_hist|k_emit_nodes|k_occupancy|k_context" -s 4 -c 4 -o gpurun_out/prof_octree3 -f \
    python tools/prof_octree.py > gpurun_out/ncu_octree3.log 2>&1
tail -3 gpurun_out/ncu_octree3.log
