#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_quantise_frames|k_frame_stats|k_onesweep" -s 2 -c 4 -o gpurun_out/prof_quant -f \
    python tools/prof_octree.py > gpurun_out/ncu_quant.log 2>&1
tail -2 gpurun_out/ncu_quant.log
