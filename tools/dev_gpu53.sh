#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_quantise_frames" -s 1 -c 1 -o gpurun_out/prof_quant2 -f \
    python tools/prof_octree.py > gpurun_out/ncu_quant2.log 2>&1
tail -1 gpurun_out/ncu_quant2.log
