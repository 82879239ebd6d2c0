#!/bin/bash
# ncu: per-kernel durations of the octree pipeline (256 K16-mullevel frames) + full capture of the new kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_04_launches.csv python tools/prof_octree.py > gpurun_out/r2_04_ll.log 2>&1
tail -2 gpurun_out/r2_04_ll.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tree_rows|k_tree_occ|k_quantise_fused" -c 6 -f -o gpurun_out/r2_04_octree python tools/prof_octree.py > gpurun_out/r2_04_ncu.log 2>&1
tail -2 gpurun_out/r2_04_ncu.log
ls -la gpurun_out/r2_04_octree.ncu-rep
