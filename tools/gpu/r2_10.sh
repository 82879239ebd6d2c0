#!/bin/bash
# ncu --set full: octree kernels of the final design (256 K16-mullevel frames) and the OctAttention tensor-core attention
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep|k_quantise_fused|k_tree_occ|k_context_lean|k_frame_stats|k_sort_hist|k_head_hist" --launch-skip 13 -c 13 -f -o gpurun_out/r2_10_octree python tools/prof_octree.py > gpurun_out/r2_10_ncu.log 2>&1
tail -2 gpurun_out/r2_10_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_octattn_attn_h|k_octattn_prep" --launch-skip 2 -c 2 -f -o gpurun_out/r2_10_octattn python tools/prof_octattn.py > gpurun_out/r2_10_ncu2.log 2>&1
tail -2 gpurun_out/r2_10_ncu2.log
ls -la gpurun_out/r2_10_*.ncu-rep
