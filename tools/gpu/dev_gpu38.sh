#!/bin/bash
# round-1 final evidence: launch list, ncu --set full captures of the octree kernels, the 3xFP16 GEMM and the kNN kernel, bench line
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01d.csv \
    python tools/prof_step.py 1 > gpurun_out/ncu_ll38.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_quantise_frames|k_filter_count|k_compact_hist|k_onesweep|k_head_hist|k_emit_nodes|k_occupancy|k_context" -s 15 -c 15 -o gpurun_out/prof_octree4 -f \
    python tools/prof_octree.py > gpurun_out/ncu_octree4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_knn_small|k_knn_rerank" -s 3 -c 3 -o gpurun_out/prof_knn_h3b -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_knn_h3b.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_38.log
ls -la gpurun_out | tail -8
