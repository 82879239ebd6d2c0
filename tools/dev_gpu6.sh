#!/bin/bash
# bench + ncu evidence (launch list and full capture of the top kernels)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_r01.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --frames-per-step 1 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn$ -s 1 -c 1 -o gpurun_out/prof_knn \
    python bench.py --steps 1 --warmup 1 --frames-per-step 1 > gpurun_out/ncu_knn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 -s 40 -c 2 -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 1 --frames-per-step 1 > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep|k_context|k_emit_nodes|k_quantise_keys" -s 40 -c 12 -o gpurun_out/prof_octree \
    python tools/bench_octree.py > gpurun_out/ncu_octree.log 2>&1
ls -la gpurun_out
