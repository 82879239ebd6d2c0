#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_tc_gpu.py tests/test_models_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_12.log
timeout 600 python tools/exp_step.py 2>&1 | tail -6 | tee gpurun_out/exp_step4.log
# launch list of one step (second step: skip the warm-up step's launches by taking everything and filtering later)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01b.csv \
    python tools/prof_step.py 1 > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc" -s 2 -c 2 -o gpurun_out/prof_knn_tc -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_knn_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_swin_attn_tc" -s 24 -c 2 -o gpurun_out/prof_attn_tc -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_attn_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tf32" -s 150 -c 4 -o gpurun_out/prof_gemm2 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_gemm2.log 2>&1
ls -la gpurun_out | tail -8
