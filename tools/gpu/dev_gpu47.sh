#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_47.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_swin_attn_h|k_attn_prep" -s 0 -c 4 -o gpurun_out/prof_attn_h2 -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_attn_h2.log 2>&1
tail -1 gpurun_out/ncu_attn_h2.log
