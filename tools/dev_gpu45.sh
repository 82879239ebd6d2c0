#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "swin_attention" 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_swin_attn_h" -s 12 -c 2 -o gpurun_out/prof_attn_h -f \
    python tools/prof_step.py 1 > gpurun_out/ncu_attn_h.log 2>&1
tail -2 gpurun_out/ncu_attn_h.log
