#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_ctx.py 2>&1 | tee gpurun_out/exp_ctx.log
