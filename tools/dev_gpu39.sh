#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke39.log
