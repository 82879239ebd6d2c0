#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_roundtrip_gpu.py -x -q -s 2>&1 | tail -30 | tee gpurun_out/pytest_30.log
