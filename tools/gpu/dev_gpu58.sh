#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_roundtrip_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_58.log
timeout 900 python bench.py --steps 2 --warmup 2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['decode'])" | tee gpurun_out/bench_58.log
