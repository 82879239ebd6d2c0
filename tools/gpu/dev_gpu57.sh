#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_57.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_57.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})"
