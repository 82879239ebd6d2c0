#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_roundtrip_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_55.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_55.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})"
