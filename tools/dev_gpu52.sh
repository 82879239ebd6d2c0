#!/bin/bash
mkdir -p gpurun_out
for as in 1 0; do
SCP_GEMM_AS=$as timeout 900 python bench.py --steps 3 --warmup 3 --frames-per-step 2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('AS=$as', d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})" | tee -a gpurun_out/bench_52.log
done
