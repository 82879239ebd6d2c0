#!/bin/bash
mkdir -p gpurun_out
for f in 1 4 8; do
timeout 600 python bench.py --steps 2 --warmup 2 --frames-per-step $f 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('frames/step', d['config']['frames_per_step_per_gpu'], 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],1))" | tee -a gpurun_out/bench_fps.log
done
