#!/bin/bash
# experiment: tokens per ragged model call (2^19 default vs 2^20, 3*2^18)
mkdir -p gpurun_out
for mt in 524288 1048576 786432; do
  SCP_MAX_TOKENS=$mt timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_mt_$mt.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_mt_$mt.log").read())
print($mt, d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done
