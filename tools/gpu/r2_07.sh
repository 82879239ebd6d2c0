#!/bin/bash
# OctAttention tensor-core attention: op test (short timeout: a protocol bug must not burn the budget), model parity, e2e
# goldens, then config-4 timing
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k octattn 2>&1 | tail -15
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "op test failed or hung: stop"; exit 1; }
timeout 300 python -m pytest tests/test_models_gpu.py tests/test_octattn_e2e.py -q -m gpu -x -k "octattn or compress or encoder or pmf" 2>&1 | tail -8
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "model tests failed or hung: stop"; exit 1; }
SCP_OCTATTN_ENGINE=1 timeout 300 python bench.py --config 4 --steps 3 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_07_bench_1.log 2> gpurun_out/r2_07_bench_1.err
tail -2 gpurun_out/r2_07_bench_1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_07_bench_1.log").read().strip().splitlines()[-1])
print("engine 1", d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["bpp_mean"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
print(d["roofline"]["kernel"], d["roofline"]["frac"])
PY
