#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_octattn_e2e.py -q -m gpu -x -k "layernorm or octattn or compress or encoder" 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
timeout 300 python bench.py --config 4 --steps 3 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_13_bench.log 2> gpurun_out/r2_13_bench.err
tail -2 gpurun_out/r2_13_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_13_bench.log").read().strip().splitlines()[-1])
print("config4", d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["bpp_mean"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
