#!/bin/bash
# OctAttention kernel with four threads per row: op test, e2e goldens, config-4 bench; then the sanitizer pass over both models
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "octattn" 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "octattn op test failed: stop"; exit 1; }
timeout 900 python -m pytest tests/test_octattn_e2e.py tests/test_models_gpu.py tests/test_e2e_gpu.py -q -m gpu -x 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_44_bench4.log 2> gpurun_out/r2_44_bench4.err
tail -2 gpurun_out/r2_44_bench4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_44_bench4.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
SAN_TOOLS="memcheck racecheck" bash tools/gpu/sanitize.sh
