#!/bin/bash
# window attention with four threads per row (16x256b TMEM loads, quad-shuffle max / sum): op tests, parity, round trips, bench kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "attention op tests failed: stop"; exit 1; }
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_e2e_gpu.py tests/test_roundtrip_gpu.py tests/test_dropin_gpu.py -q -m gpu -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_42_bench.log 2> gpurun_out/r2_42_bench.err
tail -3 gpurun_out/r2_42_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_42_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"].get("bpp_dev"), d["decode"]["round_trip_exact"], d["clocks"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
