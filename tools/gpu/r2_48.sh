#!/bin/bash
# kNN scan without the per-score -|q|^2 (two instructions per candidate): op tests, parity, round trips, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_roundtrip_gpu.py tests/test_e2e_gpu.py tests/test_octattn_e2e.py -q -m gpu -x 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
for d in 192 144; do timeout 300 python tools/exp_knn.py $d 2>&1 | grep "dbg=0"; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_48_bench.log 2> gpurun_out/r2_48_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_48_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["decode"]["round_trip_exact"], d["clocks"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
