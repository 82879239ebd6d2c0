#!/bin/bash
# packed f32x2 GELU epilogue: GEMM / op / model parity tests, bench kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_e2e_gpu.py tests/test_roundtrip_gpu.py -q -m gpu -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_22_bench.log 2> gpurun_out/r2_22_bench.err
tail -3 gpurun_out/r2_22_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_22_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"].get("bpp_dev"), d["decode"]["round_trip_exact"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
SCP_GEMM_TRACE=1 timeout 300 python tools/exp_gemm_trace.py 2>&1 | grep -A14 "N=1024 K=256" | head -34
