#!/bin/bash
# deep-ring A-stationary nn.Linear: GEMM tests, per-layer times at CL = 1 / 2 / 4, MMA timeline, bench kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -q -m gpu -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: stop"; exit 1; }
for cl in 1 2 4; do echo "== SCP_GEMM_CL=$cl"; SCP_GEMM_CL=$cl timeout 200 python tools/exp_gemm_time.py 2>&1 | tail -5; done | tee gpurun_out/r2_28_gemm_time.log
SCP_GEMM_CL=1 SCP_GEMM_TRACE=1 timeout 300 python tools/exp_gemm_trace.py 2>&1 | grep -A9 "N=768 K=256" | head -10
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_e2e_gpu.py tests/test_roundtrip_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_28_bench.log 2> gpurun_out/r2_28_bench.err
tail -3 gpurun_out/r2_28_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_28_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"].get("bpp_dev"), d["decode"]["round_trip_exact"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
