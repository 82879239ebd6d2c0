#!/bin/bash
# experiment: conv2 / mlp2 of the DGCNN stage on the 3xFP16 tensor path instead of the fp32 SIMT GEMM
mkdir -p gpurun_out
SCP_GEO_ENGINE=auto timeout 600 python -m pytest tests/test_models_gpu.py tests/test_e2e_gpu.py tests/test_roundtrip_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/geo_tests.log
tail -12 gpurun_out/geo_tests.log
for g in auto simt; do
  SCP_GEO_ENGINE=$g timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_geo_$g.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_geo_$g.log").read())
print("$g", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["e2e"]["bpp_mean"], d["decode"]["round_trip_exact"])
PY
done
