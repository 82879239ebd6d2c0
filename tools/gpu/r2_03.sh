#!/bin/bash
# round 2, call 3: fused quantise + key-pass tree builder: bit-identity tests, then the octree stage timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_octree_gpu.py tests/test_e2e_gpu.py tests/test_roundtrip_gpu.py tests/test_dropin_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_03_tests.log
cat gpurun_out/r2_03_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_03_bench.log 2> gpurun_out/r2_03_bench.err
tail -3 gpurun_out/r2_03_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_03_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"].get("bpp_dev"), d["decode"]["round_trip_exact"])
for k,v in d["octree_stages"]["stages"].items(): print(k, v)
PY
