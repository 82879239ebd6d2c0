#!/bin/bash
# edge_gather_max with a second output view (three DGCNN copies removed): full gpu suite + bench kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "gpu suite failed: stop"; exit 1; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_47_bench.log 2> gpurun_out/r2_47_bench.err
tail -3 gpurun_out/r2_47_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_47_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"].get("bpp_dev"), d["decode"]["round_trip_exact"], d["clocks"], d["gpu_launches"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
