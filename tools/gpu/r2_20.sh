#!/bin/bash
# k_knn_tc2 pairwise-interleaved candidate tiles, two epilogue teams: tests + timeline + bench kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "knn" 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "knn tests failed: stop"; exit 1; }
for eng in 0; do for d in 192 144; do
SCP_KNN_ENGINE=$eng SCP_KNN_TRACE=1 timeout 300 python tools/exp_knn.py $d 2>&1 | grep -E "dbg=|tiles 2-|tiles 68-" | awk '/dbg=/{print; next} {k=$0} 1' | grep -B2 "dbg=" | sed "s/^/eng=$eng /"
done; done > gpurun_out/r2_20_knn_trace.log 2>&1
grep -v -- "--" gpurun_out/r2_20_knn_trace.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_20_bench.log 2> gpurun_out/r2_20_bench.err
tail -3 gpurun_out/r2_20_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_20_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"].get("bpp_dev"), d["decode"]["round_trip_exact"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
PY
