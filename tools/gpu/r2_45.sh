#!/bin/bash
# k_knn_tc4 (eight epilogue warps, four threads per query row): op tests, same-box A/B against k_knn_tc, parity + round trips, bench kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "knn" 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "knn tests failed: stop"; exit 1; }
for eng in 0 1; do for d in 192 144; do
SCP_KNN_ENGINE=$eng SCP_KNN_TRACE=1 timeout 300 python tools/exp_knn.py $d 2>&1 | grep -E "dbg=|tiles 2-" | awk '/dbg=/{print; next} {k=$0} 1' | grep -B1 "dbg=" | grep -v "^--" | sed "s/^/eng=$eng /"
done; done > gpurun_out/r2_45_knn_trace.log 2>&1
cut -c1-290 gpurun_out/r2_45_knn_trace.log
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_roundtrip_gpu.py tests/test_e2e_gpu.py -q -m gpu -x 2>&1 | tail -3
for eng in 1 0; do
SCP_KNN_ENGINE=$eng timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_45_bench_$eng.log 2> gpurun_out/r2_45_bench_$eng.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_45_bench_$eng.log").read().strip().splitlines()[-1])
print("engine $eng:", d["value"], d["e2e"]["value"], d["ms_per_step"], d["decode"]["round_trip_exact"], d["clocks"]["sm_mhz"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"]) for k,v in d["kernels"].items()})
PY
done
