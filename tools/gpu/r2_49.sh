#!/bin/bash
# final state of round 2: launch list, full gpu suite, smoke and both bench arms (reference first)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_49_launches.csv \
    python tools/prof_step.py 1 > gpurun_out/r2_49_ll.log 2>&1
tail -1 gpurun_out/r2_49_ll.log
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "gpu suite failed: stop"; exit 1; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r2_49_bench_ref.log 2> gpurun_out/r2_49_bench_ref.err
tail -3 gpurun_out/r2_49_bench_ref.err | head -1
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_49_bench.log 2> gpurun_out/r2_49_bench.err
tail -4 gpurun_out/r2_49_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_49_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"], d["decode"]["round_trip_exact"], d["clocks"])
print({k:(round(v["ms_per_step"],2), v["launches_per_step"], round(v["frac_of_peak"] or 0,3)) for k,v in d["kernels"].items()})
for k,v in d["octree_stages"]["stages"].items(): print(k, v)
for k,v in d["other_configs"].items(): print(k, v.get("value"), v.get("e2e",{}).get("value"), v.get("ms_per_step"))
print(d["cpu_baseline"])
PY
