#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config 5 --no-other-configs --no-cpu-parity > gpurun_out/r2_34_sweep_n1.log 2> gpurun_out/r2_34_sweep_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_34_sweep_n1.log").read().strip().splitlines()[-1])
print(d["n_gpus"], d["config"]["baseline_config"], round(d["value"],2), round(d["e2e"]["value"],2), d["ms_per_step"], d["scaling"], d["clocks"])
PY
