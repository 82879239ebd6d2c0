#!/bin/bash
# multi-GPU: default headline (config 2) and the config-5 1000-frame sweep under torchrun, N = $1 ranks
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-other-configs --no-cpu-parity > gpurun_out/r2_34_bench_n$N.log 2> gpurun_out/r2_34_bench_n$N.err
tail -2 gpurun_out/r2_34_bench_n$N.err
timeout 900 $TR bench.py --gpus $N --config 5 --no-other-configs --no-cpu-parity > gpurun_out/r2_34_sweep_n$N.log 2> gpurun_out/r2_34_sweep_n$N.err
tail -2 gpurun_out/r2_34_sweep_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r2_34_bench_n$N.log", "gpurun_out/r2_34_sweep_n$N.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["config"]["baseline_config"], round(d["value"],2), round(d["e2e"]["value"],2), d["ms_per_step"], d["scaling"], d.get("config",{}).get("frames_total"), d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
