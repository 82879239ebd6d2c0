#!/bin/bash
# compute-sanitizer over the round-2 nn.Linear paths: deep-ring A-stationary kernel (CL = 1) and the cluster / multicast kernel (CL = 2, 4)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  for cfg in "1 40001 768 256" "2 40001 768 256" "4 40001 512 256"; do
    timeout 280 $CS --tool $tool --print-limit 5 --error-exitcode 0 python tools/exp_gemm_cl.py $cfg > gpurun_out/r2_37_san.log 2>&1
    echo "== $tool CL,M,N,K = $cfg : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok, max err' gpurun_out/r2_37_san.log | tr '\n' ' ')"
  done
done 2>&1 | tee gpurun_out/r2_37_sanitize.log
