#!/bin/bash
for cfg in "1 70001 768 256" "1 70016 768 256" "1 70016 256 256" "1 70016 512 256" "2 70016 768 256" "2 70001 768 256" "4 70016 768 256" "1 70016 768 192"; do
  timeout 60 python tools/exp_gemm_cl.py $cfg 2>&1 | tail -1 | cut -c1-200
done
