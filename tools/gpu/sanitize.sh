#!/bin/bash
# compute-sanitizer over the whole hot path on a small frame (SURVEY section 5 hygiene): memcheck (out-of-bounds / misaligned
# accesses, incl. the TMA / tcgen05 kernels) and racecheck (shared-memory hazards in the mbarrier pipelines and the
# decoupled-look-back sort).  Logs -> gpurun_out/sanitize_*.log; the summary lines are copied to profiles/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  for which in ehem octattn; do
    timeout ${SAN_TIMEOUT:-150} $CS --tool $tool --print-limit 10 --error-exitcode 0 python tools/sanitize_case.py $which ${SAN_POINTS:-1000} \
      > gpurun_out/sanitize_${tool}_${which}.log 2>&1
    echo "== $tool $which rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|nodes' gpurun_out/sanitize_${tool}_${which}.log | tr '\n' ' ')"
  done
done
