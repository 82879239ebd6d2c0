#!/bin/bash
# round 2, call 2: new unit tests, then both bench arms (reference first, like the driver)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py tests/test_roundtrip_gpu.py tests/test_library_baseline_gpu.py tests/test_octattn_e2e.py -q -m gpu -x -s -k "cache or batch or single_node or library or two_ranks" 2>&1 | grep -v "^$" | tail -25 | cut -c1-1500 > gpurun_out/r2_02_tests.log
cat gpurun_out/r2_02_tests.log
nproc
( time timeout 1200 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r2_02_bench_ref.log 2> gpurun_out/r2_02_bench_ref.err
tail -3 gpurun_out/r2_02_bench_ref.err; tail -1 gpurun_out/r2_02_bench_ref.log | cut -c1-2500
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_02_bench.log 2> gpurun_out/r2_02_bench.err
tail -5 gpurun_out/r2_02_bench.err; tail -1 gpurun_out/r2_02_bench.log | cut -c1-6000
