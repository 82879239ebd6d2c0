#!/bin/bash
# kNN kernel timeline (clock64 stamps of CTA 0) at d = 192 and 144, full / no-accept / no-scan
mkdir -p gpurun_out
for d in 192 144; do
SCP_KNN_TRACE=1 timeout 300 python tools/exp_knn.py $d 2>&1 | grep -E "dbg=|tiles 2-|tiles 68-" | awk '!seen[$0]++' | head -40
done > gpurun_out/r2_15_knn_trace.log 2>&1
cat gpurun_out/r2_15_knn_trace.log | cut -c1-330
