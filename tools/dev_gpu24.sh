#!/bin/bash
mkdir -p gpurun_out
SCP_OCT_DEBUG=1 timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/dbg_24.log
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_octree
bench_octree.main(64, 16, False)
PY
