#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_coder_gpu.py -q -m gpu -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "coder tests failed: stop"; exit 1; }
timeout 120 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from scp_b200 import coder
n = 514344
x = torch.randn(n, 255, device="cuda") * 3
sym = torch.randint(0, 255, (n,), device="cuda").to(torch.int16)
iv = torch.empty((n, 2), dtype=torch.int32, device="cuda")
cdf = torch.empty((n, 256), dtype=torch.uint16, device="cuda")
for what, out in (("interval", {"interval": iv}), ("cdf+interval", {"interval": iv, "cdf": cdf})):
    for _ in range(3): coder.pmf_to_cdf(x, sym=sym, is_logits=True, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): coder.pmf_to_cdf(x, sym=sym, is_logits=True, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = n * (1020 + (8 if what == "interval" else 520))
    print(what, round(ms, 4), "ms", round(by / ms / 1e6, 1), "GB/s", round(by / ms / 1e6 / 6547.2, 3), "of HBM peak")
PY
