"""tcgen05 (TF32, TMEM accumulators, TMA-fed) GEMM engine against an fp64 statement of nn.Linear."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def V(t, col=0, ncol=None):
    return (t, col, t.shape[1] - col if ncol is None else ncol)


@pytest.mark.parametrize("M,N,K,act,res,step", [
    (128, 256, 256, "none", False, 1), (1000, 256, 256, "none", False, 1), (5000, 768, 256, "none", False, 1),
    (4096, 1024, 1280, "leaky", False, 1), (2500, 255, 512, "none", False, 1), (640, 240, 256, "relu", False, 2),
    (300, 64, 80, "leaky", False, 1), (3000, 256, 1024, "none", True, 1), (2500, 1024, 256, "gelu", False, 1),
    (70000, 256, 256, "none", True, 1), (33, 128, 448, "none", False, 1)])
@pytest.mark.parametrize("engine,tol", [("tf32", 4e-3), ("tf32x3", 1e-5)])
def test_linear_tensor_core_engines(M, N, K, act, res, step, engine, tol):
    from scp_b200.ops import CudaOps
    cu = CudaOps(engine=engine)
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M * step, K + 8, generator=g)
    w = torch.randn(N, K, generator=g) * 0.1
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    xs = x[step - 1::step, 4:4 + K].double()
    ref = xs @ w.double().T + b.double()
    ref = {"none": lambda t: t, "leaky": lambda t: torch.nn.functional.leaky_relu(t, 0.01),
           "gelu": torch.nn.functional.gelu, "relu": torch.relu}[act](ref)
    if res:
        ref = ref + r.double()
    y = torch.full((M, N + 4), float("nan"), device="cuda")
    cu.linear(V(x.cuda(), 4, K), w.cuda(), b.cuda(), V(y, 0, N), act=act, res=V(r.cuda()) if res else None,
              row_step=step, row_off=step - 1, rows=M)
    torch.cuda.synchronize()
    got = y[:, :N].cpu().double()
    assert torch.isfinite(got).all()
    assert torch.isnan(y[:, N:]).all()                       # nothing written outside the view
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(M, N, K, engine, "max err", err, "scale", scale)
    assert err < tol * scale * max(1.0, (K / 256) ** 0.5)
