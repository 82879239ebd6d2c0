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
@pytest.mark.parametrize("engine,tol", [("tf32", 4e-3), ("tf32x3", 1e-5), ("f16x3", 1e-5)])
def test_linear_tensor_core_engines(M, N, K, act, res, step, engine, tol):
    from scp_b200.ops import CudaOps
    cu = CudaOps(engine=engine)
    cu.lib.scp_gemm_cache_clear()          # split weights are cached per device pointer (the models clear it in _prepare)
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


@pytest.mark.parametrize("xs,ws,tol", [(1.0, 1.0, 2e-6), (300.0, 1e-3, 2e-6), (1e-3, 30.0, 2e-5), (20.0, 1e-5, 2e-6)])
def test_f16x3_dynamic_range(xs, ws, tol):
    """3xFP16 engine: fp16 has 5 exponent bits, so the weights are rescaled per matrix and the activations are taken as they
    are.  Rows mixing large and tiny activations, tiny / large weights: the error stays at the 3xTF32 level relative to the
    size of the result (what falls under the fp16 subnormal step is an absolute error of 2^-25 per term)."""
    from scp_b200.ops import CudaOps
    cu = CudaOps(engine="f16x3")
    cu.lib.scp_gemm_cache_clear()          # split weights are cached per device pointer (the models clear it in _prepare)
    g = torch.Generator().manual_seed(7)
    M, N, K = 3000, 256, 1024
    x = torch.randn(M, K, generator=g) * xs
    x[:, ::7] *= 1e-4                                          # tiny entries next to large ones
    x[::5] *= 50.0
    w = torch.randn(N, K, generator=g) * ws
    w[:, ::11] *= 1e-3
    b = torch.randn(N, generator=g) * xs * ws
    ref = x.double() @ w.double().T + b.double()
    y = torch.empty((M, N), device="cuda")
    cu.linear(V(x.cuda()), w.cuda(), b.cuda(), V(y))
    got = y.cpu().double()
    assert torch.isfinite(got).all()
    row_scale = (x.double().abs() @ w.double().abs().T)        # size of the terms that were summed
    rel = ((got - ref).abs() / row_scale).max().item()
    print("xs", xs, "ws", ws, "max err / sum|terms|", rel)
    assert rel < tol          # activations of 1e-3 and below sit on the fp16 subnormal step: absolute 3e-8 per term


def test_weight_split_cache_follows_the_weights():
    """The hi/lo weight splits are cached per weight matrix (csrc/gemm_tc.cu).  The cache entry must die with the weights:
    (a) a new tensor that the caching allocator places at the address of a freed one, (b) an in-place update
    (``load_state_dict`` copies into the same storage), for both entropy models."""
    from scp_b200.ops import CudaOps, V
    from test_models_cpu import cfg_ehem, cfg_oct
    ops = CudaOps()
    torch.manual_seed(0)
    x = torch.randn(256, 256, device="cuda")
    seen = set()
    for i in range(4):                                            # (a) same shape, freed and re-allocated
        w = torch.randn(128, 256, device="cuda") * (i + 1)
        seen.add(w.data_ptr())
        y = torch.empty(256, 128, device="cuda")
        ops.linear(V(x), w, None, V(y))
        ref = (x.double() @ w.double().T).float()
        assert (y - ref).abs().max() <= 2e-5 * ref.abs().max(), i
        del w
    w = torch.randn(128, 256, device="cuda")
    y = torch.empty(256, 128, device="cuda")
    ops.linear(V(x), w, None, V(y))
    w.mul_(-3.0)                                                  # (b) in place
    ops.linear(V(x), w, None, V(y))
    ref = (x.double() @ w.double().T).float()
    assert (y - ref).abs().max() <= 2e-5 * ref.abs().max()
    from scp_b200 import weights as W
    from scp_b200.models import EHEM, OctAttention
    g = torch.Generator().manual_seed(1)
    for cls, cfg, spec in ((OctAttention, cfg_oct(), W.octattn_spec()), (EHEM, cfg_ehem(), W.ehem_spec(19))):
        m = cls(cfg).cuda()
        if cls is EHEM:
            data = torch.stack((torch.randint(1, 12, (1, 64, 4), generator=g), torch.randint(1, 9, (1, 64, 4), generator=g),
                                torch.randint(0, 255, (1, 64, 4), generator=g)), -1).cuda()
            pos = torch.rand((1, 3, 64), generator=g).cuda()
        else:
            data = torch.stack((torch.randint(0, 255, (1, 64, 4), generator=g), torch.randint(1, 12, (1, 64, 4), generator=g),
                                torch.randint(1, 9, (1, 64, 4), generator=g)), -1).cuda()
            pos = torch.rand((1, 64, 4, 3), generator=g).cuda()
        run = lambda mod: [t.clone() for t in (lambda o: o if isinstance(o, tuple) else (o,))(mod(data.clone(), pos))]
        before = run(m)
        m.load_state_dict(W.synth_state_dict(spec, seed=5, sharpen=True))        # in-place copy into the same storage
        after = run(m)
        fresh = cls(cfg, seed=5).cuda()
        want = run(fresh)
        assert not torch.equal(before[0], after[0])
        for a, b in zip(after, want):
            assert torch.equal(a, b), cls.__name__


@pytest.mark.parametrize("M,N,K,act,res", [(70001, 768, 256, "none", False), (75000, 1024, 256, "gelu", False),
                                           (90000, 256, 256, "none", True), (60000, 512, 192, "leaky", False)])
def test_cluster_multicast_is_bit_identical(M, N, K, act, res):
    """K <= 256 layers with many rows run in thread-block clusters that share the weight stream by TMA multicast (csrc/gemm_tc.cu,
    CL = 2 or 4): same MMAs in the same order as the single-CTA kernel, so the result must not change by a bit -- including the
    phantom M block of an odd tail and rows past M."""
    from scp_b200.ops import CudaOps
    cu = CudaOps(engine="f16x3")
    cu.lib.scp_gemm_cache_clear()
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * 0.1).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).cuda() if res else None
    outs = []
    for cl in (1, 2, 4):
        old = cu.lib.scp_set_gemm_cluster(cl)
        y = torch.full((M + 300, N), float("nan"), device="cuda")
        cu.linear(V(x), w, b, V(y[:M]), act=act, res=V(r) if res else None)
        torch.cuda.synchronize()
        cu.lib.scp_set_gemm_cluster(old)
        assert torch.isnan(y[M:]).all()                      # nothing written past the last row
        outs.append(y[:M].clone())
    ref = x.double() @ w.double().T + b.double()
    ref = {"none": lambda t: t, "leaky": lambda t: torch.nn.functional.leaky_relu(t, 0.01), "gelu": torch.nn.functional.gelu}[act](ref)
    if res:
        ref = ref + r.double()
    assert (outs[0].double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
