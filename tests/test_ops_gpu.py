"""Every CUDA operator of the entropy model (A8-A12) against the plain-torch fp32 statement of the same
operator (tests/emu_ops.py) on seeded random inputs.  fp32 SIMT engine: tolerance 1e-4 relative."""
import numpy as np
import pytest
import torch

from emu_ops import EmuOps

pytestmark = pytest.mark.gpu


def both(shape_rows, cols, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(shape_rows, cols, generator=g) * scale
    return c, c.cuda()


def V(t, col=0, ncol=None):
    return (t, col, t.shape[1] - col if ncol is None else ncol)


def close(a, b, tol=1e-4):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-6
    assert err <= tol * max(1.0, ref), (err, ref)


@pytest.fixture(scope="module")
def ops():
    from scp_b200.ops import CudaOps
    return CudaOps(engine="simt"), EmuOps()


@pytest.mark.parametrize("M,N,K,act,res,step", [
    (1, 255, 256, "none", False, 1), (37, 64, 80, "leaky", False, 1), (300, 16, 16, "leaky", False, 1),
    (5000, 768, 256, "none", False, 1), (3000, 256, 1024, "none", True, 1), (2500, 1024, 256, "gelu", False, 1),
    (1100, 128, 3, "none", False, 1), (700, 255, 512, "none", False, 1), (640, 240, 256, "relu", False, 2),
    (4096, 1024, 1280, "leaky", False, 1)])
def test_linear(ops, M, N, K, act, res, step):
    cu, em = ops
    x, xc = both(M * step, K + 8, 1)
    w, wc = both(N, K, 2, 0.1)
    b, bc = both(1, N, 3)
    r, rc = both(M, N, 4)
    y, yc = torch.zeros(M, N + 4), torch.zeros(M, N + 4, device="cuda")
    kw = dict(act=act, res=V(r) if res else None, row_step=step, row_off=step - 1, rows=M)
    em.linear(V(x, 4, K), w, b[0], V(y, 0, N), **kw)
    kw["res"] = V(rc) if res else None
    cu.linear(V(xc, 4, K), wc, bc[0], V(yc, 0, N), **kw)
    close(yc, y, 2e-5 * max(1, K // 64))


@pytest.mark.parametrize("M,C,res", [(1, 256, False), (1000, 256, False), (333, 512, False), (77, 600, True)])
def test_layernorm(ops, M, C, res):
    cu, em = ops
    x, xc = both(M, C, 1, 3.0)
    r, rc = both(M, C, 2)
    g, gc = both(1, C, 3)
    b, bc = both(1, C, 4)
    y, yc = torch.zeros(M, C), torch.zeros(M, C, device="cuda")
    em.layernorm(V(x), g[0], b[0], V(y), res=V(r) if res else None)
    cu.layernorm(V(xc), gc[0], bc[0], V(yc), res=V(rc) if res else None)
    close(yc, y, 1e-5)


def rand_ctx(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    c = torch.stack([torch.randint(0, 18, (n, 4), generator=g), torch.randint(0, 9, (n, 4), generator=g),
                     torch.randint(0, 256, (n, 4), generator=g)], 2).to(torch.uint8)
    return c


def test_ehem_embeddings(ops):
    cu, em = ops
    n = 1000
    ctx = rand_ctx(n)
    oe, oec = both(256, 16, 1)
    le, lec = both(19, 4, 2)
    qe, qec = both(9, 4, 3)
    y, yc = torch.zeros(n, 144), torch.zeros(n, 144, device="cuda")
    em.ehem_embed(ctx, oe, le, qe, V(y, 64, 80))
    cu.ehem_embed(ctx.cuda().reshape(n, 12), oec, lec, qec, V(yc, 64, 80))
    assert torch.equal(yc.cpu(), y)
    z, zc = torch.zeros(n // 2, 16), torch.zeros(n // 2, 16, device="cuda")
    em.ehem_embed_occ(ctx, oe, V(z))
    cu.ehem_embed_occ(ctx.cuda().reshape(n, 12), oec, V(zc))
    assert torch.equal(zc.cpu(), z)


@pytest.mark.parametrize("tensor_cores", [1, 0])
@pytest.mark.parametrize("d,lens,grid", [(3, [2, 700, 64, 37 + 1, 8192], True), (144, [600, 130, 2, 8192], False),
                                         (192, [2100, 129], False)])
def test_knn(ops, d, lens, grid, tensor_cores):
    cu, _ = ops
    old = cu.lib.scp_set_knn_engine(tensor_cores)
    offs = np.concatenate([[0], np.cumsum(lens)])
    g = torch.Generator().manual_seed(d)
    x = torch.randn(offs[-1], d, generator=g)
    if grid:                                     # integer-grid positions: many exact ties (SURVEY hard part)
        x = torch.randint(0, 40, (offs[-1], d), generator=g).float() / 40
    k = 20
    idx = cu.knn(V(x.cuda()), cu.seqs(list(offs)), k).cpu().long()
    cu.lib.scp_set_knn_engine(old)
    xd = x.double()
    for a, b in zip(offs[:-1], offs[1:]):
        s = xd[a:b]
        dist = torch.cdist(s, s) ** 2
        kk = min(k, b - a)
        ref = dist.topk(kk, dim=1, largest=False)[0]
        sel = idx[a:b]
        assert ((sel >= a) & (sel < b)).all()
        got = torch.gather(dist, 1, sel - a).sort(1)[0]
        if kk < k:      # short sequence: all points + repeats of self (distance 0)
            assert (got[:, : k - kk + 1] == 0).all()
            got = got[:, k - kk:]
            for r in range(b - a):
                assert len(set(sel[r].tolist())) == kk
        else:
            for r in range(0, b - a, 97):
                assert len(set(sel[r].tolist())) == k
        assert (got - ref).abs().max() < 1e-4 * max(1.0, ref.max().item())


@pytest.mark.parametrize("d", [144, 192])
def test_knn_is_independent_of_the_rest_of_the_batch(ops, d):
    """Lossless decoding needs the encoder (windows of whole frames in one ragged call) and the decoder (one level per call)
    to see identical neighbour sets: the indices of a window must not change when windows with a 4000x larger value range
    share the call (the fp16 hi/lo split of the Gram tiles is scaled per window, csrc/knn_tc.cu)."""
    cu, _ = ops
    g = torch.Generator().manual_seed(d)
    a = torch.randn(1500, d, generator=g) * 0.01
    big = torch.randn(3000, d, generator=g) * 40.0
    alone = cu.knn(V(a.cuda()), cu.seqs([0, 1500]), 20).cpu()
    x = torch.cat([big[:2000], a, big[2000:]]).cuda()
    inside = cu.knn(V(x), cu.seqs([0, 2000, 3500, 4500]), 20).cpu()
    assert torch.equal(inside[2000:3500] - 2000, alone)


def test_edge_gather(ops):
    cu, em = ops
    n, k = 900, 20
    for C in (64, 128, 256):
        uv, uvc = both(n, 2 * C, C)
        idx = torch.randint(0, n, (n, k), generator=torch.Generator().manual_seed(1)).int()
        s, sc = both(1, C, 5)
        t, tc = both(1, C, 6)
        y, yc = torch.zeros(n, C + 8), torch.zeros(n, C + 8, device="cuda")
        em.edge_gather_max(V(uv), C, idx, s[0], t[0], V(y, 8, C))
        y2c = torch.zeros(n, C + 4, device="cuda")
        cu.edge_gather_max(V(uvc), C, idx.cuda(), sc[0], tc[0], V(yc, 8, C), y2=V(y2c, 4, C))      # second view: same values
        close(yc, y, 1e-6)
        assert torch.equal(y2c[:, 4:], yc[:, 8:]) and (y2c[:, :4] == 0).all()


@pytest.mark.parametrize("tensor_cores", [2, 1, 0])          # 2 = 3xFP16 two CTAs/SM (default), 1 = 3xTF32, 0 = fp32 SIMT
@pytest.mark.parametrize("lens,shift", [([37], 0), ([37], 256), ([600, 2, 512], 256), ([1100], 0), ([1100, 300], 256),
                                        ([2048], 256), ([8192, 130], 0)])
def test_swin_attention(ops, lens, shift, tensor_cores):
    cu, em = ops
    old = cu.lib.scp_set_attn_engine(tensor_cores)
    offs = [0] + list(np.cumsum(lens))
    T = offs[-1]
    qkv, qkvc = both(T, 768, 1)
    b, bc = both(3, 256, 2)
    rel, relc = both(1023, 4, 3, 0.5)
    y, yc = torch.zeros(T, 256), torch.zeros(T, 256, device="cuda")
    em.swin_attention(V(qkv, 0, 256), V(qkv, 256, 256), V(qkv, 512, 256), b[0], b[1], b[2], rel, 4, em.seqs(offs), shift, V(y))
    cu.swin_attention(V(qkvc, 0, 256), V(qkvc, 256, 256), V(qkvc, 512, 256), bc[0], bc[1], bc[2], relc, 4,
                      cu.seqs(offs), shift, V(yc))
    cu.lib.scp_set_attn_engine(old)
    close(yc, y, 2e-5)


def test_pair_concat_upsample_copy(ops):
    cu, em = ops
    lens = [37, 2, 600, 1]
    offs = [0] + list(np.cumsum(lens))
    x, xc = both(offs[-1], 256, 1)
    se, sc = em.seqs(offs), cu.seqs(offs)
    y, yc = torch.zeros(se.half().total, 512), torch.zeros(se.half().total, 512, device="cuda")
    em.pair_concat(V(x), se, se.half(), V(y))
    cu.pair_concat(V(xc), sc, sc.half(), V(yc))
    assert torch.equal(yc.cpu(), y)
    h, hc = both(se.half().half().total, 256, 2)
    z, zc = torch.zeros(offs[-1], 1280), torch.zeros(offs[-1], 1280, device="cuda")
    em.upsample_cols(V(h), se.half().half(), se, 2, V(z, 512, 256))
    cu.upsample_cols(V(hc), sc.half().half(), sc, 2, V(zc, 512, 256))
    assert torch.equal(zc.cpu(), z)
    em.copy_cols(V(x), V(z, 1024, 256), row_step=2, row_off=1, rows=offs[-1] // 2)
    cu.copy_cols(V(xc), V(zc, 1024, 256), row_step=2, row_off=1, rows=offs[-1] // 2)
    assert torch.equal(zc.cpu(), z)


def test_octattn_ops(ops):
    cu, em = ops
    from scp_b200 import weights as W
    sd = W.synth_state_dict(W.octattn_spec(), 0, True)
    sdc = {k: v.cuda() for k, v in sd.items()}
    lens = [1024, 300, 5]
    offs = [0] + list(np.cumsum(lens))
    T = offs[-1]
    ctx = rand_ctx(T, 3)
    cpos = torch.randint(0, 4096, (T, 4, 3), generator=torch.Generator().manual_seed(2)).int()
    e, eu = torch.zeros(T, 600), torch.zeros(T, 600)
    ec, euc = torch.zeros(T, 600, device="cuda"), torch.zeros(T, 600, device="cuda")
    em.octattn_embed(ctx, cpos, 1 / 4096, 12, 12, em.seqs(offs), sd, e, eu)
    cu.octattn_embed(ctx.cuda().reshape(T, 12), cpos.cuda().reshape(T, 12), 1 / 4096, 12, 12, cu.seqs(offs), sdc, ec, euc)
    close(ec, e, 1e-6)
    close(euc, eu, 1e-6)
    a, ac = both(T, 3000, 7, 0.3)
    o, ou = torch.zeros(T, 600), torch.zeros(T, 600)
    oc, ouc = torch.zeros(T, 600, device="cuda"), torch.zeros(T, 600, device="cuda")
    em.octattn_attention(V(a, 1800, 600), V(a, 0, 600), V(a, 1200, 600), V(a, 600, 600), V(a, 2400, 600), 4, 150,
                         em.seqs(offs), V(o), V(ou))
    cu.octattn_attention(V(ac, 1800, 600), V(ac, 0, 600), V(ac, 1200, 600), V(ac, 600, 600), V(ac, 2400, 600), 4, 150,
                         cu.seqs(offs), V(oc), V(ouc))
    close(oc, o, 2e-5)
    close(ouc, ou, 2e-5)
