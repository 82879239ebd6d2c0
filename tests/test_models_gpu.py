"""CUDA EHEM / OctAttention forward against the golden logits of the unmodified reference (CPU fp32).
Tolerance (BASELINE.json north_star): PMF max-abs <= 1e-3 for the fp32 path, on every row; the only discontinuity of the
model (kNN ties) is handled by the explained-rows criterion of tests/parity_explain.py, not by an allowance."""
import numpy as np
import pytest
import torch

from conftest import golden
from parity_explain import check_explained_parity
from test_models_cpu import cfg_ehem, cfg_oct

pytestmark = pytest.mark.gpu
PMF_TOL = 1e-3


def pmf_err(a, b):
    return (torch.softmax(a.float().cpu(), -1) - torch.softmax(torch.as_tensor(b).float(), -1)).abs().max().item()


@pytest.fixture(scope="module")
def ehem():
    from scp_b200.models import EHEM
    return EHEM(cfg_ehem()).cuda()


@pytest.fixture(scope="module")
def sd():
    from scp_b200 import weights as W
    return W.synth_state_dict(W.ehem_spec(19), 0, True)


@pytest.mark.parametrize("tag", ["n1", "n2", "n37", "j600", "j1100", "n600", "n1100"])
def test_ehem_vs_reference_logits(ehem, sd, tag):
    """EHEM.forward against the unmodified reference's logits, by the explained-rows criterion of tests/parity_explain.py:
    (A) every PMF row within 1e-3 of the oracle run on the device's neighbour sets, (B) those sets are k-nearest sets up to
    float32 ties, (C) windows without near-ties match the reference's own logits within 1e-3 on every row.  ``j*`` = jittered
    (tie-free) positions, ``n*`` = octree grid positions (exact ties at the k-th neighbour on ~1-3 % of the rows)."""
    g = golden("ehem_logits.npz")
    data = torch.from_numpy(g[f"{tag}_data"].astype(np.int64))
    pos = torch.from_numpy(g[f"{tag}_pos"])
    l1, l2 = ehem(data[None].cuda(), pos[None].cuda())
    assert tuple(l1.shape[1:]) == g[f"{tag}_logits1"].shape and tuple(l2.shape[1:]) == g[f"{tag}_logits2"].shape
    check_explained_parity(ehem, sd, data, pos, g[f"{tag}_logits1"], g[f"{tag}_logits2"], what=tag)


@pytest.mark.parametrize("jitter", [True, False])
def test_ehem_full_window_vs_reference(ehem, sd, jitter):
    """One full 8192-token context window (the bench's window size), jittered and gridded positions."""
    g = golden("ehem_logits_full.npz")
    j = golden("ehem_logits_full_jit.npz")
    data = torch.from_numpy(g["data"].astype(np.int64))
    pos = torch.from_numpy(j["pos"] if jitter else g["pos"])
    r = j if jitter else g
    check_explained_parity(ehem, sd, data, pos, r["logits1_s16"], r["logits2_s16"], ref_slice=slice(None, None, 16),
                           what="full 8192 window, " + ("jittered" if jitter else "grid") + " positions")


def test_ehem_ragged_batch_equals_single_windows(ehem):
    """Batch invariance (what the encoder's ragged batches and the decoder's level-wise calls rely on): a window's logits do
    not depend on what else is in the call -- bit-identical."""
    g = golden("ehem_logits.npz")
    tags = ["n2", "n600", "n37", "j1100"]
    ctxs, poss, offs, single = [], [], [0], []
    for t in tags:
        d = torch.from_numpy(g[f"{t}_data"].astype(np.int64))
        p = torch.from_numpy(g[f"{t}_pos"])
        single.append(ehem(d[None].cuda(), p[None].cuda()))
        d8, pt = d.to(torch.uint8), p.T
        if len(d8) % 2:
            pad = torch.zeros_like(d8[:1]); pad[:, :, 2] = 255
            d8 = torch.cat([d8, pad]); pt = torch.cat([pt, torch.zeros_like(pt[:1])])
        ctxs.append(d8); poss.append(pt); offs.append(offs[-1] + len(d8))
    l1, l2 = ehem.forward_ragged(torch.cat(ctxs).cuda(), torch.cat(poss).contiguous().cuda(), offs)
    for (s1, s2), a, b in zip(single, offs[:-1], offs[1:]):
        assert torch.equal(l1[a // 2:b // 2], s1[0]) and torch.equal(l2[a // 2:a // 2 + s2.shape[1]], s2[0])


@pytest.mark.parametrize("tag", ["w0", "w3", "tail"])
def test_octattn_vs_reference_logits(tag):
    from scp_b200.models import OctAttention
    g = golden("octattn_logits.npz")
    m = OctAttention(cfg_oct()).cuda()
    data = torch.from_numpy(g[f"{tag}_data"].astype(np.int64))[None].cuda()
    pos = torch.from_numpy(g[f"{tag}_pos"])[None].cuda()
    out = m(data, pos)[0, ::2]
    assert pmf_err(out, g[f"{tag}_logits_s2"]) < PMF_TOL
