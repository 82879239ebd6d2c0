"""CUDA EHEM / OctAttention forward against the golden logits of the unmodified reference (CPU fp32).
Tolerance (BASELINE.json north_star): PMF max-abs <= 1e-3 for the fp32 path."""
import numpy as np
import pytest
import torch

from conftest import golden
from test_models_cpu import cfg_ehem, cfg_oct

pytestmark = pytest.mark.gpu
PMF_TOL = 1e-3


def pmf_err(a, b):
    return (torch.softmax(a.float().cpu(), -1) - torch.softmax(torch.as_tensor(b).float(), -1)).abs().max().item()


@pytest.fixture(scope="module")
def ehem():
    from scp_b200.models import EHEM
    return EHEM(cfg_ehem()).cuda()


@pytest.mark.parametrize("tag", ["n1", "n2", "n37", "n600", "n1100"])
def test_ehem_vs_reference_logits(ehem, tag):
    g = golden("ehem_logits.npz")
    data = torch.from_numpy(g[f"{tag}_data"].astype(np.int64))[None].cuda()
    pos = torch.from_numpy(g[f"{tag}_pos"])[None].cuda()
    l1, l2 = ehem(data, pos)
    assert tuple(l1.shape[1:]) == g[f"{tag}_logits1"].shape and tuple(l2.shape[1:]) == g[f"{tag}_logits2"].shape
    assert pmf_err(l1[0], g[f"{tag}_logits1"]) < PMF_TOL
    if l2.shape[1]:
        assert pmf_err(l2[0], g[f"{tag}_logits2"]) < PMF_TOL


def test_ehem_full_window_vs_reference(ehem):
    g = golden("ehem_logits_full.npz")
    data = torch.from_numpy(g["data"].astype(np.int64))[None].cuda()
    pos = torch.from_numpy(g["pos"])[None].cuda()
    l1, l2 = ehem(data, pos)
    e1 = pmf_err(l1[0, ::16], g["logits1_s16"])
    e2 = pmf_err(l2[0, ::16], g["logits2_s16"])
    m1 = torch.softmax(l1[0], 1).max(1)[0].cpu().numpy()
    print("full-window pmf err", e1, e2, "max-pmf err", np.abs(m1 - g["pmf1_max"]).max())
    assert e1 < PMF_TOL and e2 < PMF_TOL
    assert np.abs(m1 - g["pmf1_max"]).max() < PMF_TOL


@pytest.mark.parametrize("tag", ["w0", "w3", "tail"])
def test_octattn_vs_reference_logits(tag):
    from scp_b200.models import OctAttention
    g = golden("octattn_logits.npz")
    m = OctAttention(cfg_oct()).cuda()
    data = torch.from_numpy(g[f"{tag}_data"].astype(np.int64))[None].cuda()
    pos = torch.from_numpy(g[f"{tag}_pos"])[None].cuda()
    out = m(data, pos)[0, ::2]
    assert pmf_err(out, g[f"{tag}_logits_s2"]) < PMF_TOL
