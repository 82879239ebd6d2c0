"""Pins the torch fp32 model oracle (oracle/ehem_torch.py) to logits of the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ehem_torch as O
from scp_b200 import weights as W


@pytest.fixture(scope="module")
def sd_ehem():
    return W.synth_state_dict(W.ehem_spec(19), 0, True)


@pytest.mark.parametrize("tag", ["n1", "n2", "n37", "n600", "n1100", "j600", "j1100"])
def test_ehem_oracle_vs_reference(sd_ehem, tag):
    g = golden("ehem_logits.npz")
    l1, l2 = O.ehem_forward(sd_ehem, torch.from_numpy(g[f"{tag}_data"].astype(np.int64)), torch.from_numpy(g[f"{tag}_pos"]))
    assert np.abs(l1.numpy() - g[f"{tag}_logits1"]).max() < 2e-3          # logits, not PMFs
    assert (torch.softmax(l1, 1) - torch.softmax(torch.from_numpy(g[f"{tag}_logits1"]), 1)).abs().max() < 1e-5
    if g[f"{tag}_logits2"].shape[0]:
        assert (torch.softmax(l2, 1) - torch.softmax(torch.from_numpy(g[f"{tag}_logits2"]), 1)).abs().max() < 1e-5


@pytest.mark.parametrize("tag", ["w0", "w3", "tail"])
def test_octattn_oracle_vs_reference(tag):
    g = golden("octattn_logits.npz")
    sd = W.synth_state_dict(W.octattn_spec(), 0, True)
    out = O.octattn_forward(sd, torch.from_numpy(g[f"{tag}_data"].astype(np.int64)), torch.from_numpy(g[f"{tag}_pos"]))
    ref = torch.from_numpy(g[f"{tag}_logits_s2"])
    assert (torch.softmax(out[::2], 1) - torch.softmax(ref, 1)).abs().max() < 1e-5


def test_canonical_knn_only_differs_on_ties(sd_ehem):
    g = golden("ehem_logits.npz")
    pos = torch.from_numpy(g["j600_pos"])
    a = torch.sort(O.knn_torch_topk(pos, 20), 1)[0]
    b = torch.sort(O.knn_canonical(pos, 20), 1)[0]
    assert torch.equal(a, b)                                   # no ties -> identical neighbour sets
