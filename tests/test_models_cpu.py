"""Host-side orchestration of scp_b200.models (EHEM / OctAttention) checked on CPU against the golden logits of
the UNMODIFIED reference, with the operator contracts emulated in plain torch (tests/emu_ops.py)."""
import types

import numpy as np
import pytest
import torch

from conftest import golden
from emu_ops import EmuOps


def cfg_ehem():
    NS = types.SimpleNamespace
    return NS(model=NS(context_size=8192, token_num=255, max_level=19), train=NS(type="kitti"), data=NS(extra_pos=False))


def cfg_oct():
    NS = types.SimpleNamespace
    return NS(model=NS(max_octree_level=12, context_size=1024, token_num=255, layer_num=3, head_num=4,
                       abs_pos_embed_dim=12, occ_embed_dim=128, level_embed_dim=6, octant_embed_dim=4,
                       hidden_dimension=300), train=NS(type="kitti", dropout=0.0))


def test_state_dict_names_match_spec():
    from scp_b200.models import EHEM, OctAttention
    m = EHEM(cfg_ehem(), ops=EmuOps())
    assert list(m.state_dict().keys()) == [n for n, _, _ in m.spec]
    assert len(m.state_dict()) == 528
    o = OctAttention(cfg_oct(), ops=EmuOps())
    assert list(o.state_dict().keys()) == [n for n, _, _ in o.spec]
    assert len(o.state_dict()) == 53


@pytest.mark.parametrize("tag", ["n1", "n2", "n37", "n600", "n1100"])
def test_ehem_orchestration_vs_reference_logits(tag):
    from scp_b200.models import EHEM
    g = golden("ehem_logits.npz")
    m = EHEM(cfg_ehem(), ops=EmuOps())
    data = torch.from_numpy(g[f"{tag}_data"].astype(np.int64))[None]
    pos = torch.from_numpy(g[f"{tag}_pos"])[None]
    l1, l2 = m(data, pos)
    r1, r2 = g[f"{tag}_logits1"], g[f"{tag}_logits2"]
    assert l1.shape[1:] == r1.shape and l2.shape[1:] == r2.shape
    p1, q1 = torch.softmax(l1[0], 1).numpy(), torch.softmax(torch.from_numpy(r1), 1).numpy()
    assert np.abs(p1 - q1).max() < 1e-3
    if r2.shape[0]:
        p2, q2 = torch.softmax(l2[0], 1).numpy(), torch.softmax(torch.from_numpy(r2), 1).numpy()
        assert np.abs(p2 - q2).max() < 1e-3


def test_ehem_ragged_batch_equals_single_windows():
    from scp_b200.models import EHEM
    g = golden("ehem_logits.npz")
    m = EHEM(cfg_ehem(), ops=EmuOps())
    tags = ["n2", "n600", "n37"]
    ctxs, poss, offs = [], [], [0]
    for t in tags:
        d = torch.from_numpy(g[f"{t}_data"].astype(np.int64)).to(torch.uint8)
        p = torch.from_numpy(g[f"{t}_pos"]).T
        if len(d) % 2:
            pad = torch.zeros_like(d[:1]); pad[:, :, 2] = 255
            d = torch.cat([d, pad]); p = torch.cat([p, torch.zeros_like(p[:1])])
        ctxs.append(d); poss.append(p); offs.append(offs[-1] + len(d))
    l1, l2 = m.forward_ragged(torch.cat(ctxs), torch.cat(poss).contiguous(), offs)
    for t, a, b in zip(tags, offs[:-1], offs[1:]):
        r1 = g[f"{t}_logits1"]
        got = l1[a // 2:b // 2].numpy()
        assert np.abs(torch.softmax(torch.from_numpy(got), 1).numpy() - torch.softmax(torch.from_numpy(r1), 1).numpy()).max() < 1e-3


@pytest.mark.parametrize("tag", ["w0", "w3", "tail"])
def test_octattn_orchestration_vs_reference_logits(tag):
    from scp_b200.models import OctAttention
    g = golden("octattn_logits.npz")
    m = OctAttention(cfg_oct(), ops=EmuOps())
    data = torch.from_numpy(g[f"{tag}_data"].astype(np.int64))[None]
    pos = torch.from_numpy(g[f"{tag}_pos"])[None]
    out = m(data, pos)[0, ::2]
    ref = torch.from_numpy(g[f"{tag}_logits_s2"])
    assert out.shape == ref.shape
    assert (torch.softmax(out, 1) - torch.softmax(ref, 1)).abs().max() < 1e-3


@pytest.mark.parametrize("tag", ["n37", "n600"])
def test_explained_parity_criterion_on_emulated_ops(tag):
    """tests/parity_explain.py itself (the criterion the -m gpu model tests apply), run on the CPU emulation."""
    from parity_explain import check_explained_parity
    from scp_b200 import weights as W
    from scp_b200.models import EHEM
    g = golden("ehem_logits.npz")
    m = EHEM(cfg_ehem(), ops=EmuOps())
    rep = check_explained_parity(m, W.synth_state_dict(W.ehem_spec(19), 0, True),
                                 torch.from_numpy(g[f"{tag}_data"].astype(np.int64)), torch.from_numpy(g[f"{tag}_pos"]),
                                 g[f"{tag}_logits1"], g[f"{tag}_logits2"], what=tag)
    assert rep["arith_max"] < 1e-4 and len(rep["stages"]) == 3 and rep["direct_max"] < 1e-4


def test_octattn_without_positional_encoding():
    """cfg.model.pos_embed False (attention_model.py:142-149): no ``position_enc.pe`` in the state_dict (a reference checkpoint
    trained that way loads strictly) and nothing is added to the embeddings."""
    from oracle import ehem_torch as O
    from scp_b200 import weights as W
    from scp_b200.models import OctAttention
    cfg = cfg_oct()
    cfg.model.pos_embed = False
    m = OctAttention(cfg, ops=EmuOps())
    assert "transformer_encoder.position_enc.pe" not in m.state_dict() and len(m.state_dict()) == 52
    g = golden("octattn_logits.npz")
    data = torch.from_numpy(g["w3_data"].astype(np.int64))[:300]
    pos = torch.from_numpy(g["w3_pos"])[:300]
    out = m(data[None].clone(), pos[None])[0]
    sd = W.synth_state_dict(W.octattn_spec(pos_embed=False), 0, True)
    ref = O.octattn_forward(sd, data, pos)
    assert (torch.softmax(out, 1) - torch.softmax(ref, 1)).abs().max() < 1e-4
    with_pe = O.octattn_forward(W.synth_state_dict(W.octattn_spec(), 0, True), data, pos)
    assert (torch.softmax(out, 1) - torch.softmax(with_pe, 1)).abs().max() > 1e-3
