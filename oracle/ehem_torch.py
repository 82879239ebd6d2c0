"""CPU ORACLE (test infrastructure, NOT a product path): plain-PyTorch fp32 restatement of the forward
passes of SCP-EHEM and SCP-OctAttention in the reference's own formulation (explicit edge features,
unfused attention, materialised scores).  Pinned by tests/test_oracle_models.py against logits produced
by the UNMODIFIED reference (tests/golden/ehem_logits*.npz, octattn_logits.npz).

Follows: models/ehem.py:72-136; models/dgcnn.py:10-71,121-154; models/swin_transformer.py:217-222,
322-367,406-501,583-871; models/oct_attention.py:48-99; models/attention_model.py:6-155.

``knn`` is pluggable because the reference's neighbour choice under EXACT distance ties is whatever
``torch.topk`` (libstdc++ partial_sort / nth_element) happens to return; ``knn_torch_topk`` restates that,
``knn_canonical`` is the deterministic rule of the CUDA path (exact float64 distance of the float32 rows,
ties -> lowest index)."""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# kNN variants
# ---------------------------------------------------------------------------------------------
def knn_torch_topk(x, k):
    """dgcnn.py:17-28.  x [C,N] -> idx [N,k]."""
    inner = -2 * torch.matmul(x.t(), x)
    xx = torch.sum(x ** 2, dim=0, keepdim=True)
    pd = -inner - xx - xx.t()
    return pd.topk(k=k, dim=-1)[1]


def knn_canonical(x, k):
    """Deterministic rule of scp_knn: the k nearest rows by the EXACT squared distance of the float32 rows (float64
    accumulation of (a-b)^2 in channel order), ties -> lowest index."""
    C, N = x.shape
    xd = x.double().t().contiguous()
    d = torch.zeros(N, N, dtype=torch.float64)
    for c in range(C):
        diff = xd[:, None, c] - xd[None, :, c]
        d = d + diff * diff
    return torch.argsort(d, dim=1, stable=True)[:, :k]


# ---------------------------------------------------------------------------------------------
# EHEM
# ---------------------------------------------------------------------------------------------
def _mlp(sd, pre, x):
    x = F.leaky_relu(F.linear(x, sd[f"{pre}.0.weight"], sd[f"{pre}.0.bias"]))
    x = F.leaky_relu(F.linear(x, sd[f"{pre}.2.weight"], sd[f"{pre}.2.bias"]))
    return F.linear(x, sd[f"{pre}.4.weight"], sd[f"{pre}.4.bias"])


def _edge_conv(sd, pre, x, k, knn):
    """x [N,C]: gather k neighbours, feature [nbr - x, x], 1x1 conv (no bias), BatchNorm (eval),
    LeakyReLU(0.2), max over k  (dgcnn.py:48-71,79-87,133-134)."""
    idx = knn(x.t().contiguous(), k)
    nbr = x[idx]                                             # [N,k,C]
    ctr = x[:, None, :].expand(-1, k, -1)
    f = torch.cat((nbr - ctr, ctr), 2)                       # [N,k,2C]
    y = F.linear(f, sd[f"{pre}.0.weight"][:, :, 0, 0])
    y = (y - sd[f"{pre}.1.running_mean"]) / torch.sqrt(sd[f"{pre}.1.running_var"] + 1e-5) * sd[f"{pre}.1.weight"] + sd[f"{pre}.1.bias"]
    return F.leaky_relu(y, 0.2).max(1)[0]


def _geo_features(sd, data11, pos, knn):
    g = "geo_feat_generator"
    occ, level, octant = data11[:, 2::3], data11[:, 0::3], data11[:, 1::3]
    n = data11.shape[0]
    x = torch.cat((sd[f"{g}.occ_enc.weight"][occ].reshape(n, -1), sd[f"{g}.level_enc.weight"][level].reshape(n, -1),
                   sd[f"{g}.octant_enc.weight"][octant].reshape(n, -1)), 1)
    k = min(20, n)
    p1 = _edge_conv(sd, f"{g}.conv1", pos, k, knn)
    p2 = _edge_conv(sd, f"{g}.conv2", torch.cat((p1, x), 1), k, knn)
    x = _mlp(sd, f"{g}.mlp2", x)
    p3 = _edge_conv(sd, f"{g}.conv3", torch.cat((p2, x), 1), k, knn)
    x = _mlp(sd, f"{g}.mlp3", x)
    ec = _mlp(sd, f"{g}.edge_mlp1", torch.cat((p1, p2, p3), 1))
    ec = _mlp(sd, f"{g}.edge_mlp2", torch.cat((p3, ec), 1))
    return torch.cat((x, ec), 1)


def _window_attention(sd, pre, h, q_src, shift):
    """One SwinLayer attention (swin_transformer.py:631-697): LN -> zero-pad to 512-multiple -> roll ->
    windows -> QK^T/8 + rel-pos bias (+ -100 mask on the last window) -> softmax -> PV -> unroll -> crop."""
    ws, heads = 512, 4
    S = h.shape[0]
    a = f"{pre}.attention.self"
    ln = lambda t: F.layer_norm(t, (256,), sd[f"{pre}.layernorm_before.weight"], sd[f"{pre}.layernorm_before.bias"], 1e-5)

    def windows(t):
        t = F.pad(ln(t), (0, 0, 0, (-S) % ws))
        if shift:
            t = torch.roll(t, -shift, 0)
        return t.reshape(-1, ws, 256)
    kv = windows(h)
    qw = kv if q_src is None else windows(q_src)
    split = lambda t: t.reshape(t.shape[0], ws, heads, 64).permute(0, 2, 1, 3)
    Q = split(F.linear(qw, sd[f"{a}.query.weight"], sd[f"{a}.query.bias"]))
    K = split(F.linear(kv, sd[f"{a}.key.weight"], sd[f"{a}.key.bias"]))
    Vv = split(F.linear(kv, sd[f"{a}.value.weight"], sd[f"{a}.value.bias"]))
    sc = Q @ K.transpose(-1, -2) / math.sqrt(64)
    i = torch.arange(ws)
    sc = sc + sd[f"{a}.relative_position_bias_table"][(i[:, None] - i[None, :]) + ws - 1].permute(2, 0, 1)[None]
    if shift:
        Sp = kv.shape[0] * ws
        region = torch.zeros(Sp)
        region[Sp - ws:Sp - shift] = 1
        region[Sp - shift:] = 2
        rw = region.reshape(-1, ws)
        mask = (rw[:, None, :] != rw[:, :, None]).float() * -100.0
        sc = sc + mask[:, None]
    o = (torch.softmax(sc, -1) @ Vv).permute(0, 2, 1, 3).reshape(-1, 256)
    if shift:
        o = torch.roll(o, shift, 0)
    return o[:S]


def _swin_layer(sd, pre, h, q_src, shift):
    att = _window_attention(sd, pre, h, q_src, shift)
    h = h + F.linear(att, sd[f"{pre}.attention.output.dense.weight"], sd[f"{pre}.attention.output.dense.bias"])
    y = F.layer_norm(h, (256,), sd[f"{pre}.layernorm_after.weight"], sd[f"{pre}.layernorm_after.bias"], 1e-5)
    y = F.gelu(F.linear(y, sd[f"{pre}.intermediate.dense.weight"], sd[f"{pre}.intermediate.dense.bias"]))
    return h + F.linear(y, sd[f"{pre}.output.dense.weight"], sd[f"{pre}.output.dense.bias"])


def _merge(sd, pre, h):
    if h.shape[0] % 2:
        h = F.pad(h, (0, 0, 0, 1))
    y = torch.cat((h[0::2], h[1::2]), 1)
    y = F.layer_norm(y, (512,), sd[f"{pre}.norm.weight"], sd[f"{pre}.norm.bias"], 1e-5)
    return F.linear(y, sd[f"{pre}.reduction.weight"])


def _swin_encoder(sd, enc, depths, h, query=None):
    """Returns the concat_states tensor (ehem.py:75-86): stage outputs upsampled x2^i to the finest length."""
    n = h.shape[0]
    outs = []
    for i, depth in enumerate(depths):
        for j in range(depth):
            h = _swin_layer(sd, f"{enc}.layers.{i}.blocks.{j}", h, query, 0 if j % 2 == 0 else 256)
        outs.append(h[torch.arange(n) >> i])
        if i < len(depths) - 1:
            pre = f"{enc}.layers.{i}.downsample"
            h = _merge(sd, pre, h)
            if query is not None:
                query = _merge(sd, pre, query)
    return torch.cat(outs, 1)


@torch.no_grad()
def ehem_forward(sd, data, pos, knn=knn_torch_topk):
    """data int64 [csz,4,3] (level, octant, occ), pos float32 [3,csz] -> (logits1, logits2) like
    EHEM.forward(enc=True) with batch size 1 (ehem.py:88-136)."""
    csz = data.shape[0]
    padded = csz % 2 == 1
    if padded:
        pad = torch.zeros_like(data[:1])
        pad[:, :, 2] = 255
        data = torch.cat((data, pad), 0)
        pos = torch.cat((pos, torch.zeros_like(pos[:, :1])), 1)
    pre_occ = data[0::2, -1, -1]
    feat = _geo_features(sd, data.reshape(data.shape[0], 12)[:, :-1], pos.t().contiguous(), knn)
    feat_a = _mlp(sd, "ancient_mlp", _swin_encoder(sd, "swin_self_transformer", (4, 4, 4, 4, 2), feat))
    a1, a2 = feat_a[0::2], feat_a[1::2]
    logits1 = _mlp(sd, "prob_pred_mlp1", a1)
    pre = torch.cat((_mlp(sd, "pre_occ_mlp", sd["geo_feat_generator.occ_enc.weight"][pre_occ]), _mlp(sd, "pre_attn_mlp", a1)), 1)
    cross = _swin_encoder(sd, "swin_cross_transformer", (2, 2, 1, 1), pre, query=a2)
    logits2 = _mlp(sd, "prob_pred_mlp2", torch.cat((cross, a2), 1))
    if padded:
        logits2 = logits2[:-1]
    return logits1, logits2


# ---------------------------------------------------------------------------------------------
# OctAttention
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def octattn_forward(sd, data, pos, train_type="kitti", max_octree_level=12, heads=4):
    """data int64 [csz,4,3] (occ, level, octant), pos float32 [csz,4,3] -> logits [csz,255]
    (oct_attention.py:48-83, attention_model.py:58-155)."""
    csz = data.shape[0]
    occ, level, octant = data[..., 0], data[..., 1].clone(), data[..., 2]
    level = level - torch.clamp(level[:, -1:] - (10 if train_type == "obj" else 12), min=0)
    level = level.clamp(0, max_octree_level)
    oe = sd["occ_enc.weight"][occ]
    ou = oe.clone()
    ou[:, -1] = sd["occ_enc.weight"][255]
    rest = torch.cat((sd["level_enc.weight"][level], sd["octant_enc.weight"][octant],
                      F.linear(pos, sd["abs_pos_enc.weight"], sd["abs_pos_enc.bias"])), 2)
    pe = sd["transformer_encoder.position_enc.pe"][:csz] if "transformer_encoder.position_enc.pe" in sd else 0.0   # cfg.model.pos_embed (attention_model.py:142-149)
    e = torch.cat((oe, rest), 2).reshape(csz, 600) * math.sqrt(600) + pe
    u = torch.cat((ou, rest), 2).reshape(csz, 600) * math.sqrt(600) + pe
    mask = sd["mask"][:csz, :csz]
    hd = 600 // heads
    sp = lambda t: t.reshape(csz, heads, hd).permute(1, 0, 2)
    for i in range(3):
        p = f"transformer_encoder.layers.{i}"
        lin = lambda nm, t: F.linear(t, sd[f"{p}.attn.{nm}.weight"], sd[f"{p}.attn.{nm}.bias"])
        k, ku, qu, v, vu = sp(lin("mlp_key", e)), sp(lin("mlp_key", u)), sp(lin("mlp_query", u)), sp(lin("mlp_value", e)), sp(lin("mlp_value", u))
        sc = qu @ k.transpose(1, 2) / math.sqrt(hd)
        out = (torch.softmax(sc + mask, -1) @ v).permute(1, 0, 2).reshape(csz, 600)
        eye = torch.eye(csz)
        scu = (1 - eye) * sc + torch.diag_embed((qu * ku).sum(2) / math.sqrt(hd))
        au = torch.softmax(scu + mask, -1)
        outu = (((1 - eye) * au) @ v + torch.diagonal(au, dim1=1, dim2=2)[..., None] * vu).permute(1, 0, 2).reshape(csz, 600)
        res = []
        for src, att in ((e, out), (u, outu)):
            x = F.layer_norm(att + src, (600,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5)
            f = F.linear(F.relu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])), sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
            res.append(F.layer_norm(x + f, (600,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5))
        e, u = res
    return F.linear(F.relu(F.linear(u, sd["decoder0.weight"], sd["decoder0.bias"])), sd["decoder1.weight"], sd["decoder1.bias"])
