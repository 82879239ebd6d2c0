"""Decode path: bitstream -> occupancy codes -> voxels -> points.

Replaces the per-node Python of the reference's ``decode_ehem.py`` (decodeOct :56-188) and
``decode_ehem_mullevel.py`` (sub_decode :56-176, decodeOct :179-206) and ``DeOctree`` (Octree.py:68-99):

  level L nodes (cell origin, octant, 3 ancestor rows)        device state, rebuilt level by level
      --> context windows of encode.py:112-115                 (scp_gather_windows)
      --> EHEM phase 1 for ALL windows of the level at once    (models.EHEM.phase1; needs the ancestors only)
      --> softmax -> integer CDF rows                          (scp_pmf_to_cdf)
      --> per window, in coding order: decode the even nodes (host range decoder), EHEM phase 2 on that window with
          the decoded occupancies, decode the odd nodes        (scp_range_decode, models.EHEM.phase2)
      --> children of the decoded occupancy bytes              (child expansion = decode_ehem.py:116-140)

The decoder feeds the model EXACTLY what the encoder fed it (same context bytes, same float32 positions, same
operators), so the CDFs are bit-identical and the stream decodes losslessly; that identity is what
tests/test_roundtrip_gpu.py checks at full size.  Differences to the reference's decoder, on purpose:
* positions are normalised with the level's (min, max) pair the encoder stored (``FrameResult.pos_mm``, the
  ``.dat`` side file of encode.py:150) exactly as encode_dataset_ehem.py:70-72 does; decode_ehem.py:42-54 re-derives
  them from the parent's normalised float and only uses the max;
* termination comes from the number of levels in the header, not from the length of the original sequence
  (decode_ehem.py:66 reads it from the uncompressed .npy "for checking").
The child expansion and the window bookkeeping are torch index operations on the device (plumbing); the entropy
model, the CDF table and the range decoder are the library's.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import _lib, coder


@dataclass
class DecodedFrame:
    occ: List[np.ndarray] = field(default_factory=list)        # per sub-octree: occupancy bytes 1..255, BFS order (coded rows)
    voxels: List[np.ndarray] = field(default_factory=list)     # per sub-octree: int64 [V,3] quantised coordinates, Morton order
    depths: List[int] = field(default_factory=list)
    n_symbols: int = 0


class Decoder:
    """``Decoder(model, lidar_level, mode, mullevel)`` then ``decode(frame_result)`` for a ``FrameResult`` of
    ``Encoder.encode`` (bitstream + the header fields the reference puts in the file name and the .dat file)."""

    def __init__(self, model, lidar_level=12, mode="spher", mullevel=False, kind="kitti", max_tokens=1 << 19):
        if model.__class__.__name__ != "EHEM":
            raise NotImplementedError("decode path: EHEM only (the reference's decode.py for OctAttention is out of scope)")
        self.model = model
        self.level = lidar_level
        self.mode = mode
        self.mullevel = mullevel
        self.kind = kind
        self.context = model.cfg.model.context_size
        self.max_tokens = max_tokens
        self.lib = _lib.require_device()
        self.dev = torch.device("cuda")

    # ------------------------------------------------------------------------------------------------------------
    def _windows(self, n):
        starts = np.arange(0, n, self.context, dtype=np.int64)
        lens = np.minimum(self.context, n - starts).astype(np.int32)
        toks = np.concatenate([[0], np.cumsum(lens + (lens & 1))]).astype(np.int64)
        return starts, lens, toks

    def _decode_level(self, dec, ctx, pos_norm, n_code):
        """ctx uint8 [N,4,3] (self occupancy column = 255), pos_norm float32 [N,3]; decodes the first n_code nodes
        (the windows of encode.py:112-115) and returns their occupancy symbols 0..254 (host int16 [n_code])."""
        dev, st = self.dev, _lib.stream_ptr()
        starts, lens, toks = self._windows(n_code)
        out = np.empty(n_code, np.int16)
        ctx = ctx.contiguous()
        pos_norm = pos_norm.contiguous()
        w0 = 0
        while w0 < len(starts):                                   # chunks of windows bounded by max_tokens, like the encoder
            w1 = w0 + 1
            while w1 < len(starts) and toks[w1 + 1] - toks[w0] <= self.max_tokens:
                w1 += 1
            T = int(toks[w1] - toks[w0])
            ctxp = torch.empty((T, 4, 3), dtype=torch.uint8, device=dev)
            posp = torch.empty((T, 3), dtype=torch.float32, device=dev)
            re = torch.empty(T // 2, dtype=torch.int64, device=dev)
            ro = torch.empty(T // 2, dtype=torch.int64, device=dev)
            tk = np.ascontiguousarray(toks[w0:w1] - toks[w0])
            _lib.check(self.lib.scp_gather_windows(_lib.ptr(ctx), _lib.ptr(pos_norm), _lib.ptr(np.ascontiguousarray(starts[w0:w1])),
                                                   _lib.ptr(np.ascontiguousarray(lens[w0:w1])), _lib.ptr(tk), w1 - w0,
                                                   _lib.ptr(ctxp), _lib.ptr(posp), _lib.ptr(re), _lib.ptr(ro), st),
                       "scp_gather_windows")
            offs = [int(x) for x in np.append(tk, T)]
            feat_a, l1 = self.model.phase1(ctxp, posp, offs)
            cdf1 = coder.pmf_to_cdf(l1, is_logits=True, want_cdf=True)["cdf"].cpu().numpy()
            for w in range(w0, w1):
                t0, ln, s = int(tk[w - w0]), int(lens[w]), int(starts[w])
                ne, no, Tw = (ln + 1) // 2, ln // 2, ln + (ln & 1)
                sym_e = dec.decode(cdf1[t0 // 2: t0 // 2 + ne])
                out[s: s + ln: 2] = sym_e
                if no == 0:
                    continue
                cw = ctxp[t0: t0 + Tw]
                cw[0: 2 * ne: 2, 3, 2] = torch.from_numpy(sym_e.astype(np.uint8)).to(dev)
                l2 = self.model.phase2(cw, feat_a[t0: t0 + Tw], [0, Tw])
                cdf2 = coder.pmf_to_cdf(l2[:no].contiguous(), is_logits=True, want_cdf=True)["cdf"].cpu().numpy()
                out[s + 1: s + ln: 2] = dec.decode(cdf2)
            w0 = w1
        return out

    def _decode_tree(self, dec, depth, pos_mm, drop_last, pos_eps_last):
        dev = self.dev
        n = depth
        # level 1: the root (decode_ehem.py:79-82: ancestors (0,0,255), self (level 1, octant 1))
        pos = torch.zeros((1, 3), dtype=torch.int64, device=dev)
        anc = torch.zeros((1, 3, 3), dtype=torch.uint8, device=dev)
        anc[:, :, 2] = 255
        octant = torch.ones(1, dtype=torch.uint8, device=dev)
        bits = torch.tensor([[(d >> 2) & 1, (d >> 1) & 1, d & 1] for d in range(8)], dtype=torch.int64, device=dev)
        occ_all = []
        for L in range(1, n + 1):
            N = pos.shape[0]
            own = torch.stack([torch.full((N,), L, dtype=torch.uint8, device=dev), octant,
                               torch.full((N,), 255, dtype=torch.uint8, device=dev)], 1)
            ctx = torch.cat([anc, own[:, None, :]], 1)                      # [N,4,3], unclipped levels
            ctx_model = ctx
            if L == n and n > self.level:                                   # encode_dataset_ehem.py:86
                ctx_model = ctx.clone()
                ctx_model[:, :, 0] = torch.clamp(ctx_model[:, :, 0], max=self.level)
            mn, mx = pos_mm[L - 1]
            den = float(mx - mn) + (0.0 if (L == n and not pos_eps_last) else 1e-9)
            pos_norm = ((pos.to(torch.float64) - float(mn)) / den).to(torch.float32)     # encode_dataset_ehem.py:70-72
            n_code = N - 1 if (drop_last and L == n) else N
            if n_code == 1 and L > 1 and not self.mullevel:
                raise NotImplementedError("single-node level below the root: encode.py:123 codes the root again instead "
                                          "of this node (reference defect), the stream is not decodable")
            sym = self._decode_level(dec, ctx_model, pos_norm, n_code) if n_code > 0 else np.empty(0, np.int16)
            occ_all.append((sym + 1).astype(np.uint8))
            occ = torch.zeros(N, dtype=torch.int64, device=dev)              # a dropped last node contributes no children
            occ[:n_code] = torch.from_numpy(sym.astype(np.int64) + 1).to(dev)
            # children in BFS order: parents in order, child digit ascending (bit d of the byte <-> digit d, Octree.py:175)
            child = ((occ[:, None] >> torch.arange(8, device=dev)[None, :]) & 1).nonzero()
            par, dig = child[:, 0], child[:, 1]
            cell = 1 << (n - L)                                              # cell size one level down
            pos = pos[par] + bits[dig] * cell
            if L < n:
                ctx[:, 3, 2] = (occ - 1).clamp(min=0).to(torch.uint8)
                anc = ctx[par][:, 1:4].contiguous()
                octant = (dig + 1).to(torch.uint8)
        return np.concatenate(occ_all), pos.cpu().numpy()

    @torch.no_grad()
    def decode(self, fr) -> DecodedFrame:
        """fr: FrameResult (bitstream, depths, pos_mm)."""
        dec = coder.RangeDecoder(fr.bitstream)
        out = DecodedFrame(depths=list(fr.depths))
        lv = 0
        for depth in fr.depths:
            occ, vox = self._decode_tree(dec, depth, fr.pos_mm[lv: lv + depth], drop_last=self.mullevel,
                                         pos_eps_last=not self.mullevel)
            out.occ.append(occ)
            out.voxels.append(vox)
            lv += depth
        out.n_symbols = dec.count
        return out


def dequantise(voxels, steps, offset, mode):
    """decode_ehem.py:226-243: quantised coordinates -> points (float64): ``v * steps + offset`` then spher2cart /
    cylin2cart (data_preprocess.py:179-229).  ``voxels`` int64 [V,3] (host or device); returns a float64 array [V,3]."""
    v = torch.as_tensor(voxels).to(torch.float64)
    p = v * torch.as_tensor(np.asarray(steps, np.float64), device=v.device) + torch.as_tensor(np.asarray(offset, np.float64), device=v.device)
    if mode == "spher":
        rho, phi, th = p[:, 0], p[:, 1], p[:, 2]
        p = torch.stack([rho * torch.sin(th) * torch.cos(phi), rho * torch.sin(th) * torch.sin(phi), rho * torch.cos(th)], 1)
    elif mode == "cylin":
        rho, phi, z = p[:, 0], p[:, 1], p[:, 2]
        p = torch.stack([rho * torch.cos(phi), rho * torch.sin(phi), z], 1)
    return p.cpu().numpy()
