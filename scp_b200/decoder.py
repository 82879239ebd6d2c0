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
The level inputs (scp_decode_level_inputs), the child expansion (scp_expand_children), the entropy model, the CDF table
and the range decoder are the library's; torch only owns the device buffers.
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
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
        self.pool = ThreadPoolExecutor(min(16, os.cpu_count() or 1))

    # ------------------------------------------------------------------------------------------------------------
    def _windows(self, n):
        starts = np.arange(0, n, self.context, dtype=np.int64)
        lens = np.minimum(self.context, n - starts).astype(np.int32)
        toks = np.concatenate([[0], np.cumsum(lens + (lens & 1))]).astype(np.int64)
        return starts, lens, toks

    def _decode_level_batch(self, items):
        """items: list of (decoder, ctx uint8 [N,4,3] (self occupancy column = 255), pos_norm float32 [N,3], n_code), one per
        frame.  Decodes, for every frame, its first n_code nodes (the windows of encode.py:112-115) and returns the list of
        occupancy symbol arrays (host int16 [n_code]).  Phase 1 runs over the windows of ALL frames at once; the per-window
        part (even symbols -> phase 2 -> odd symbols, the order the stream was written in) runs in lock-step over the
        frames: window w of every frame that has one forms one phase-2 batch, and the host decoders of the frames run
        in parallel threads (the C call releases the GIL)."""
        dev, st = self.dev, _lib.stream_ptr()
        outs = [np.empty(n, np.int16) for (_, _, _, n) in items]
        # global window table: (frame, start row inside the frame's level, length)
        wins = []
        row0 = [0]
        for f, (_, ctx, _, n) in enumerate(items):
            starts, lens, _ = self._windows(n)
            wins += [(f, int(s0), int(l)) for s0, l in zip(starts, lens)]
            row0.append(row0[-1] + ctx.shape[0])
        if not wins:
            return outs
        ctx_all = torch.cat([it[1] for it in items]).contiguous()
        pos_all = torch.cat([it[2] for it in items]).contiguous()
        lens = np.asarray([w[2] for w in wins], np.int32)
        rows = np.asarray([row0[w[0]] + w[1] for w in wins], np.int64)
        toks = np.concatenate([[0], np.cumsum(lens + (lens & 1))]).astype(np.int64)
        # phase 1, chunks of windows bounded by max_tokens (like the encoder)
        where = [None] * len(wins)                                 # window -> (chunk, token offset inside the chunk)
        chunks = []
        w0 = 0
        while w0 < len(wins):
            w1 = w0 + 1
            while w1 < len(wins) and toks[w1 + 1] - toks[w0] <= self.max_tokens:
                w1 += 1
            T = int(toks[w1] - toks[w0])
            ctxp = torch.empty((T, 4, 3), dtype=torch.uint8, device=dev)
            posp = torch.empty((T, 3), dtype=torch.float32, device=dev)
            re = torch.empty(T // 2, dtype=torch.int64, device=dev)
            ro = torch.empty(T // 2, dtype=torch.int64, device=dev)
            tk = np.ascontiguousarray(toks[w0:w1] - toks[w0])
            _lib.check(self.lib.scp_gather_windows(_lib.ptr(ctx_all), _lib.ptr(pos_all), _lib.ptr(np.ascontiguousarray(rows[w0:w1])),
                                                   _lib.ptr(np.ascontiguousarray(lens[w0:w1])), _lib.ptr(tk), w1 - w0,
                                                   _lib.ptr(ctxp), _lib.ptr(posp), _lib.ptr(re), _lib.ptr(ro), st),
                       "scp_gather_windows")
            feat_a, l1 = self.model.phase1(ctxp, posp, [int(x) for x in np.append(tk, T)])
            cdf1 = coder.pmf_to_cdf(l1, is_logits=True, want_cdf=True)["cdf"].cpu().numpy()
            for w in range(w0, w1):
                where[w] = (len(chunks), int(tk[w - w0]))
            chunks.append((ctxp, feat_a, cdf1))
            w0 = w1
        # per-frame window lists, then lock-step over the window index
        per_frame = [[] for _ in items]
        for w, (f, _, _) in enumerate(wins):
            per_frame[f].append(w)
        for k in range(max(len(p) for p in per_frame)):
            group = [p[k] for p in per_frame if k < len(p)]

            def even(w):
                f, s0, ln = wins[w]
                c, t0 = where[w]
                ne = (ln + 1) // 2
                sym = items[f][0].decode(chunks[c][2][t0 // 2: t0 // 2 + ne])
                outs[f][s0: s0 + ln: 2] = sym
                return sym
            syms_e = list(self.pool.map(even, group)) if len(group) > 1 else [even(group[0])]
            g2 = [(w, se) for w, se in zip(group, syms_e) if wins[w][2] > 1]
            if not g2:
                continue
            cws, fas, offs = [], [], [0]
            for w, se in g2:
                f, s0, ln = wins[w]
                c, t0 = where[w]
                Tw = ln + (ln & 1)
                cw = chunks[c][0][t0: t0 + Tw]
                cw[0: 2 * len(se): 2, 3, 2] = torch.from_numpy(se.astype(np.uint8)).to(dev)
                cws.append(cw)
                fas.append(chunks[c][1][t0: t0 + Tw])
                offs.append(offs[-1] + Tw)
            if len(g2) == 1:
                cw_all, fa_all = cws[0], fas[0]
            else:
                cw_all, fa_all = torch.cat(cws), torch.cat(fas)
            l2 = self.model.phase2(cw_all, fa_all, offs)
            # rows of logits2 that belong to real odd nodes (the pad token of an odd-length window is dropped)
            keep = np.concatenate([np.arange(o // 2, o // 2 + wins[w][2] // 2) for (w, _), o in zip(g2, offs[:-1])])
            l2k = l2 if len(keep) == l2.shape[0] else l2[torch.from_numpy(keep).to(dev)]
            cdf2 = coder.pmf_to_cdf(l2k.contiguous(), is_logits=True, want_cdf=True)["cdf"].cpu().numpy()
            cuts = np.concatenate([[0], np.cumsum([wins[w][2] // 2 for w, _ in g2])])

            def odd(i):
                w = g2[i][0]
                f, s0, ln = wins[w]
                outs[f][s0 + 1: s0 + ln: 2] = items[f][0].decode(cdf2[cuts[i]: cuts[i + 1]])
            if len(g2) > 1:
                list(self.pool.map(odd, range(len(g2))))
            else:
                odd(0)
        return outs

    def _tree_level_inputs(self, st_, L, n, pos_mm, pos_eps_last):
        """Model inputs of level L from the per-tree state (pos int32 [N,3], anc uint8 [N,3,3], octant uint8 [N]):
        (ctx with unclipped levels, ctx for the model, pos_norm) -- scp_decode_level_inputs."""
        dev = self.dev
        pos, anc, octant = st_
        N = pos.shape[0]
        ctx = torch.empty((N, 4, 3), dtype=torch.uint8, device=dev)
        ctx_model = torch.empty((N, 4, 3), dtype=torch.uint8, device=dev)
        pos_norm = torch.empty((N, 3), dtype=torch.float32, device=dev)
        mn, mx = pos_mm[L - 1]
        den = float(mx - mn) + (0.0 if (L == n and not pos_eps_last) else 1e-9)      # encode_dataset_ehem.py:70-72 / mullevel :80
        clip = self.level if (L == n and n > self.level) else 255                    # encode_dataset_ehem.py:86
        _lib.check(self.lib.scp_decode_level_inputs(_lib.ptr(pos), _lib.ptr(anc), _lib.ptr(octant), N, L, clip, float(mn), den,
                                                    _lib.ptr(ctx), _lib.ptr(ctx_model), _lib.ptr(pos_norm), _lib.stream_ptr()),
                   "scp_decode_level_inputs")
        return ctx, ctx_model, pos_norm

    _POPC = np.array([bin(i).count("1") for i in range(256)], np.int64)

    def _expand(self, st_, ctx, sym, n_code, L, n):
        """Next level's node state from the decoded symbols of level L (scp_expand_children).  A dropped last node
        (encode_mullevel, Octree.py:259-262) has no decoded occupancy and contributes no children."""
        dev = self.dev
        pos = st_[0]
        N = pos.shape[0]
        occ_h = np.zeros(N, np.uint8)
        occ_h[:n_code] = (sym + 1).astype(np.uint8)
        M = int(self._POPC[occ_h].sum())
        occ = torch.from_numpy(occ_h).to(dev)
        cpos = torch.empty((M, 3), dtype=torch.int32, device=dev)
        canc = torch.empty((M, 3, 3), dtype=torch.uint8, device=dev)
        coct = torch.empty((M,), dtype=torch.uint8, device=dev)
        if M == 0:                                              # (a tree whose only node on this level was dropped)
            return (cpos, canc, coct)
        _lib.check(self.lib.scp_expand_children(_lib.ptr(occ), _lib.ptr(pos), _lib.ptr(ctx), N, L, 1 << (n - L), _lib.ptr(cpos),
                                                _lib.ptr(canc), _lib.ptr(coct), _lib.stream_ptr()), "scp_expand_children")
        return (cpos, canc, coct)

    @torch.no_grad()
    def decode_batch(self, frs) -> List[DecodedFrame]:
        """frs: list of FrameResult (bitstream, depths, pos_mm).  Frames are decoded in lock-step (tree by tree, level by
        level, window by window), so that every model call covers all frames."""
        dev = self.dev
        decs = [coder.RangeDecoder(fr.bitstream) for fr in frs]
        outs = [DecodedFrame(depths=list(fr.depths)) for fr in frs]
        drop_last, pos_eps_last = self.mullevel, not self.mullevel
        lv0 = [0] * len(frs)
        for j in range(max(len(fr.depths) for fr in frs)):
            act = [f for f, fr in enumerate(frs) if j < len(fr.depths)]
            state, occs = {}, {f: [] for f in act}
            for f in act:
                # level 1: the root (decode_ehem.py:79-82: ancestors (0,0,255), self (level 1, octant 1))
                anc = torch.zeros((1, 3, 3), dtype=torch.uint8, device=dev)
                anc[:, :, 2] = 255
                state[f] = (torch.zeros((1, 3), dtype=torch.int32, device=dev), anc, torch.ones(1, dtype=torch.uint8, device=dev))
            for L in range(1, max(frs[f].depths[j] for f in act) + 1):
                cur = [f for f in act if L <= frs[f].depths[j]]
                items, ctxs = [], {}
                for f in cur:
                    n = frs[f].depths[j]
                    ctx, ctx_model, pos_norm = self._tree_level_inputs(state[f], L, n, frs[f].pos_mm[lv0[f]: lv0[f] + n], pos_eps_last)
                    N = ctx.shape[0]
                    n_code = N - 1 if (drop_last and L == n) else N
                    ctxs[f] = ctx
                    items.append((decs[f], ctx_model, pos_norm, n_code))
                syms = self._decode_level_batch(items)
                for f, sym, (_, _, _, n_code) in zip(cur, syms, items):
                    occs[f].append((sym + 1).astype(np.uint8))
                    state[f] = self._expand(state[f], ctxs[f], sym, n_code, L, frs[f].depths[j])
            for f in act:
                outs[f].occ.append(np.concatenate(occs[f]))
                outs[f].voxels.append(state[f][0].cpu().numpy().astype(np.int64))
                lv0[f] += frs[f].depths[j]
        for f in range(len(frs)):
            outs[f].n_symbols = decs[f].count
        return outs

    def decode(self, fr) -> DecodedFrame:
        """fr: FrameResult (bitstream, depths, pos_mm)."""
        return self.decode_batch([fr])[0]


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
