"""TEST INFRASTRUCTURE: the "explained rows" PMF parity criterion for SCP-EHEM (BASELINE.json: PMFs within 1e-3).

The three kNN graphs of the DGCNN front end (dgcnn.py:131-142) are the only discontinuous step of the model: when the k-th
and (k+1)-th candidate of a row are equally far (exactly, on gridded octree positions; or up to the float32 rounding noise
of the reference's ``2 x.y - |x|^2 - |y|^2`` on learned features), the reference's pick is an accident of ``torch.topk`` /
MKL summation order, and one swapped neighbour moves a handful of PMFs by a few 1e-3.  Instead of
tolerating "a few percent of rows", parity is split into three assertions, each strict:

 (A) ARITHMETIC: with the oracle (oracle/ehem_torch.py, pinned on the reference's golden logits) forced to use the
     neighbour sets the device chose, every PMF row agrees within 1e-3.  No allowance.
 (B) NEIGHBOURS: every neighbour set the device chose is a correct k-nearest set of the oracle's own float32 features up
     to ties: wherever it differs from the oracle's ``topk`` pick, the swapped candidates are equally far from the query
     within TIE_EPS (relative to |x|^2 + |y|^2, the magnitude whose rounding noise decides the reference's order).
     Stage s is judged on features computed with the device's sets for the stages before it, so one early swap does
     not excuse later ones.
 (C) DIRECT: when no row of the window has a near-tie at its k/k+1 boundary, the reference's own logits (golden files)
     must be matched within 1e-3 on every row; otherwise the deviation from the raw reference run is reported and must
     stay below the bound a swapped neighbour can cause (5e-2).
"""
import torch

from oracle import ehem_torch as O

PMF_TOL = 1e-3
TIE_EPS = 1e-6          # float32 noise of the score 2x.y-|x|^2-|y|^2 over 144-192 channels and of the features themselves,
                        # relative to |x|^2 + |y|^2.  Measured on B200 over all test windows: the largest gap at a row where the
                        # device's set differs from the oracle's is 3.2e-8 (every test prints "max_gap_at_differing_rows")


class RecordKnn:
    """Wraps ``ops.knn`` of a model to keep the index tables of the calls of one forward pass (stage order)."""

    def __init__(self, ops):
        self.ops, self.idx, self._orig = ops, [], ops.knn

    def __enter__(self):
        def knn(x, seqs, k):
            out = self._orig(x, seqs, k)
            self.idx.append(out.detach().cpu().long())
            return out
        self.ops.knn = knn
        return self

    def __exit__(self, *a):
        self.ops.knn = self._orig


def _sqdist64(x, i, j):
    """Exact squared distances |x_i - x_j|^2 in float64; x [N,C] float32, i [R], j [R,m]."""
    xd = x.double()
    return ((xd[i][:, None, :] - xd[j]) ** 2).sum(-1)


class ForcedKnn:
    """``knn`` hook of the oracle: returns the device's neighbour sets and records how they relate to the oracle's own."""

    def __init__(self, device_idx):
        self.device_idx, self.stage, self.report = device_idx, 0, []

    def __call__(self, x, k):                       # x [C,N] float32
        own = O.knn_torch_topk(x, k)
        N = x.shape[1]
        forced = self.device_idx[self.stage][:N, :k]
        xs = x.t().contiguous()
        xx = (xs.double() ** 2).sum(1)
        rows = torch.arange(N)
        d_own, d_dev = _sqdist64(xs, rows, own), _sqdist64(xs, rows, forced)
        scale = xx[:, None] + torch.maximum(xx[own], xx[forced]).max(1, keepdim=True)[0] + 1e-30
        # a device set is a valid kNN set iff its farthest member is no farther than the oracle's farthest (k-th
        # distance) beyond a tie, and it has k distinct members
        distinct = torch.tensor([len(set(r.tolist())) for r in forced]) == min(k, N)
        excess = ((d_dev.max(1)[0] - d_own.max(1)[0]).clamp(min=0) / scale[:, 0])
        differs = torch.tensor([set(a.tolist()) != set(b.tolist()) for a, b in zip(own, forced)])
        # near-tie rows of the window: gap between the k-th and the (k+1)-th distance below TIE_EPS
        if N > k:
            xd = xs.double()                                  # all pairs: |x|^2 + |y|^2 - 2 x.y in float64 (1e-16 relative)
            d_all = (xx[:, None] + xx[None, :] - 2.0 * (xd @ xd.t())).clamp_(min=0)
            srt = d_all.topk(k + 1, dim=1, largest=False)[0]
            gap = (srt[:, k] - srt[:, k - 1]) / scale[:, 0]
            near = gap < TIE_EPS
        else:
            gap = torch.full((N,), float("inf"), dtype=torch.float64)
            near = torch.zeros(N, dtype=torch.bool)
        self.report.append({"stage": self.stage, "rows": N, "differ": int(differs.sum()), "near_tie_rows": int(near.sum()),
                            "max_excess": float(excess.max()) if N else 0.0, "all_distinct": bool(distinct.all()),
                            "differ_outside_near_ties": int((differs & ~near).sum()),
                            "max_gap_at_differing_rows": float(gap[differs].max()) if bool(differs.any()) else 0.0})
        self.stage += 1
        return forced


def pmf_rows_err(a, b):
    return (torch.softmax(torch.as_tensor(a).float().cpu(), -1) - torch.softmax(torch.as_tensor(b).float().cpu(), -1)).abs().amax(-1)


def check_explained_parity(model, sd, data, pos, ref1=None, ref2=None, ref_slice=slice(None), what=""):
    """data int64 [csz,4,3], pos float32 [3,csz] (CPU tensors of one window); ``model`` on the device under test.
    ref1/ref2: golden logits of the unmodified reference (optionally strided by ``ref_slice``).  Returns the report."""
    dev = next(model.parameters()).device
    with RecordKnn(model.ops) as rec:
        l1, l2 = model(data[None].to(dev), pos[None].to(dev))
    l1, l2 = l1[0].cpu(), l2[0].cpu()
    forced = ForcedKnn(rec.idx)
    o1, o2 = O.ehem_forward(sd, data, pos, knn=forced)
    e = torch.cat((pmf_rows_err(l1, o1), pmf_rows_err(l2, o2)))
    rep = {"what": what, "arith_max": float(e.max()), "arith_median": float(e.median()), "stages": forced.report}
    # (A) arithmetic parity, every row
    assert rep["arith_max"] <= PMF_TOL, rep
    # (B) the device's neighbour sets are k-nearest sets up to ties
    for s in forced.report:
        assert s["all_distinct"] and s["max_excess"] <= TIE_EPS and s["differ_outside_near_ties"] == 0, rep
    # (C) direct comparison with the reference's own run
    if ref1 is not None:
        d = torch.cat((pmf_rows_err(l1[ref_slice], ref1), pmf_rows_err(l2[ref_slice], ref2) if len(ref2) else torch.zeros(0)))
        rep["direct_max"] = float(d.max())
        rep["direct_rows_above_tol"] = int((d > PMF_TOL).sum())
        tie_free = all(s["near_tie_rows"] == 0 for s in forced.report)
        rep["window_tie_free"] = tie_free
        assert rep["direct_max"] <= (PMF_TOL if tie_free else 5e-2), rep
    print(rep)
    return rep
