"""CPU ORACLE (test infrastructure, NOT a product path) for the integer stages of SCP's encoder.

Vectorised numpy restatement of what the reference computes; every function cites the
reference lines it follows (paths under /root/reference).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this module; ``scp_b200`` never does.

Parity pin: ``oracle/make_golden.py`` ran the *unmodified reference* in the build container
(numpy 2.3.5) and stored its outputs under ``tests/golden``; ``tests/test_oracle_vs_golden.py``
checks every function here bit-exactly against those fixtures.
"""
import math

import numpy as np

PAD_OCC = 256      # Octree.py:105  missing ancestor occupancy before the "-1" of the datasets


# --------------------------------------------------------------------------------------
# A1  coordinate transform + quantisation
# --------------------------------------------------------------------------------------
def cart2spher(p):
    """data_preprocess.py:200-207 (float32 in, float32 out under numpy>=2)."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    rho = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    phi = np.arctan2(y, x + 1e-9)
    phi[np.where(phi < 0)[0]] += 2 * math.pi
    theta = np.arccos(z / rho)
    return np.vstack((rho, phi, theta)).transpose(1, 0)


def cart2cylin(p):
    """data_preprocess.py:171-177."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    rho = np.sqrt(x ** 2 + y ** 2)
    phi = np.arctan2(y, x + 1e-9)
    phi[np.where(phi < 0)[0]] += 2 * math.pi
    return np.vstack((rho, phi, z)).transpose(1, 0)


def quantize(xyz, qs0, mode="spher", cart_offset=0.0):
    """data_preprocess.py:41-56,68 (proc_pc) == :111-125,137 (mul_proc_pc).

    xyz: (n,3) float32.  mode 'spher' | 'cylin' | 'cart'.
    Returns dict(q=(n,3) int64 *before* np.unique, bin_num, steps (3,) f64, offset (3,) f64)."""
    ref_pt = np.ascontiguousarray(np.asarray(xyz)[:, :3], dtype=np.float32)
    if mode == "cylin":
        points = cart2cylin(ref_pt)
        bin_num = np.round(points[:, 0].max() / qs0) + 1
        steps = np.array([qs0, 2 * math.pi / (bin_num - 1), qs0])[True]
        offset = np.array([0.0, 0.0, min(points[:, 2])])[True]
    elif mode == "spher":
        points = cart2spher(ref_pt)
        bin_num = np.round(points[:, 0].max() / qs0) + 1
        steps = np.array([qs0, 2 * math.pi / (bin_num - 1), math.pi / (bin_num - 1)])[True]
        offset = np.zeros((1, 3))
        points = points - 0
    else:
        points = ref_pt
        bin_num = np.float32(0)
        steps = np.array([qs0, qs0, qs0], dtype=np.float64)[True]
        offset = np.array([cart_offset] * 3, dtype=np.float64)[True]
    if mode != "spher":
        points = points - offset
    pt = np.round(points / steps)
    return dict(q=pt.astype(np.int64), bin_num=float(bin_num), steps=steps[0].astype(np.float64),
                offset=np.asarray(offset, np.float64).reshape(3))


# --------------------------------------------------------------------------------------
# A2/A3/A4  Morton octree + K=4 ancestor rows
# --------------------------------------------------------------------------------------
def depth_of(q):
    """Octree.py:58  n = ceil(log2(max(A)+1)) over all three axes."""
    return int(np.ceil(np.log2(np.max(q) + 1)))


def morton_keys(q, n):
    """Octree.py:56-65: bit b of (x,y,z) interleaved, x most significant."""
    q = q.astype(np.uint64)
    key = np.zeros(len(q), np.uint64)
    for b in range(n):
        for c in range(3):
            key |= ((q[:, c] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + (2 - c))
    return key


def tree_rows(q, morton_path=None, drop_last=False, K=4):
    """Octree.py:148-181 (GenOctree) + :102-137 (gen_K_parent_seq), or with ``morton_path``
    Octree.py:184-221 + :224-272 (mullevel variant: filter on the rho bits, drop the last row).

    q: (V,3) non-negative ints (duplicates allowed).  Returns dict with
      rows  (N,K,6) int64  [occ 1..256, level, octant, x, y, z]
      level_counts (n,)    nodes per level
      codes (N,) occupancy bytes in BFS order, depth n
    """
    q = np.asarray(q, np.int64)
    n = depth_of(q)
    if morton_path is not None:
        keep = np.ones(len(q), bool)
        for j, bit in enumerate(morton_path):      # Octree.py:188 (x bits are every third Morton bit)
            keep &= ((q[:, 0] >> (n - 1 - j)) & 1) == bit
        q = q[keep]
    keys = np.unique(morton_keys(q, n))
    V = len(keys)
    occ_l, oct_l, par_l, pos_l, lvl_l = [], [], [], [], []
    prev_prefix = None
    for L in range(1, n + 1):
        sh = np.uint64(3 * (n - L + 1))
        prefix = keys >> sh if 3 * (n - L + 1) < 64 else np.zeros(V, np.uint64)
        digit = (keys >> np.uint64(3 * (n - L))) & np.uint64(7)
        nodes, inv = np.unique(prefix, return_inverse=True)
        occ = np.zeros(len(nodes), np.int64)
        np.bitwise_or.at(occ, inv, (1 << digit.astype(np.int64)))
        octant = (nodes & np.uint64(7)).astype(np.int64) + 1 if L > 1 else np.ones(1, np.int64)
        if L > 1:
            parent = np.searchsorted(prev_prefix, nodes >> np.uint64(3))
        else:
            parent = np.zeros(1, np.int64)
        # Octree.py:140-145 get_pos: origin of the node's own cell in full-resolution units
        pos = np.zeros((len(nodes), 3), np.int64)
        for j in range(L - 1):                      # digit j of the (L-1)-digit prefix
            dj = (nodes >> np.uint64(3 * (L - 2 - j))) & np.uint64(7)
            for c in range(3):
                pos[:, c] |= (((dj >> np.uint64(2 - c)) & np.uint64(1)).astype(np.int64)) << (n - 1 - j)
        occ_l.append(occ); oct_l.append(octant); par_l.append(parent); pos_l.append(pos)
        lvl_l.append(np.full(len(nodes), L, np.int64))
        prev_prefix = nodes
    counts = np.array([len(o) for o in occ_l], np.int64)
    starts = np.concatenate([[0], np.cumsum(counts)])
    N = int(starts[-1])
    occ = np.concatenate(occ_l); octant = np.concatenate(oct_l); level = np.concatenate(lvl_l)
    pos = np.concatenate(pos_l)
    parent_row = np.concatenate([par_l[i] + (starts[i - 1] if i > 0 else 0) for i in range(n)])
    rows = np.zeros((N, K, 6), np.int64)
    rows[:, :, 0] = PAD_OCC
    rows[:, K - 1, 0] = occ; rows[:, K - 1, 1] = level; rows[:, K - 1, 2] = octant
    rows[:, K - 1, 3:] = pos
    anc = np.arange(N)
    valid = np.ones(N, bool)
    for k in range(K - 2, -1, -1):                 # parent, grandparent, great-grandparent
        valid = valid & (level[anc] > 1)
        anc = np.where(valid, parent_row[anc], 0)
        rows[valid, k, 0] = occ[anc[valid]]
        rows[valid, k, 1] = level[anc[valid]]
        rows[valid, k, 2] = octant[anc[valid]]
        rows[valid, k, 3:] = pos[anc[valid]]
    if drop_last:                                   # Octree.py:259-262  Seq[1:n]
        rows = rows[:-1]
    return dict(rows=rows, level_counts=counts, codes=occ, depth=n, n_voxels=V)


def voxels_unique(q):
    """data_preprocess.py:69 np.unique(axis=0): lexicographically sorted unique voxels."""
    return np.unique(np.asarray(q, np.int64), axis=0)


# --------------------------------------------------------------------------------------
# A5  EHEM dataset level split
# --------------------------------------------------------------------------------------
def ehem_level_split(rows, lidar_level, mullevel=False):
    """encode_dataset_ehem.py:52-105 (spher/cylin branch) / encode_dataset_ehem_mullevel.py:47-85.

    Returns (ids, poss, pos_mm, data, oct_seq) lists per level like the reference dataset."""
    oct_seq = np.array(rows, np.int64, copy=True)
    oct_seq[:, :, 0] -= 1
    lv = oct_seq[:, -1, 1]
    cuts = np.flatnonzero(lv[1:] > np.maximum.accumulate(lv)[:-1]) + 1
    bounds = np.concatenate([[0], cuts, [len(oct_seq)]])
    ids, poss, pos_mm, data = [], [], [], []
    nlev = len(bounds) - 1
    for i in range(nlev):
        s, e = bounds[i], bounds[i + 1]
        blk = oct_seq[s:e, :, :3]
        last = i == nlev - 1
        if last:
            blk[:, :, 1] = np.minimum(blk[:, :, 1], lidar_level)   # in place, like :86
        data.append(np.concatenate((blk[:, :, 1:], blk[:, :, :1]), axis=2))
        cur = oct_seq[s:e, -1, 3:6]
        mx, mn = cur.max(), cur.min()
        eps = 0.0 if (last and mullevel) else 1e-9
        with np.errstate(divide="ignore", invalid="ignore"):
            poss.append(((cur - mn) / (mx - mn + eps)).astype(np.float32).transpose((1, 0)))
        pos_mm.append((mn, mx))
        ids.append(np.arange(e - s, dtype=np.int64))
    return ids, poss, pos_mm, data, oct_seq


# --------------------------------------------------------------------------------------
# A6  OctAttention dataset
# --------------------------------------------------------------------------------------
def octattn_dataset(rows, context_size=1024):
    """encode_dataset.py:32-55 with level_wise=False as ``encode.py compress`` is run for
    OctAttention (cur_level starts at 100, so there is a single block)."""
    oct_seq = np.array(rows, np.int64, copy=True)
    pad = np.zeros((context_size - 1, oct_seq.shape[1], oct_seq.shape[2]), np.int64)
    pad[:, :, 0] = 255
    oct_seq[:, :, 0] -= 1
    max_level = oct_seq[:, -1, 1].max()
    data = np.vstack((pad[:, :, :3], oct_seq[:, :, :3]))
    pos = np.vstack((pad[:, :, 3:].astype(np.float32),
                     (oct_seq[:, :, 3:] / (2 ** max_level)).astype(np.float32)))
    ids = np.hstack((-np.ones(context_size - 1, np.int64), np.arange(len(oct_seq), dtype=np.int64)))
    return ids, pos, data, oct_seq


# --------------------------------------------------------------------------------------
# A7  coding order
# --------------------------------------------------------------------------------------
def coding_order(level_sizes, context_size=8192, mullevel=False):
    """encode.py:109-136 / encode_mullevel.py:106-133: even ids then odd ids per window."""
    out, base = [], 0
    for n in level_sizes:
        for i in range(0, n, context_size):
            ids = np.arange(i, min(i + context_size, n), dtype=np.int64)
            if n == 1:
                # encode.py:123 does NOT add coded_cnt, encode_mullevel.py:120 does
                out.append(ids[-1:] + (base if mullevel else 0))
                continue
            out.append(ids[::2] + base)
            out.append(ids[1::2] + base)
        base += n
    return np.concatenate(out) if out else np.zeros(0, np.int64)


# --------------------------------------------------------------------------------------
# A13  PMF -> integer CDF
# --------------------------------------------------------------------------------------
def pmf_to_cdf_u16(pmf):
    """numpyAc.py:109-114 + :80-107, then reinterpreted as uint16 by the coder
    (numpyAc_backend.cpp:233-234).  pmf (N,255) float32 -> (N,256) uint16."""
    pmf = np.asarray(pmf, np.float32)
    c = np.cumsum(pmf, axis=1)                 # float32, sequential
    c = c / c[:, -1:]
    F = np.hstack((np.zeros((pmf.shape[0], 1)), c))   # float64
    F = np.round(F * (2 ** 16 - (F.shape[1] - 1)))
    cdf = F.astype(np.int16)
    cdf += np.arange(F.shape[1]).astype(np.int16)
    return cdf.view(np.uint16)


def softmax_f32(logits):
    x = np.asarray(logits, np.float32)
    x = x - x.max(-1, keepdims=True)
    e = np.exp(x)
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)
