"""Drop-in for the encode-path functions of data_preproc/Octree.py: ``gen_K_parent_seq`` (:102-137) over the
octree object returned by ``gen_octree``.  Vectorised on the host over the node records the CUDA pipeline produced
(the reference loops over every node in Python with ctypes attribute reads)."""
import numpy as np


def gen_K_parent_seq(octree, K):
    LevelNum = len(octree)
    recs, lvls = [], []
    for L in range(LevelNum):
        lv = octree[L]
        n = len(lv)
        a = np.zeros((n, 7), np.int64)
        for i in range(n):
            nd = lv[i]
            a[i] = (nd.nodeid, nd.parent, nd.oct, nd.octant, nd.pos[0], nd.pos[1], nd.pos[2])
        recs.append(a)
        lvls.append(np.full(n, L + 1, np.int64))
    rec = np.concatenate(recs)
    level = np.concatenate(lvls)
    N = len(rec)
    Seq = np.ones((N, K), "int") * 256
    LevelOctant = np.zeros((N, K, 2), "int")
    Pos = np.zeros((N, K, 3), "int")
    anc = np.arange(N)
    valid = np.ones(N, bool)
    for k in range(K - 1, -1, -1):
        Seq[valid, k] = rec[anc[valid], 2]
        LevelOctant[valid, k, 0] = level[anc[valid]]
        LevelOctant[valid, k, 1] = rec[anc[valid], 3]
        Pos[valid, k] = rec[anc[valid], 4:7]
        valid = valid & (level[anc] > 1)
        anc = np.where(valid, rec[anc, 1] - 1, 0)       # parent nodeid is 1-based
    assert N == rec[-1, 0]
    return {"Seq": Seq, "Level": LevelOctant, "ChildID": [[] for _ in range(N)], "Pos": Pos}
