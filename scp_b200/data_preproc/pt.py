"""Point-cloud readers with the semantics of the reference's data_preproc/pt.py:162-281 that the encode path
uses (KITTI ``.bin`` rows x,y,z,intensity float32; ASCII ``.ply`` vertex lists).  Host-side I/O only."""
import os

import numpy as np


def loadbin(file):
    """pt.py:190-192"""
    points = np.fromfile(file, dtype=np.float32).reshape(-1, 4)
    return points[:, 0:3], points[:, 3:4]


def loadply(path, color_format="rgb"):
    """ASCII PLY with x y z as the first three vertex properties (pt.py:224-281 semantics: coords only)."""
    with open(path, "rb") as f:
        n, header_done = 0, False
        while not header_done:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line.startswith("format") and "ascii" not in line:
                raise ValueError("only ASCII .ply is supported (like the reference's loadply)")
            header_done = line == "end_header"
        data = np.loadtxt(f, dtype=np.float32, max_rows=n, ndmin=2)
    return data[:, :3], data[:, 3:]


def pcread(path, color_format="rgb"):
    if not os.path.exists(path):
        raise Exception("no such file:" + path)       # same error type/message as pt.py:176-177
    if path.endswith(".ply"):
        return loadply(path, color_format)
    if path.endswith(".bin"):
        return loadbin(path)
    raise ValueError("unsupported point cloud file: " + path)


def ptread(path):
    """pt.py:162-168"""
    return pcread(path, "geometry")[0]


def write_ply_data(filename, points, attributeName=[], attriType=[]):
    """pt.py:114-151: ASCII PLY, ``%f`` coordinates, optional integer / float attribute columns."""
    points = np.asarray(points)
    assert points.shape[1] >= len(attributeName) + 3
    filename = str(filename)
    d = os.path.dirname(filename)
    if d != "" and not os.path.exists(d):
        os.makedirs(d)
    head = ["ply", "format ascii 1.0", "element vertex " + str(points.shape[0]),
            "property float x", "property float y", "property float z"]
    head += ["property " + t + " " + n for n, t in zip(attributeName, attriType)]
    kinds = {"uint16": "%d", "float": "%f", "uchar": "%d"}
    fmt = " ".join(["%f", "%f", "%f"] + [kinds[t] for t in attriType])
    with open(filename, "w") as f:
        f.write("\n".join(head + ["end_header"]) + "\n")
        rows = points[:, :3 + len(attriType)]
        for a in range(0, len(rows), 65536):              # one C-level format call per block instead of one per row
            blk = rows[a:a + 65536]
            f.write((fmt + "\n") * len(blk) % tuple(blk.ravel().tolist()))


def distChamfer(f1, f2, scale=1.0):
    """pt.py:88-95 on the GPU (exact brute-force nearest neighbours, ``scp_b200.metrics``)."""
    from .. import metrics
    return metrics.distChamfer(f1, f2, scale)
