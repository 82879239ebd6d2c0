"""Seeded synthetic LiDAR sweeps (KITTI- and Ford-shaped) for parity tests and benchmarks.

No datasets are available offline, so every test/bench input comes from here (SURVEY.md
section 8d).  The generator is plain numpy on the host: it produces *inputs*, it is not part
of the accelerated path.

``guard_band`` removes the points whose quantised bin coordinate lies close to a rounding
boundary.  numpy's float32 ``arctan2``/``arccos`` (SVML) are not correctly rounded, the CUDA
path rounds correctly; on guard-banded inputs both agree on every quantised integer, which is
where bit-exactness is defined from (DESIGN.md "Float stage").
"""
import math

import numpy as np

KITTI_QS = lambda level: 400.0 / (2 ** level - 1)   # reference: encode_dataset_ehem.py:164
FORD_QS = lambda level: float(2 ** (18 - level))    # reference: test_gene.py:77


def _sweep(rng, n_beams, n_az, elev_hi_deg, elev_lo_deg, r_min, r_max, sensor_h=1.73):
    elev = np.deg2rad(np.linspace(elev_hi_deg, elev_lo_deg, n_beams))
    az = np.linspace(0.0, 2 * math.pi, n_az, endpoint=False)
    e, a = np.meshgrid(elev, az, indexing="ij")
    e = e + rng.normal(0.0, 2e-4, e.shape)
    a = a + rng.normal(0.0, 2e-4, a.shape)
    n_sectors = 90
    wall_d = rng.uniform(6.0, 80.0, n_sectors)
    wall_h = rng.uniform(1.0, 12.0, n_sectors)
    open_sector = rng.random(n_sectors) < 0.25
    sec = np.floor((a % (2 * math.pi)) / (2 * math.pi) * n_sectors).astype(int) % n_sectors
    se, ce = np.sin(e), np.cos(e)
    with np.errstate(divide="ignore", invalid="ignore"):
        r_ground = np.where(se < 0, sensor_h / (-se), np.inf)
        r_wall = wall_d[sec] / ce
    z_wall = r_wall * se
    hit_wall = (~open_sector[sec]) & (z_wall <= -sensor_h + wall_h[sec]) & (z_wall >= -sensor_h)
    r = np.where(hit_wall, np.minimum(r_wall, r_ground), r_ground)
    r = r + rng.normal(0.0, 0.02, r.shape)
    keep = np.isfinite(r) & (r > r_min) & (r < r_max) & (rng.random(r.shape) > 0.03)
    r, e, a = r[keep], e[keep], a[keep]
    x = r * np.cos(e) * np.cos(a)
    y = r * np.cos(e) * np.sin(a)
    z = r * np.sin(e)
    return np.stack([x, y, z], 1)


def kitti_sweep(seed=0, n_points=120000, r_max=105.0):
    """(n,4) float32 rows x,y,z,intensity in metres, like a KITTI velodyne .bin (pt.py:190-192)."""
    rng = np.random.default_rng(seed)
    xyz = _sweep(rng, 64, 2900, 2.0, -24.8, 2.5, r_max)
    if len(xyz) > n_points:
        xyz = xyz[np.sort(rng.choice(len(xyz), n_points, replace=False))]
    out = np.empty((len(xyz), 4), np.float32)
    out[:, :3] = xyz.astype(np.float32)
    out[:, 3] = rng.random(len(xyz)).astype(np.float32)
    return out


def ford_sweep(seed=0, n_points=80000, r_max=100.0):
    """(n,4) float32; coordinates are integer millimetres stored as float32 (Ford-shaped)."""
    rng = np.random.default_rng(seed + 100003)
    xyz = _sweep(rng, 32, 3600, 10.0, -30.0, 2.5, r_max, sensor_h=2.4)
    if len(xyz) > n_points:
        xyz = xyz[np.sort(rng.choice(len(xyz), n_points, replace=False))]
    out = np.empty((len(xyz), 4), np.float32)
    out[:, :3] = np.round(xyz * 1000.0).astype(np.float32)
    out[:, 3] = 0
    return out


def bin_coordinates_f64(xyz, qs0, mode):
    """float64 bin coordinates (before rint) of the reference's quantiser, mirroring its
    float32/float64 dtype chain (data_preprocess.py:42-56,68) with correctly rounded float32
    angles.  mode: 'spher' | 'cylin'.  Returns (coords f64 (n,3), bin_num float32)."""
    p = np.asarray(xyz, np.float32)[:, :3]
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    if mode == "spher":
        rho = np.sqrt(x * x + y * y + z * z, dtype=np.float32)
    else:
        rho = np.sqrt(x * x + y * y, dtype=np.float32)
    xe = (x + np.float32(1e-9)).astype(np.float32)
    phi = np.arctan2(y.astype(np.float64), xe.astype(np.float64)).astype(np.float32)
    phi = np.where(phi < 0, (phi + np.float32(2 * math.pi)).astype(np.float32), phi)
    if mode == "spher":
        third = np.arccos((z / rho).astype(np.float32).astype(np.float64)).astype(np.float32)
    else:
        third = z
    bin_num = np.float32(np.rint(rho.max() / np.float32(qs0))) + np.float32(1)
    s_phi = np.float64(np.float32(2 * math.pi) / (bin_num - np.float32(1)))
    if mode == "spher":
        s3 = np.float64(np.float32(math.pi) / (bin_num - np.float32(1)))
        off3 = 0.0
    else:
        s3 = float(qs0)
        off3 = np.float64(z.min())
    c = np.stack([rho.astype(np.float64) / float(qs0),
                  phi.astype(np.float64) / s_phi,
                  (third.astype(np.float64) - off3) / s3], 1)
    return c, bin_num


def guard_band(points, qs0, mode, margin=0.02):
    """Drops points within ``margin`` bins of a rounding boundary in any axis, iterating until the
    frame-level scalars (rho max -> bin_num, z min) are stable, and requires rho.max()/qs to stay
    away from .5 as well (it decides bin_num)."""
    pts = np.asarray(points, np.float32)
    for _ in range(8):
        c, _ = bin_coordinates_f64(pts, qs0, mode)
        frac = np.abs(c - np.floor(c) - 0.5)
        ok = (frac > margin).all(1)
        if ok.all():
            break
        pts = pts[ok]
    return pts


def make_frame(kind="kitti", seed=0, level=12, mode="spher", guard=True, n_points=None):
    """Returns (points (n,4) float32, qs0)."""
    if kind == "kitti":
        pts = kitti_sweep(seed, n_points or 120000)
        qs0 = KITTI_QS(level)
    elif kind == "ford":
        pts = ford_sweep(seed, n_points or 80000)
        qs0 = FORD_QS(level)
    else:
        raise ValueError(kind)
    if guard and mode in ("spher", "cylin"):
        pts = guard_band(pts, qs0, mode)
    return pts, qs0
