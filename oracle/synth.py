"""Seeded KITTI-shape synthetic LiDAR frames.

TEST INFRASTRUCTURE (see oracle/__init__.py).  A 64-beam spinning LiDAR is
ray-cast against a ground plane, two side walls and 12 car-sized boxes, in the
camera frame the reference feeds its graph generator with
(x right, y down, z forward; reference dataset/kitti_dataset.py:998-1006).
Frame ``i`` uses ``numpy.random.default_rng(1000 + i)``.
"""
import numpy as np

SENSOR_HEIGHT = 1.73
MAX_RANGE = 80.0


def _ray_box(d, lo, hi):
    """Slab test for rays from the origin with directions d [R,3] -> t [R] (inf = miss)."""
    with np.errstate(divide='ignore', invalid='ignore'):
        t0 = lo[None, :] / d
        t1 = hi[None, :] / d
    tmin = np.minimum(t0, t1)
    tmax = np.maximum(t0, t1)
    tn = np.nanmax(tmin, axis=1)
    tf = np.nanmin(tmax, axis=1)
    hit = (tn <= tf) & (tf > 0) & (tn > 0)
    return np.where(hit, tn, np.inf)


def lidar_frame(frame_idx=0, num_points=20000, full_360=False):
    """-> (xyz [N,3] float32, intensity [N,1] float32).

    num_points=None keeps every return (about 28 k front crop / 127 k full 360).
    """
    rng = np.random.default_rng(1000 + int(frame_idx))
    elev = np.deg2rad(np.linspace(2.0, -24.8, 64))
    azim = np.deg2rad(np.arange(2000) * (360.0 / 2000) - 180.0)
    if not full_360:
        azim = azim[np.abs(azim) < np.deg2rad(40.5)]
    az, el = np.meshgrid(azim, elev, indexing='ij')
    az = az.ravel() + rng.normal(0.0, 1e-3, az.size)
    el = el.ravel() + rng.normal(0.0, 1e-3, el.size)
    d = np.stack([np.cos(el) * np.sin(az), -np.sin(el), np.cos(el) * np.cos(az)], axis=1)
    # ground plane y = +SENSOR_HEIGHT (y points down)
    with np.errstate(divide='ignore'):
        t = np.where(d[:, 1] > 1e-9, SENSOR_HEIGHT / d[:, 1], np.inf)
    # side walls x = +-12 m, 4 m tall
    for xw in (-12.0, 12.0):
        with np.errstate(divide='ignore'):
            tw = np.where(d[:, 0] * xw > 1e-9, xw / d[:, 0], np.inf)
        yw = tw * d[:, 1]
        ok = np.isfinite(tw) & (yw <= SENSOR_HEIGHT) & (yw >= SENSOR_HEIGHT - 4.0)
        t = np.minimum(t, np.where(ok, tw, np.inf))
    # 12 axis-aligned boxes, 1.6 wide (x) x 1.5 tall (y) x 4.0 long (z), on the ground
    bx = rng.uniform(-9.0, 9.0, 12)
    bz = rng.uniform(6.0, 60.0, 12)
    if full_360:
        bz = bz * rng.choice([-1.0, 1.0], 12)
    for cx, cz in zip(bx, bz):
        lo = np.array([cx - 0.8, SENSOR_HEIGHT - 1.5, cz - 2.0])
        hi = np.array([cx + 0.8, SENSOR_HEIGHT, cz + 2.0])
        t = np.minimum(t, _ray_box(d, lo, hi))
    keep = np.isfinite(t) & (t < MAX_RANGE)
    xyz = d[keep] * t[keep, None] + rng.normal(0.0, 0.02, (int(keep.sum()), 3))
    if num_points is not None:
        if xyz.shape[0] < num_points:
            raise ValueError('scene produced only %d returns' % xyz.shape[0])
        sel = rng.choice(xyz.shape[0], size=num_points, replace=False)
        xyz = xyz[np.sort(sel)]
    xyz = xyz.astype(np.float32)
    intensity = rng.random((xyz.shape[0], 1), dtype=np.float32)
    return xyz, intensity
