"""Deterministic synthetic scans of the shapes BASELINE.json names.

All generators return float64 arrays holding float32-representable values (a ROS
PointCloud2 delivers float32 xyz which the reference widens to double,
ufomap_ros/src/conversions.cpp:88-95).  Noise comes from a counter-based
splitmix64 stream so the same (seed, scan index) gives the same scan on every
machine and numpy version.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(idx, seed):
    """Vectorised splitmix64 of (seed + idx*golden); returns uint64."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform01(n, seed, stream=0):
    idx = np.arange(n, dtype=np.uint64) + np.uint64(stream) * np.uint64(1 << 40)
    return (_splitmix(idx, seed) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def _to_f32_grid(p):
    return p.astype(np.float32).astype(np.float64)


def _box_range(origin, dirs, lo, hi):
    """Distance from `origin` along unit `dirs` to the inside of the box [lo, hi]."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - origin) / dirs
        t2 = (hi - origin) / dirs
    t = np.where(dirs > 0, t2, np.where(dirs < 0, t1, np.inf))
    return t.min(axis=1)


def sensor_origin(k, step=(0.25, 0.10, 0.0), base=(0.013, 0.021, 0.007)):
    """Pose of scan k of a stream."""
    return np.array([base[0] + step[0] * k, base[1] + step[1] * k, base[2] + step[2] * k])


def velodyne64(k=0, rings=64, azimuths=2048, seed=88172645463325252, noise=0.02,
               room_lo=(-60.0, -40.0, -1.8), room_hi=(60.0, 40.0, 12.0), origin=None):
    """Velodyne-64-shaped scan inside a box room: `rings` x `azimuths` returns
    (131 072 by default), elevation -24.8..+2.0 deg, range = first wall hit +
    U[-noise, +noise].  Returns (origin[3], xyz[n,3]) in the map frame, ring-major."""
    o = sensor_origin(k) if origin is None else np.asarray(origin, float)
    elev = np.deg2rad(np.linspace(-24.8, 2.0, rings))
    az = (np.arange(azimuths) + 0.5) * (2 * np.pi / azimuths)
    e, a = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)], axis=-1).reshape(-1, 3)
    r = _box_range(o, d, np.asarray(room_lo), np.asarray(room_hi))
    n = len(d)
    r = r + (uniform01(n, seed, stream=k) * 2.0 - 1.0) * noise
    return o, _to_f32_grid(o + d * r[:, None])


def rgbd(k=0, width=640, height=480, seed=88172645463325252, noise=0.002,
         room_lo=(-3.5, -2.5, -1.2), room_hi=(3.5, 2.5, 1.5), origin=None):
    """RGB-D-shaped scan: pinhole camera looking along +x (fx=fy=525 at 640x480, scaled
    with the resolution).  Returns (origin[3], xyz[n,3], rgb[n,3])."""
    o = (np.array([0.0013 + 0.02 * k, 0.0021 + 0.01 * k, 0.0007])
         if origin is None else np.asarray(origin, float))
    f = 525.0 * width / 640.0
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    v, u = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    d = np.stack([np.ones(u.shape), -(u - cx) / f, -(v - cy) / f], axis=-1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = _box_range(o, d, np.asarray(room_lo), np.asarray(room_hi))
    n = len(d)
    r = r + (uniform01(n, seed, stream=1000 + k) * 2.0 - 1.0) * noise
    uu, vv = u.reshape(-1), v.reshape(-1)
    rgb = np.stack([uu & 255, vv & 255, (7 * uu + 13 * vv) & 255], axis=-1).astype(np.uint8)
    return o, _to_f32_grid(o + d * r[:, None]), rgb


def random_shell(n=10000, seed=42, rmin=2.0, rmax=6.0, origin=(0.013, 0.021, 0.007)):
    """n directions uniform on the sphere, ranges U[rmin, rmax] (config #1 shape)."""
    o = np.asarray(origin, float)
    z = uniform01(n, seed, 1) * 2.0 - 1.0
    phi = uniform01(n, seed, 2) * 2.0 * np.pi
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = np.stack([s * np.cos(phi), s * np.sin(phi), z], axis=-1)
    r = rmin + (rmax - rmin) * uniform01(n, seed, 3)
    return o, _to_f32_grid(o + d * r[:, None])


def sensor_ring(i, n_sensors=8, radius=10.0, z=0.007):
    """Origin of sensor i of config #5 (sensors on a ring of `radius` m)."""
    a = 2 * np.pi * i / n_sensors
    return np.array([radius * np.cos(a) + 0.013, radius * np.sin(a) + 0.021, z])
