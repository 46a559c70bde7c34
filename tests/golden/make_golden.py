"""Generates the committed golden fixtures from the UNMODIFIED reference
(oracle/_ref/libufo_ref.so, built from /root/reference by oracle/Makefile).

    python tests/golden/make_golden.py

The reference ships no test vectors of its own (ufomap/tests/CMakeLists.txt is
empty), so these are outputs of the reference itself run in the build container.
Each .npz holds the exact inputs and the reference's results; tests/test_golden.py
replays them through the CPU oracle (no GPU) and the CUDA path (-m gpu).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle_lib import REF_SO, RefMap, _load, build_oracle  # noqa: E402
from ufomap_b200 import scans  # noqa: E402


def scan_case(name, map_kw, inserts, color=False):
    m = RefMap(color=color, **map_kw)
    out = {"n_inserts": len(inserts), "color": int(color)}
    for k, v in map_kw.items():
        out["map_" + k] = v
    for i, ins in enumerate(inserts):
        m.insert(**ins)
        for k, v in ins.items():
            out["ins%d_%s" % (i, k)] = np.asarray(v)
    codes, occ, rgb = m.value_field()
    out["codes"], out["occ"], out["rgb"] = codes, occ, rgb
    ic, idp, iocc, irgb, ifl = m.walk(False)
    keep = np.arange(len(ic))
    out["inner_codes"], out["inner_depths"], out["inner_occ"] = ic[keep], idp[keep], iocc[keep]
    out["inner_rgb"], out["inner_flags"] = irgb[keep], ifl[keep]
    mn, mx = m.change_bbox()
    out["bbox_min"], out["bbox_max"] = mn, mx
    out["sensor_model"] = m.sensor_model()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, len(codes), "voxels", len(ic), "inner nodes")


def indexing_case():
    m = RefMap(0.05)
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.uniform(-200, 200, (400, 3)), rng.uniform(-1, 1, (100, 3)),
                          [[0, 0, 0], [0.05, 0.05, 0.05], [-0.05, 0, 0.025], [1638.4, -1638.4, 0.0]]])
    depths = np.array([0, 1, 2, 5, 9, 15], np.uint32)
    keys = np.zeros((len(depths), len(pts), 3), np.uint32)
    codes = np.zeros((len(depths), len(pts)), np.uint64)
    coords = np.zeros((len(depths), len(pts), 3))
    for i, d in enumerate(depths):
        for j, p in enumerate(pts):
            keys[i, j] = m.to_key(p, int(d))
            codes[i, j] = m.to_code(p, int(d))
            coords[i, j] = m.key_to_coord(keys[i, j], int(d))
    a, b = pts[:80] / 20, pts[80:160] / 20
    specs = [(0, -1.0), (0, 6.0), (1, -1.0), (3, 10.0)]
    rays = {}
    for si, (d, mr) in enumerate(specs):
        flat, lens = [], []
        for x, y in zip(a, b):
            r = m.compute_ray(x, y, mr, d)
            flat.append(r)
            lens.append(len(r))
        rays["ray%d_codes" % si] = np.concatenate(flat)
        rays["ray%d_lens" % si] = np.array(lens)
    ends = pts[:150] / 8
    free = {}
    for fi, (d, simple) in enumerate([(0, False), (2, False), (0, True)]):
        free["free%d" % fi] = np.sort(m.free_set([0.011, -0.02, 0.03], ends, d, simple))
    np.savez_compressed(os.path.join(HERE, "indexing.npz"), resolution=0.05, pts=pts, depths=depths,
                        keys=keys, codes=codes, coords=coords, ray_a=a, ray_b=b,
                        ray_specs=np.array(specs), free_origin=np.array([0.011, -0.02, 0.03]),
                        free_ends=ends, free_specs=np.array([(0, 0), (2, 0), (0, 1)]), **rays, **free)
    print("indexing", len(pts), "points")


def frame_case():
    """insertPointCloud(sensor_origin, cloud, frame_origin, ...): sensor-frame clouds + poses."""
    api = _load(REF_SO, "ufo_ref_")
    m = RefMap(0.1)
    out = {"resolution": 0.1, "max_range": 20.0, "n_inserts": 3}
    rpys = [(0.0, 0.0, 0.0), (0.02, -0.03, 0.7), (-0.4, 1.2, -2.9)]
    for k, rpy in enumerate(rpys):
        o, p = scans.velodyne64(k=k, rings=8, azimuths=96)
        local = (p - o).astype(np.float32).astype(np.float64)   # what a sensor driver delivers
        pose = np.empty(7)
        api["pose_from_rpy"](*[float(v) for v in o], *rpy, pose.ctypes.data)
        world = np.empty_like(local)
        api["transform"](pose.ctypes.data, local.ctypes.data, len(local), world.ctypes.data)
        m.insert(origin=pose[:3], xyz=world, max_range=20.0)
        out["rpy%d" % k], out["pose%d" % k] = np.array(rpy), pose
        out["local%d" % k], out["world%d" % k] = local, world
    out["codes"], out["occ"], out["rgb"] = m.value_field()
    np.savez_compressed(os.path.join(HERE, "frame_velodyne_10cm.npz"), **out)
    print("frame_velodyne_10cm", len(out["codes"]), "voxels")


def file_image_case():
    """Octree::write (uncompressed) of a small colour map: the reference's own file image."""
    m = RefMap(0.08, color=True, automatic_pruning=False)
    out = {"resolution": 0.08, "n_inserts": 2, "max_range": 3.0}
    for k in range(2):
        o, p, c = scans.rgbd(k=k, width=24, height=18)
        m.insert(origin=o, xyz=p, rgb=c, max_range=3.0, discrete=True)
        out["origin%d" % k], out["xyz%d" % k], out["rgb%d" % k] = o, p, c
    image = m.write()
    out["image"] = np.frombuffer(image, np.uint8)
    np.savez_compressed(os.path.join(HERE, "fileimage_rgbd_8cm.npz"), **out)
    print("fileimage_rgbd_8cm", len(image), "bytes")


def main():
    build_oracle()
    frame_case()
    file_image_case()
    o, p = scans.random_shell(n=2000)
    scan_case("shell_16cm", dict(resolution=0.16), [dict(origin=o, xyz=p, max_range=5.0)])
    ins = []
    for k in range(3):
        o, p = scans.velodyne64(k=k, rings=8, azimuths=128)
        ins.append(dict(origin=o, xyz=p, max_range=25.0))
    scan_case("velodyne_stream_20cm", dict(resolution=0.2), ins)
    ins = []
    for k in range(2):
        o, p, c = scans.rgbd(k=k, width=40, height=30)
        ins.append(dict(origin=o, xyz=p, rgb=c, max_range=3.0, discrete=True))
    scan_case("rgbd_color_discrete_4cm", dict(resolution=0.04, automatic_pruning=False), ins, color=True)
    o, p, c = scans.rgbd(width=40, height=30)
    scan_case("rgbd_depth1_simple_5cm", dict(resolution=0.05),
              [dict(origin=o, xyz=p, max_range=3.0, depth=1), dict(origin=o, xyz=p, max_range=2.5, simple=True)])
    indexing_case()


if __name__ == "__main__":
    main()
