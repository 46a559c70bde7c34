#!/usr/bin/env python
"""Full-size parity + timing script (test infrastructure: it loads the reference harness under oracle/).
BASELINE config #3 (OccupancyMapColor 2 mm, 640x480 RGB-D, 5 m, insertPointCloudDiscrete) at
insert_depth 4 -- the setting at which the reference's CPU path can run this config at all
(depth 0 exceeds 62 GB of host RAM).  Times the CUDA path and the unmodified
reference on the same full-size scans and compares the occupancy of sampled voxels bit for bit.
usage (on the GPU box): python tests/full_size/config3_depth4.py [--scans 3] [--samples 200000]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import RefMap, have_ref  # noqa: E402
from ufomap_b200 import capi, scans  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=3)
ap.add_argument("--samples", type=int, default=200000)
ap.add_argument("--depth", type=int, default=4)
args = ap.parse_args()

gpu = capi.Map(0.002, color=True, initial_bricks=1 << 19)
gpu.set_profiling(1)
ref = RefMap(0.002, color=True) if have_ref() else None
rng = np.random.default_rng(3)
gpu_ms, ref_s, sample_pts = [], [], []
for k in range(args.scans):
    o, p, c = scans.rgbd(k=k)
    gpu.insert(o, p, rgb=c, max_range=5.0, depth=args.depth, discrete=True, dtype=np.float32)
    gpu_ms.append(gpu.stats()["ms_total"])
    if ref is not None:
        t0 = time.perf_counter()
        ref.insert(origin=o, xyz=p, rgb=c, max_range=5.0, depth=args.depth, discrete=True)
        ref_s.append(time.perf_counter() - t0)
    # sample positions: along random rays (free space) and at end points (hits)
    idx = rng.integers(0, len(p), args.samples // args.scans)
    frac = np.concatenate([rng.uniform(0.02, 1.0, len(idx) // 2), np.ones(len(idx) - len(idx) // 2)])
    sample_pts.append(o + (p[idx] - o) * frac[:, None])
n = len(p)
print("points/scan", n, "gpu ms/scan", ["%.3f" % v for v in gpu_ms],
      "-> %.1f M points/s (steady)" % (n / (np.mean(gpu_ms[1:]) * 1e-3) / 1e6))
if ref is not None:
    print("reference s/scan", ["%.3f" % v for v in ref_s],
          "-> %.3f M points/s; speed-up %.0fx" % (n / np.mean(ref_s[1:]) / 1e6,
                                                  np.mean(ref_s[1:]) / (np.mean(gpu_ms[1:]) * 1e-3)))
    pts = np.concatenate(sample_pts)
    codes = np.array([gpu.to_code(q, 0) for q in pts], np.uint64)
    occ, flags, rgb = gpu.query(codes, 0)
    bad = 0
    for i, cd in enumerate(codes):
        _, rocc, rrgb, _, _ = ref.node(int(cd), 0)
        if np.float32(occ[i]).view(np.uint32) != np.float32(rocc).view(np.uint32):
            bad += 1
    print("sampled voxels", len(codes), "occupancy mismatches", bad,
          "| occupied %d free %d unknown %d" % ((occ > 0).sum(), (occ < 0).sum(), (occ == 0).sum()))
    sys.exit(1 if bad else 0)
