#!/usr/bin/env python
"""Kernel-level timing of the insert pipeline on the bench workload (or another config).
usage: kbench.py [--res 0.02] [--range 30] [--scans 8] [--shape velodyne|rgbd] [--color]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ufomap_b200 import capi, scans  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=float, default=0.02)
ap.add_argument("--range", type=float, default=30.0)
ap.add_argument("--scans", type=int, default=8)
ap.add_argument("--shape", default="velodyne")
ap.add_argument("--discrete", action="store_true")
ap.add_argument("--bricks", type=int, default=1 << 19)
args = ap.parse_args()

color = args.shape == "rgbd"
m = capi.Map(args.res, color=color, initial_bricks=args.bricks)
m.set_profiling(1)
rows = []
for k in range(args.scans):
    if args.shape == "velodyne":
        o, p = scans.velodyne64(k=k)
        m.insert(o, p, max_range=args.range, dtype=np.float32, discrete=args.discrete)
    else:
        o, p, c = scans.rgbd(k=k)
        m.insert(o, p, rgb=c, max_range=args.range, dtype=np.float32, discrete=True)
    st = m.stats()
    rows.append(st)
keys = ("ms_total", "ms_points", "ms_rays", "ms_scatter", "ms_update", "ms_propagate")
tail = rows[2:] if len(rows) > 3 else rows
print(os.environ.get("UFOMAP_B200_LIB", "default"), " ".join("%s=%.3f" % (k[3:], np.mean([r[k] for r in tail])) for k in keys),
      "first_scan_total=%.3f" % rows[0]["ms_total"], "U=%d blocks=%d regrows=%d" % (rows[-1]["touched_voxels"], rows[-1]["blocks_in_map"], sum(r["regrows"] for r in rows)))
