#!/usr/bin/env python
"""How much do the insert kernels of two scans gain from running concurrently?
Two independent maps (own streams) receive the same scan stream with asynchronous inserts, so the
kernels of map A's scan and map B's scan can share the GPU.  Prints scans/s for one map alone and for
the pair.  Grid sizes of the two long kernels: UFO_B200_K3_BLOCKS / UFO_B200_WALK_BLOCKS."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ufomap_b200 import capi, scans  # noqa: E402

N, W = 16, 4
clouds = [scans.velodyne64(k=k) for k in range(N + W)]
clouds = [(o, np.ascontiguousarray(p, dtype=np.float32)) for o, p in clouds]


def run(maps):
    for k in range(W):
        for m in maps:
            m.insert(clouds[k][0], clouds[k][1], max_range=30.0, dtype=np.float32, async_=True)
    for m in maps:
        m.wait()
    t0 = time.perf_counter()
    for k in range(W, W + N):
        for m in maps:
            m.insert(clouds[k][0], clouds[k][1], max_range=30.0, dtype=np.float32, async_=True)
    for m in maps:
        m.wait()
    return (time.perf_counter() - t0) / N * 1e3


a = capi.Map(0.02, initial_bricks=1 << 19)
one = run([a])
del a
a = capi.Map(0.02, initial_bricks=1 << 19)
b = capi.Map(0.02, initial_bricks=1 << 19)
two = run([a, b])
print("K3_BLOCKS=%s WALK_BLOCKS=%s  one map: %.3f ms/scan   two maps: %.3f ms per PAIR of scans (%.3f ms/scan, x%.2f)" % (
    os.environ.get("UFO_B200_K3_BLOCKS", "-"), os.environ.get("UFO_B200_WALK_BLOCKS", "-"), one, two, two / 2, 2 * one / two))
