"""The exact ray-splitting scheme planned for K2 (experiments/ray_split_poc.py): its model of the
reference walk is the oracle's walk, and split walks reproduce it voxel for voxel."""
import importlib.util
import os

import numpy as np

from oracle_lib import OracleMap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ray_split_poc", os.path.join(ROOT, "experiments", "ray_split_poc.py"))
poc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(poc)


def test_split_walks_equal_the_oracle_walk():
    res, levels = 0.02, 16
    orc = OracleMap(res, depth_levels=levels)
    rng = np.random.default_rng(11)
    checked = 0
    for case in range(300):
        a = rng.uniform(-3, 3, 3)
        b = a + rng.normal(size=3) * rng.uniform(0.05, 6.0)
        if case % 5 == 1:       # exact diagonal from a voxel corner: three-way ties at every step
            a = np.round(a / res) * res
            b = a + np.array([1, 1, 1]) * res * int(rng.integers(3, 120)) * rng.choice([-1, 1])
        a32, b32 = a.astype(np.float32).astype(np.float64), b.astype(np.float32).astype(np.float64)
        s = poc.dda_init(a32, b32, res, levels)
        if s is None:
            continue
        full = poc.walk(s["cur"], s["end"], s["step"], s["t_delta"], s["t_max"], s["dist"])
        want = np.sort(orc.free_set(b32, a32[None, :]))
        got = np.sort(np.array([orc.key_to_code(np.array(v, np.uint32), 0) for v in full], np.uint64))
        assert np.array_equal(want, got), case
        for parts in (2, 4):
            assert [v for p in poc.split_walk(s, parts) for v in p] == full, (case, parts)
        checked += 1
    assert checked > 250
