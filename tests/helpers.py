"""Shared comparison helpers for the parity tests."""
import numpy as np

from oracle_lib import aggregate_level


def assert_value_fields_equal(a, b, color_tol=0, what=""):
    ca, va, ga = a
    cb, vb, gb = b
    assert len(ca) == len(cb), "%s: %d vs %d voxels" % (what, len(ca), len(cb))
    assert np.array_equal(ca, cb), "%s: voxel code sets differ" % what
    # float log-odds: bit-exact (the north star allows +-1 ULP; we hold 0)
    bad = np.nonzero(va.view(np.uint32) != vb.view(np.uint32))[0]
    assert len(bad) == 0, "%s: %d log-odds differ, first code %x: %r vs %r" % (
        what, len(bad), int(ca[bad[0]]), va[bad[0]], vb[bad[0]])
    if color_tol == 0:
        assert np.array_equal(ga, gb), "%s: colours differ" % what
    else:
        d = np.abs(ga.astype(np.int32) - gb.astype(np.int32))
        assert d.max(initial=0) <= color_tol, "%s: colour error %d" % (what, d.max())


def check_inner_against_field(gpu_map, field, model, levels=(1, 2, 3, 4, 5, 6, 9), max_nodes=200000):
    """Inner aggregates of the CUDA map == pure function of the value field."""
    codes, occ, rgb = field
    for lvl in levels:
        if lvl > gpu_map.depth_levels:
            continue
        parents, mx, free, unk = aggregate_level(codes, occ, rgb, lvl, model)
        if len(parents) > max_nodes:
            sel = np.linspace(0, len(parents) - 1, max_nodes).astype(np.int64)
            parents, mx, free, unk = parents[sel], mx[sel], free[sel], unk[sel]
        q = parents << np.uint64(3 * lvl)
        o, f, _ = gpu_map.query(q, lvl)
        assert np.array_equal(o.view(np.uint32), mx.view(np.uint32)), "level %d max mismatch" % lvl
        assert np.array_equal((f & 1).astype(bool), free), "level %d contains_free mismatch" % lvl
        assert np.array_equal(((f >> 1) & 1).astype(bool), unk), "level %d contains_unknown mismatch" % lvl
