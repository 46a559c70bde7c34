"""The C-ABI shared library loads and exports every symbol include/ufomap_b200.h
declares; argument validation that needs no device."""
import ctypes as C
import os
import re

import pytest

from ufomap_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ufomap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ufo_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = capi.load()
    names = _declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(capi.SYMBOLS)


def test_version_and_defaults():
    lib = capi.load()
    assert b"sm_100a" in lib.ufo_b200_version()
    p = capi.Params()
    lib.ufo_b200_default_params(C.byref(p))
    assert (p.depth_levels, p.prob_hit, p.prob_miss) == (16, 0.7, 0.4)
    assert (p.clamping_thres_min, p.clamping_thres_max) == (0.1192, 0.971)


def test_invalid_depth_levels_rejected():
    # octree.h:931-935: depth_levels outside [2, 21] throws std::invalid_argument
    for bad in (0, 1, 22, 40):
        with pytest.raises(ValueError):
            capi.Map(0.1, depth_levels=bad, device=-2)


def test_geometry_only_handle_refuses_state_calls():
    m = capi.Map(0.1, device=-2)
    with pytest.raises(capi.UfoError) as e:
        m.insert([0, 0, 0], [[1.0, 0, 0]])
    assert e.value.status == capi.E_CUDA  # no CPU fallback
    with pytest.raises(capi.UfoError):
        m.value_field()
    m.close()


def test_no_oracle_in_product_path():
    """The product (ufomap_b200/ and include/) never references oracle/ code."""
    for base in ("ufomap_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "ufo_oracle" not in text and "libufo_ref" not in text and "oracle_lib" not in text, f


def test_device_entry_points_bind_and_fail_loudly_without_a_device(tmp_path):
    """Every wrapper that needs device state reaches the library with a well-formed call (no
    ctypes TypeError) and, on a geometry-only handle, fails with UFO_B200_E_CUDA instead of
    falling back to anything."""
    import numpy as np
    from ufomap_b200 import capi
    m = capi.Map(0.1, device=-2)
    box = (np.zeros(3), np.ones(3))
    pts = np.ones((4, 3))
    rec = np.zeros((4, 16), np.uint8)
    calls = [
        lambda: m.insert([0, 0, 0], pts),
        lambda: m.insert_frame([0, 0, 0], pts, capi.pose_from_rpy(0, 0, 0, 0, 0, 0)),
        lambda: m.insert_pointcloud2([0, 0, 0], rec, 16),
        lambda: m.write(), lambda: m.write(box=box, min_depth=2, expanded=True),
        lambda: m.write_file(str(tmp_path / "x.ufo")), lambda: m.write_file(str(tmp_path / "y.ufo"), box=box),
        lambda: m.write_data(box, 1), lambda: m.set_value_volume(box, 0.2, 0),
        lambda: m.clear_resize(0.2, 12), lambda: m.clear(), lambda: m.value_field(),
        lambda: m.query(np.zeros(1, np.uint64), 0), lambda: m.stats(), lambda: m.wait(),
    ]
    for i, call in enumerate(calls):
        with pytest.raises(capi.UfoError) as e:
            call()
        assert e.value.status == capi.E_CUDA, i
    m.close()
