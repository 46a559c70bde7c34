"""Host-side pieces of bench.py that the driver's run depends on (no GPU needed)."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_formula():
    """N*P_in + U*S_leaf + D1*8*S_leaf + sum_{l>=1} D_l*S_inner + sum_{l>=2} D_l*8*S_inner."""
    b = _bench()
    st = dict(points=10, touched_voxels=100, touched_octets=40, touched_blocks=20, touched_d3=8,
              touched_bricks=3, upper_nodes=5)
    mono = 10 * 12 + 100 * 4 + 40 * 8 * 4 + (40 + 20 + 8 + 3 + 5) * 8 + (20 + 8 + 3 + 5) * 8 * 8
    assert b.algorithmic_bytes(st, False, 12) == mono
    upd = 100 * 4 + 40 * 8 * 4 + (40 + 20 + 8 + 3) * 8 + (20 + 8 + 3) * 8 * 8
    assert b.algorithmic_bytes(st, False, 12, part="update") == upd
    color = 10 * 16 + 100 * 8 + 40 * 8 * 8 + (40 + 20 + 8 + 3 + 5) * 12 + (20 + 8 + 3 + 5) * 8 * 12
    assert b.algorithmic_bytes(st, True, 16) == color


def test_committed_measurement_files_are_readable():
    """bench.py folds three committed measurement files into its JSON line."""
    b = _bench()
    traffic, src = b.measured_traffic()
    assert traffic > 1e9 and "profiles/" in src
    ceil = b.atomic_ceiling()
    assert ceil and ceil["l2_resident_gops"] > ceil["at_128mb_gops"] > 0
    sc = b.sector_ceiling()
    assert sc and sc["k3_all_ms"] > sc["k3_leaf_ms"] > 0
    assert sc["rmw_gbs_at_density_1.0"] > sc["rmw_gbs_at_density_0.24"] > sc["rmw_gbs_at_density_0.06"]
    peak, peak_src = b.measured_peak()
    assert 1000.0 < peak < 10000.0 and peak_src


def test_clock_sampler_never_raises_without_a_gpu():
    b = _bench()
    s = b.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(("sm_mhz", "sm_max_mhz", "reasons", "samples")) <= set(out)
    json.dumps(out)


def test_configs_match_baseline_json():
    """--config 2|3|4 are BASELINE.json's single-GPU configurations."""
    b = _bench()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert set(b.CONFIGS) == {2, 3, 4}
    assert b.CONFIGS[2]["resolution"] == 0.02 and b.CONFIGS[2]["max_range"] == 30.0 and not b.CONFIGS[2]["color"]
    assert b.CONFIGS[3]["resolution"] == 0.002 and b.CONFIGS[3]["color"]
    assert b.CONFIGS[4]["resolution"] == 0.05 and b.CONFIGS[4]["max_range"] == 100.0
    assert "points" in json.dumps(base).lower()
