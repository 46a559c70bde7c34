#!/usr/bin/env python
"""bench.py -- points integrated per second on BASELINE.json configs[1]:
OccupancyMap 2 cm, 131 072-point Velodyne-64-shaped synthetic scan stream,
max_range 30 m (one "step" = insertPointCloud of one scan).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* value  : whole-job points/s with the scans already resident in HBM
           (ufo_b200_insert_device), CUDA events on the map's stream, max over ranks.
* e2e    : the same metric through the reference-facing C-ABI call with HOST (pinned)
           buffers: H2D of the scan and D2H of the scan's result counters are inside
           the timed region (ufo_b200_insert_pointcloud).
* roofline: algorithmic bytes (SURVEY.md 8(d) formula on the live counters) over the
           CUDA-event time of the device work of one insert, against the measured HBM peak.
* cpu_baseline / --impl reference: the unmodified reference (oracle/_ref/libufo_ref.so)
           on a bounded sample of the same workload on the host cores.

Multi-GPU (torchrun, one rank per GPU): weak scaling, one sensor stream + map per GPU
(BASELINE configs[4] shape without the boundary merge); no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RESOLUTION = 0.02
MAX_RANGE = 30.0
RINGS, AZIMUTHS = 64, 2048
WORKLOAD = "OccupancyMap 2 cm depth 16, 131072-pt Velodyne-64-shaped synthetic scan stream, max_range 30 m"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(st, color=False, p_in=12, part="scan"):
    """N*P_in + U*S_leaf + D1*8*S_leaf + sum_{l>=1} D_l*S_inner + sum_{l>=2} D_l*8*S_inner.
    part="update" keeps the terms the leaf-update kernels (k_update_compact / k_update + k_brick_agg) are responsible
    for: everything except the point input and the levels above the brick (depth >= 5)."""
    s_leaf, s_inner = (8, 12) if color else (4, 8)
    n, u = st["points"], st["touched_voxels"]
    d1, d2, d3, d4, up = (st["touched_octets"], st["touched_blocks"], st["touched_d3"],
                          st["touched_bricks"], st["upper_nodes"])
    if part == "update":
        n, up = 0, 0
    inner_all = d1 + d2 + d3 + d4 + up
    inner_ge2 = d2 + d3 + d4 + up
    return n * p_in + u * s_leaf + d1 * 8 * s_leaf + inner_all * s_inner + inner_ge2 * 8 * s_inner


def measured_traffic():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        d = json.load(open(path))
        return float(d["traffic_bytes_per_launch"]), d["source"]
    except Exception:
        return None, None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = "/tmp/ufo_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def make_scans(count, rank):
    """Scans of this rank's sensor: float32 xyz, float64 origins."""
    from ufomap_b200 import scans
    base = scans.sensor_ring(rank, 8) if rank else None
    origins, clouds = [], []
    for k in range(count):
        if rank == 0:
            o, p = scans.velodyne64(k=k, rings=RINGS, azimuths=AZIMUTHS)
        else:
            o = base + np.array([0.25 * k, 0.10 * k, 0.0])
            o, p = scans.velodyne64(k=k, rings=RINGS, azimuths=AZIMUTHS, origin=o,
                                    seed=88172645463325252 + 7919 * rank)
        origins.append(o)
        clouds.append(np.ascontiguousarray(p, dtype=np.float32))
    return origins, clouds


def max_over_ranks(value, world, device=None):
    """MAX-reduce a per-rank scalar (the timed region of the slowest rank is the job's time)."""
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU insertPointCloud on the host cores,
    each step a bounded sample (every `stride`-th point of scan k) of the same workload."""
    if rank != 0:
        return
    import oracle_lib
    kind = "reference" if oracle_lib.have_ref() else "port"
    cls = oracle_lib.RefMap if kind == "reference" else oracle_lib.OracleMap
    if kind == "port":
        oracle_lib.build_oracle()
    stride = 64
    total = args.steps + args.warmup
    origins, clouds = make_scans(total, 0)
    m = cls(RESOLUTION)
    times = []
    for k in range(total):
        pts = clouds[k][::stride].astype(np.float64)
        secs = m.insert(origins[k], pts, max_range=MAX_RANGE)
        if k >= args.warmup:
            times.append(secs)
    npts = len(clouds[0][::stride])
    t = float(np.sum(times))
    value = npts * len(times) / t
    line = {
        "impl": "reference", "metric": "points_integrated_per_s", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 geometry + f32 log-odds", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "every %d-th point of each scan (%d pts/step)" % (stride, npts)},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": 2, "kind": kind,
                         "sample": "insertPointCloud of every %d-th point (%d pts) of scans %d..%d, one map, "
                                   "async=false; the reference uses 1 caller + 1 hit thread of %d host cores"
                                   % (stride, npts, args.warmup, total - 1, os.cpu_count() or 0)},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(origins, clouds):
    """Bounded CPU sample for the main line: ~10-30 s of the reference on scan 0 and 1."""
    import oracle_lib
    kind = "reference" if oracle_lib.have_ref() else "port"
    if kind == "port":
        oracle_lib.build_oracle()
    cls = oracle_lib.RefMap if kind == "reference" else oracle_lib.OracleMap
    stride = 16
    m = cls(RESOLUTION)
    secs, npts = 0.0, 0
    for k in range(2):
        pts = clouds[k][::stride].astype(np.float64)
        t = m.insert(origins[k], pts, max_range=MAX_RANGE)
        if k >= 1:  # scan 0 warms allocator and hash tables
            secs += t
            npts += len(pts)
    m.close()
    return {"value": npts / secs, "unit": "points/s", "cores": 2, "kind": kind,
            "sample": "insertPointCloud of every %d-th point (%d pts) of scan 1 after scan 0 as warm-up, "
                      "async=false, 1 caller + 1 hit thread of %d host cores" % (stride, npts, os.cpu_count() or 0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="sensors", choices=["sensors", "shard"],
                    help="N>1: 'sensors' = one sensor stream + map per GPU (weak scaling, default); "
                         "'shard' = ONE scan stream, broadcast over NCCL, map sharded by brick ownership "
                         "(strong scaling, SURVEY.md 8(e) variant 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ufomap_b200 import capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    total = args.steps + args.warmup
    shard = args.mode == "shard" and world > 1
    origins, clouds = make_scans(total, 0 if shard else rank)
    n_pts = clouds[0].shape[0]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def fresh_map():
        m = capi.Map(RESOLUTION, device=local_rank, initial_bricks=1 << 19)
        m.set_stream(stream.cuda_stream)
        if shard:
            m.set_shard(rank, world)
        return m

    def timed_loop(m, feed):
        for k in range(args.warmup):
            feed(m, k)
        m.wait()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_scan, launches = [], 0
        e0.record(stream)
        for k in range(args.warmup, total):
            feed(m, k)
            st = m.stats()  # D2H of the scan's counters (the step's result)
            per_scan.append(st)
            launches += st["launches"]
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
        return ms, per_scan, launches, clocks

    # ---- device-resident inputs: `value` ---------------------------------
    d_clouds = [torch.from_numpy(c).to(dev) for c in clouds]
    torch.cuda.synchronize(dev)
    m = fresh_map()
    m.set_profiling(1)

    bcast = torch.empty_like(d_clouds[0]) if shard else None

    def feed_device(mm, k):
        src = d_clouds[k]
        if shard:
            # the one collective of the sharded mode: rank 0's scan goes to every GPU (1.5 MB)
            if rank == 0:
                bcast.copy_(src)
            dist.broadcast(bcast, src=0)
            src = bcast
        mm.insert_packed(origins[k], src.data_ptr(), n_pts, capi.XYZ_F32, max_range=MAX_RANGE,
                         on_device=True, async_=True)

    ms_dev, per_scan, launches, clocks = timed_loop(m, feed_device)
    dev_bytes = per_scan[-1]["device_bytes"]
    m.close()
    del d_clouds
    torch.cuda.empty_cache()

    # ---- host buffers through the C ABI: `e2e` -----------------------------
    h_clouds = [torch.from_numpy(c).pin_memory() for c in clouds]
    m = fresh_map()

    stage = torch.empty(n_pts, 3, dtype=torch.float32, device=dev) if shard else None

    def feed_host(mm, k):
        if shard:
            if rank == 0:
                stage.copy_(h_clouds[k], non_blocking=True)  # H2D once, on rank 0
            dist.broadcast(stage, src=0)
            mm.insert_packed(origins[k], stage.data_ptr(), n_pts, capi.XYZ_F32, max_range=MAX_RANGE,
                             on_device=True, async_=True)
            return
        mm.insert_packed(origins[k], h_clouds[k].data_ptr(), n_pts, capi.XYZ_F32, max_range=MAX_RANGE,
                         on_device=False, async_=True)

    ms_e2e, _, _, clocks_e2e = timed_loop(m, feed_host)
    m.close()

    steps = args.steps
    streams = 1 if shard else world  # sharded mode integrates ONE stream with all GPUs
    value = streams * steps * n_pts / (ms_dev * 1e-3)
    e2e = streams * steps * n_pts / (ms_e2e * 1e-3)

    # roofline of the device work of one insert (K1..K4), averaged over the timed scans
    peak, peak_src = measured_peak()
    alg = float(np.mean([algorithmic_bytes(s) for s in per_scan]))
    alg_upd = float(np.mean([algorithmic_bytes(s, part="update") for s in per_scan]))
    t_scan_ms = float(np.mean([s["ms_total"] for s in per_scan]))
    kern = {k: float(np.mean([s[k] for s in per_scan])) for k in
            ("ms_h2d", "ms_points", "ms_rays", "ms_scatter", "ms_update", "ms_propagate")}
    # dominant kernel: the leaf update (k_update_compact + k_brick_agg between two CUDA events)
    achieved = alg_upd / (kern["ms_update"] * 1e-3) / 1e9
    pipeline = alg / (t_scan_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic()
    last = per_scan[-1]

    if rank == 0:
        line = {
            "metric": "points_integrated_per_s", "value": value, "unit": "points/s",
            "scans_per_s": value / n_pts, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "strong" if shard else "weak",
            "vs_baseline": None, "dtype": "f64 geometry + f32 log-odds", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_scan": n_pts, "input": "float32 xyz",
                       "parallelism": ("1 map per GPU" if world == 1 else
                                       ("one scan stream broadcast over NCCL, map sharded by brick ownership"
                                        if shard else "one sensor stream + map per GPU, no merge")),
                       "l2": "per-scan working set (%.1f GB leaf data touched, map %.1f GB) exceeds the 126 MB L2; no flush"
                             % (last["touched_blocks"] * 256 / 1e9, dev_bytes / 1e9)},
            "e2e": {"value": e2e, "unit": "points/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(n_pts * 12), "d2h_bytes_per_step": int(last["result_bytes"])},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k_update_compact (+k_brick_agg): hit/miss log-odds update of the marked voxels and "
                                   "depth 1-4 aggregates; CUDA events on the map's stream around every launch "
                                   "of the timed region",
                         "algorithmic_bytes_per_launch": alg_upd, "ms_per_launch": kern["ms_update"],
                         "traffic_source": traffic_src,
                         # SURVEY.md 8(d)'s whole-pipeline figure: all algorithmic bytes of a scan over the
                         # device time of the whole insert (K1..K4)
                         "pipeline": {"algorithmic_bytes_per_scan": alg, "device_ms_per_scan": t_scan_ms,
                                      "achieved": pipeline, "frac": pipeline / peak}},
            "kernels_ms": kern,
            "counters": {k: int(last[k]) for k in ("rays", "touched_voxels", "hit_voxels", "touched_octets",
                                                    "touched_blocks", "touched_d3", "touched_bricks",
                                                    "upper_nodes", "blocks_in_map", "bricks_in_map", "regrows")},
            "gpu_launches": int(launches),
            "clocks": clocks, "clocks_e2e": clocks_e2e,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_sample(origins, clouds)
            except Exception as e:  # the checker is optional for the GPU number
                line["cpu_baseline"] = {"value": None, "error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
