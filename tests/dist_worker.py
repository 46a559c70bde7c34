"""Run under torchrun with the gloo backend (CPU): checks the host-side logic of bench.py's
multi-GPU path -- per-rank sensor streams, max-over-ranks timing, rank-0-only output."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
origins, clouds, layout, xyzs, rgbs = bench.make_scans(bench.CONFIGS[2], 2, rank)
assert clouds[0].shape == (131072, 3) and clouds[0].dtype == np.float32 and layout == 1
# every rank integrates its own sensor: origins differ between ranks
gathered = [None] * world
dist.all_gather_object(gathered, origins[0].tolist())
assert len({tuple(g) for g in gathered}) == world, gathered
# routed mode: the ranks' azimuth sectors partition the rays of one scan
mine = bench.route_slice(131072, rank, world)
parts = [None] * world
dist.all_gather_object(parts, mine.tolist())
allidx = np.sort(np.concatenate([np.asarray(p_) for p_ in parts]))
assert np.array_equal(allidx, np.arange(131072)), "sectors must cover every ray exactly once"
assert len(mine) == 131072 // world and (np.diff(mine) > 0).all()
# the job's time is the slowest rank's
ms = bench.max_over_ranks(10.0 + rank, world)
assert ms == 10.0 + world - 1, ms
dist.barrier()
if rank == 0:
    print("DIST_OK world=%d" % world)
dist.destroy_process_group()
