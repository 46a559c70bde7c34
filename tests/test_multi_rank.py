"""world_size-2 CPU (gloo) coverage of the N>1 path of bench.py."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_rank_sharding_and_max_reduce_gloo():
    out = _torchrun([os.path.join("tests", "dist_worker.py")], 29541)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("DIST_OK world=2") == 1


def test_reference_arm_prints_once_under_torchrun():
    out = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], 29542)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["unit"] == "points/s"
