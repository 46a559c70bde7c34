#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --mode shard 2>&1 | tail -1 > gpurun_out/bench_shard_n2.json
python - <<'PY'
import json
for f in ("bench_n2", "bench_shard_n2"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.1f M e2e %.1f M ms/step %.3f scaling %s launches %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["scaling"], d["gpu_launches"]))
    except Exception as e:
        print(f, "FAILED", e, open("gpurun_out/%s.json" % f).read()[-600:])
PY
