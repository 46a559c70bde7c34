#!/bin/bash
# 2-GPU bench lines (run with: gpurun --gpus 2 -- tools/gpu_n2.sh): routed (strong), merge (config #5), replicas
mkdir -p gpurun_out
for mode in route merge sensors; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 10 --warmup 3 --sustain 0 --mode $mode > gpurun_out/r02_bench_n2_$mode.json 2> gpurun_out/r02_bench_n2_$mode.err
  echo "== $mode rc=$?"; tail -c 400 gpurun_out/r02_bench_n2_$mode.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_n2_$mode.json") if l.startswith("{")][-1])
    print("$mode", "value %.1f M" % (d["value"] / 1e6), "e2e %.1f M" % (d["e2e"]["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"], d["kernels_ms"])
except Exception as e:
    print("$mode: no line", e)
PY
done
