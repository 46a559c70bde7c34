#!/bin/bash
# usage (under gpurun): tools/profile_run.sh <tag>
# 1) launch list of the bench command (all kernels with device time, for kernel SHARES)
# 2) one `--set full` capture of every kernel of a steady-state scan
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_points|k_split|k_walk_mark|k_gather|k_update_brick|k_upper" -s 30 -c 10 \
    -o gpurun_out/${TAG}_hot python tools/kbench.py --scans 5 > gpurun_out/${TAG}_hot.log 2>&1
ls -la gpurun_out | tail -5
