#!/bin/bash
# usage (under gpurun): tools/profile_run.sh <tag>
# 1) launch list of the bench command (all kernels with device time, for kernel SHARES)
# 2) one `--set full` capture of the three hot kernels of a steady-state scan
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rays|k_scatter|k_update" -s 9 -c 3 \
    -o gpurun_out/${TAG}_hot python tools/kbench.py --scans 5 > gpurun_out/${TAG}_hot.log 2>&1
ls -la gpurun_out | tail -5
