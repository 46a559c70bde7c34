#!/bin/bash
# usage: scripts_profile.sh <tag>   (run under gpurun; writes gpurun_out/<tag>_*)
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rays -s 3 -c 1 -o gpurun_out/${TAG}_rays $CMD > gpurun_out/${TAG}_rays.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_update -s 3 -c 1 -o gpurun_out/${TAG}_update $CMD > gpurun_out/${TAG}_update.log 2>&1
ls -la gpurun_out/
