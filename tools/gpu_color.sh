#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in "" variants/*.so; do
  if [ -n "$v" ]; then export UFOMAP_B200_LIB=$PWD/$v; fi
  timeout 300 python tools/kbench.py --res 0.002 --range 5 --scans 4 --shape rgbd 2>&1 | tail -1
done
unset UFOMAP_B200_LIB
timeout 300 python tools/kbench.py 2>&1 | tail -1
