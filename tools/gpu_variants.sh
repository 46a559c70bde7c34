#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/kbench.py 2>&1 | tail -1
for v in variants/*.so; do [ -f $v ] && UFOMAP_B200_LIB=$PWD/$v python tools/kbench.py 2>&1 | tail -1; done
python tools/kbench.py --scans 24 2>&1 | tail -1
