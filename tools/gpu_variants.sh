#!/bin/bash
# kernel timings of the default build and of every variants/*.so on the bench workload
python tools/kbench.py --scans 8 2>&1 | tail -1
for v in variants/*.so; do [ -f $v ] && UFOMAP_B200_LIB=$PWD/$v python tools/kbench.py --scans 8 2>&1 | tail -1; done
