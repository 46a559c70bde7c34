#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-rx}
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu -k "not leaving" 2>&1 | tail -3
python tools/kbench.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:"k_rays|k_scatter|k_update" -s 9 -c 3 -o gpurun_out/${TAG}_all python tools/kbench.py --scans 5 > gpurun_out/${TAG}_all.log 2>&1
