K="python tools/kbench.py --res 0.002 --range 5 --shape rgbd --discrete --scans 6 --bricks 1048576"
$K 2>&1 | tail -1
for v in variants/*.so; do UFOMAP_B200_LIB=$PWD/$v $K 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "color or set_value or frame or pointcloud2" 2>&1 | tail -2
