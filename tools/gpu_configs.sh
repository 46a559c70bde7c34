#!/bin/bash
# other BASELINE configs at full size (parity for them is tested at oracle-feasible sizes)
mkdir -p gpurun_out
{
echo "# config 4: OccupancyMap 5 cm, Velodyne-64 131072 pts, max_range 100 m"
timeout 300 python tools/kbench.py --res 0.05 --range 100 --scans 8 2>&1 | tail -1
echo "# config 3: OccupancyMapColor 2 mm, RGB-D 640x480 (307200 pts), max_range 5 m, insertPointCloudDiscrete"
timeout 300 python tools/kbench.py --res 0.002 --range 5 --scans 4 --shape rgbd 2>&1 | tail -1
echo "# config 2 with insertPointCloudDiscrete (the variant the ROS server calls)"
timeout 300 python tools/kbench.py --discrete 2>&1 | tail -1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
} | tee gpurun_out/configs.txt
