#!/bin/bash
# compute-sanitizer memcheck + racecheck over small parity cases (every marking path and K3/K4,
# mono and colour, update and set mode), under gpurun; summaries land in gpurun_out/ (copied to
# profiles/ when kept).  usage: tools/sanitize.sh [modes...]   (default: dense probe records)
mkdir -p gpurun_out
MODES=${@:-dense probe records}
SEL="single_ray or config1_plumbing or insert_depth_1_and_2 or rays_leaving_through_plus or change_detection or color_map_plain_cloud or set_value_volume"
for tool in memcheck racecheck; do
  for mode in $MODES; do
    UFO_B200_MARK=$mode timeout 900 compute-sanitizer --tool $tool --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/r02_sanitizer_${tool}_${mode}.log 2>&1
    echo "$tool $mode: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_${tool}_${mode}.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/r02_sanitizer_${tool}_${mode}.log | tail -1)"
  done
done
