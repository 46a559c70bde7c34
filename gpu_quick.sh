#!/bin/bash
# quick GPU check: parity subset + bench line summary
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not leaving" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_quick.json
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json'))
print('value %.1f Mpts/s  e2e %.1f  ms/step %.3f  frac %.3f'%(d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['roofline']['frac']))
print(d['kernels_ms']); print(d['clocks'])"
