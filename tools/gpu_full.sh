#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_full.json
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json'))
print('value %.1f Mpts/s  e2e %.1f  ms/step %.3f  frac %.3f  cpu %s'%(d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value')))
print(d['kernels_ms']); print(d['clocks'], d['gpu_launches'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json | cut -c1-300
