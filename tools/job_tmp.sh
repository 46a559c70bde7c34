bash tools/gpu_variants.sh
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_err.log
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_a.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e'])
P
