timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "async or single_ray or empty_and_unsupported" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_a.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],'clocks',d['clocks'],'cold',d['cold_start'])
P
