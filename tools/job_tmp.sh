python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench_err.log
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['regrows'],'sus',d['sustained']['value'],d['clocks'],d['cpu_baseline']['value'])
P
bash tools/sanitize.sh dense > gpurun_out/sanitize.log 2>&1; cat gpurun_out/sanitize.log
