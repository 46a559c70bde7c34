timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 4 --steps 10 --warmup 3 --sustain 0 > gpurun_out/r02_bench_n4_route.json 2> gpurun_out/r02_bench_n4_route.err
echo "rc=$?"; tail -c 300 gpurun_out/r02_bench_n4_route.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n4_route.json') if l.startswith('{')][-1])
print('route N=4 value %.1f M e2e %.1f M ms/step %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), d['kernels_ms'], d['config']['mode'])
P
