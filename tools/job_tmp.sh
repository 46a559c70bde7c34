timeout 1100 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/r02_bench_n1.json
bash tools/profile_run.sh r02i > gpurun_out/prof.log 2>&1
