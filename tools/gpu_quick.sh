set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_quick.log
timeout 300 python tools/microbench.py 16,1217,1 > gpurun_out/microbench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/pytest_quick.log gpurun_out/microbench.log gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
