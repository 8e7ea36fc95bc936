# Round-end style check on one B200: GPU parity tests, smoke(), bench (ours + reference arm).
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err
timeout 600 python tools/config_sweep.py --steps 8 > gpurun_out/config_sweep.json 2> gpurun_out/config_sweep.err; cut -c1-1500 gpurun_out/config_sweep.json; tail -2 gpurun_out/config_sweep.err
