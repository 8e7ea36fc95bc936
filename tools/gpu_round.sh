# Round-end style check on one B200: GPU parity tests, smoke(), bench (ours + reference arm).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err
