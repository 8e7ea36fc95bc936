# Round-2 bring-up run on one B200: witness / bus / proof parity tests, then a timed N=128 skip proof with phase timings.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_witness.py tests/test_gpu_bus.py tests/test_gpu_prove.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_v2.log
cat gpurun_out/pytest_v2.log
TMX_TIMING=1 timeout 600 python tools/time_prove.py > gpurun_out/time_prove.log 2>&1
tail -80 gpurun_out/time_prove.log
