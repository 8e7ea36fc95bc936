# thread-safety / leak check of the multi-stream prover: 4 statements interleaved through pools of 2..6 provers, then a soak
set -x
mkdir -p gpurun_out
for t in 2 4 6; do timeout 600 python tools/pool_stress.py $t 240; done 2>&1 | grep -v "^+" | tee gpurun_out/pool_stress.txt
timeout 900 python tools/soak.py 1200 2>&1 | tee gpurun_out/soak.txt | tail -14
