set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/time_prove.py > gpurun_out/time_prove.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_2proofs.csv python tools/profile_prove.py 2 > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'leaf_hash_kernel|ntt_pass_kernel|quotient_kernel|merkle_subtree|ed25519_ladder|ed25519_expand|fri_batch|eval_columns|pow_grind' -s 300 -c 60 -o gpurun_out/prof_r1b python tools/profile_prove.py 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
