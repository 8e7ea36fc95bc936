# does the number of hardware work queues matter? (4 provers x 6 streams on the default 8 connections)
set -x
mkdir -p gpurun_out
for c in 8 32; do for k in 4 6; do
CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 600 python bench.py --steps 100 --no-cpu-baseline --in-flight $k > gpurun_out/bench_c${c}_if$k.json 2>> gpurun_out/bench.err
done; done
python - <<'PY'
import json
for c in (8, 32):
    for k in (4, 6):
        f = f"bench_c{c}_if{k}"
        try:
            d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
            print(f, "ms/proof", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "single", round(d["one_proof_at_a_time"]["ms_per_proof"], 2))
        except Exception as e:
            print(f, "failed", e)
PY
tail -3 gpurun_out/bench.err
