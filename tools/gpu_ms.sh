# quick check of a prover change on one B200: proof equality tests, then the bench with a few in-flight depths
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_prove.py tests/test_gpu_bus.py -m gpu -x -q 2>&1 | tail -3

for k in 4 2 1; do timeout 600 python bench.py --steps 100 --no-cpu-baseline --in-flight $k > gpurun_out/bench_if$k.json 2>> gpurun_out/bench.err; done
python - <<'PY'
import json
for f in ["bench_if4", "bench_if2", "bench_if1"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "in_flight", d["arm"]["in_flight"], "ms/proof", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "single", round(d["one_proof_at_a_time"]["ms_per_proof"], 2),
              "K1", [round(x, 2) for x in d["roofline"]["ms_per_table"]], "K2", [round(x, 2) for x in d["kernels"]["k2_ms_per_table"]], "frac", round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench.err
