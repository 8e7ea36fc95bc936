set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/pytest_quick.log; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "one-at-a-time", round(d["one_proof_at_a_time"]["ms_per_proof"],2))
print("K1 per table", [round(x,2) for x in d["roofline"]["ms_per_table"]], "K2 per table", [round(x,2) for x in d["kernels"]["k2_ms_per_table"]], "roofline", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],4))
PY
