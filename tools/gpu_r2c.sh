set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
for k in 2 6 8; do timeout 600 python bench.py --steps 60 --no-cpu-baseline --in-flight $k > gpurun_out/bench_if$k.json 2>> gpurun_out/bench.err; done
python - <<'PY'
import json
for f in ["bench", "bench_if2", "bench_if6", "bench_if8"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "in_flight", d["arm"]["in_flight"], "ms/proof", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "single", round(d["one_proof_at_a_time"]["ms_per_proof"], 2),
              "K1", [round(x, 2) for x in d["roofline"]["ms_per_table"]], "K2", [round(x, 2) for x in d["kernels"]["k2_ms_per_table"]], "frac", round(d["roofline"]["frac"], 4), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
bash tools/gpu_ncu_k1.sh > gpurun_out/ncu_k1.log 2>&1
cat gpurun_out/ncu/ntt_summary.txt | tail -80
cat gpurun_out/ncu/leaf_summary.txt | head -40
