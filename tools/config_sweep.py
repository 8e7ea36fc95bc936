"""Times the BASELINE.json configurations other than the headline one on ONE GPU (they are parity-test cases, not bench
lines; this is the table in DESIGN.md section 7).  For every configuration: one proof at a time and `in_flight` proofs
in the ProverPool, host buffers in, proof bytes out (the e2e shape of bench.py), device time between two CUDA events
recorded while the GPU is idle; every proof is verified on the CPU and its output compared with the fixture's block hash.

Usage: python tools/config_sweep.py [--steps 8] [--in-flight 4] > gpurun_out/config_sweep.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tendermintx_b200 as tmx  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "celestia")


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1), out


def statement(name, idx, kind, n_max):
    f = tmx.InputDataFetcher(os.path.join(GOLDEN, name))
    e = idx[name]
    if kind == tmx.KIND_SKIP:
        th = bytes.fromhex(e["trusted_hash"])
        blob = f.get_skip_inputs(n_max, e["trusted"], th, e["target"])
        pub = e["trusted"].to_bytes(8, "big") + th + e["target"].to_bytes(8, "big")
    else:
        th = bytes.fromhex(e["trusted_hash"])
        blob = f.get_step_inputs(n_max, e["trusted"], th)
        pub = e["trusted"].to_bytes(8, "big") + th
    return pub, blob, e["target_hash"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--in-flight", type=int, default=4)
    args = ap.parse_args()
    idx = json.load(open(os.path.join(GOLDEN, "index.json")))
    cfg = tmx.CelestiaConfig
    cases = [
        ("skip N=128 (configs[1], the bench workload)", tmx.KIND_SKIP, 128, ["skip_n128_seed0"]),
        ("step N=128 consecutive header (configs[2])", tmx.KIND_STEP, 128, ["step_n128_seed0"]),
        ("skip N=256 (configs[3])", tmx.KIND_SKIP, 256, ["skip_n256_seed0"]),
        ("batch of 8 independent skip N=128 proofs, seeds 0..7, on one GPU (configs[4] runs it one per GPU)", tmx.KIND_SKIP, 128,
         [f"skip_n128_seed{s}" for s in range(8)]),
    ]
    rows = []
    for label, kind, n_max, names in cases:
        ctx = tmx.Context(0)
        circuit = tmx.Circuit.build(ctx, kind, n_max, cfg)
        pool = tmx.ProverPool(0, kind, n_max, cfg, in_flight=args.in_flight)
        stmts = [statement(n, idx, kind, n_max) for n in names]
        # warm-up = correctness gate
        for pub, blob, want in stmts:
            proof, out = circuit.prove(pub, blob)
            circuit.verify(proof, pub, out)
            assert out.hex() == want, (label, out.hex(), want)
        pool.prove_many([(s[0], s[1]) for s in stmts] * max(1, args.in_flight // len(stmts)))
        work = [(s[0], s[1]) for s in stmts] * max(1, args.steps // len(stmts))
        one_ms, res1 = timed(lambda: [circuit.prove(p, b) for p, b in work])
        pool_ms, res2 = timed(lambda: pool.prove_many(work))
        assert [r[0] for r in res1] == [r[0] for r in res2], "pool and single prover disagree"
        dims = tmx.Context.trace_dims(kind, n_max)
        rows.append({"config": label, "tables_rows_cols": dims, "proofs_timed": len(work), "proof_bytes": len(res1[0][0]),
                     "one_at_a_time_ms_per_proof": one_ms / len(work), "pool_ms_per_proof": pool_ms / len(work),
                     "pool_proofs_per_hour": len(work) / (pool_ms / 1e3) * 3600.0, "in_flight": args.in_flight})
        print(json.dumps(rows[-1]), flush=True)
        pool.close()
        circuit.close()
        ctx.close()


if __name__ == "__main__":
    main()
