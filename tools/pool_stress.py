"""Stress: many proofs through one pool (steady state) must all equal the single-prover proof."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tendermintx_b200 as tmx
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "celestia")
T, N = int(sys.argv[1]), int(sys.argv[2])
cases = []
for seed in range(4):
    name = f"skip_n128_seed{seed}"
    idx = json.load(open(f"{root}/index.json"))[name]
    f = tmx.InputDataFetcher(f"{root}/{name}")
    th = bytes.fromhex(idx["trusted_hash"])
    cases.append((idx["trusted"].to_bytes(8, "big") + th + idx["target"].to_bytes(8, "big"), f.get_skip_inputs(128, idx["trusted"], th, idx["target"]), idx["target_hash"]))
ctx = tmx.Context(0)
c = tmx.Circuit.build(ctx, tmx.KIND_SKIP, 128, tmx.CelestiaConfig)
want = [c.prove(p, b) for p, b, _ in cases]
for (p, b, h), (proof, out) in zip(cases, want):
    assert out.hex() == h
    c.verify(proof, p, out)
pool = tmx.ProverPool(0, tmx.KIND_SKIP, 128, tmx.CelestiaConfig, in_flight=T)
stm = [(cases[i % 4][0], cases[i % 4][1]) for i in range(N)]
res = pool.prove_many(stm)
bad = [i for i, r in enumerate(res) if r != want[i % 4]]
print(f"{N} proofs, {T} in flight, 4 different statements interleaved: {len(bad)} mismatches {bad[:10]}")
