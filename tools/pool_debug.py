"""Debug aid: do provers in flight produce identical proofs?  Reports the words that differ."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tendermintx_b200 as tmx
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "celestia")
idx = json.load(open(f"{root}/index.json"))["skip_n128_seed0"]
f = tmx.InputDataFetcher(f"{root}/skip_n128_seed0")
th = bytes.fromhex(idx["trusted_hash"])
blob = f.get_skip_inputs(128, idx["trusted"], th, idx["target"])
pub = idx["trusted"].to_bytes(8, "big") + th + idx["target"].to_bytes(8, "big")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = tmx.Context(0)
c = tmx.Circuit.build(ctx, tmx.KIND_SKIP, 128, tmx.CelestiaConfig)
ref, out = c.prove(pub, blob)
c.verify(ref, pub, out)
ref2, _ = c.prove(pub, blob)
print("single prover reproducible:", ref == ref2)
refw = np.frombuffer(ref, dtype=np.uint64)
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for rnd in range(ROUNDS):
    pool = tmx.ProverPool(0, tmx.KIND_SKIP, 128, tmx.CelestiaConfig, in_flight=T)  # fresh contexts: first-use paths race
    res = pool.prove_many([(pub, blob)] * (2 * T))
    for j, (p, o) in enumerate(res):
        w = np.frombuffer(p, dtype=np.uint64)
        if w.size != refw.size or not np.array_equal(w, refw):
            d = np.nonzero(w[:refw.size] != refw[:w.size])[0]
            print(f"round {rnd} proof {j} (prover {j % T}): {d.size} words differ, first {d[:6]}, last {d[-3:]}, sizes {w.size} {refw.size}")
            try:
                c.verify(p, pub, o); print("   ... but it verifies")
            except Exception as e:
                print("   ... and does NOT verify:", str(e)[:100])
    pool.close()
print("done")
