"""Profiling driver: N_MAX=128 skip proofs of the synthetic celestia chain (the bench workload), nothing else.
Usage: python tools/profile_prove.py [n_proofs]   (run under ncu; see profiles/README.md)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tendermintx_b200 as tmx
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "celestia")
idx = json.load(open(f"{root}/index.json"))["skip_n128_seed0"]
ctx = tmx.Context(0)
c = tmx.Circuit.build(ctx, tmx.KIND_SKIP, 128, tmx.CelestiaConfig)
f = tmx.InputDataFetcher(f"{root}/skip_n128_seed0")
th = bytes.fromhex(idx["trusted_hash"])
blob = f.get_skip_inputs(128, idx["trusted"], th, idx["target"])
pub = idx["trusted"].to_bytes(8, "big") + th + idx["target"].to_bytes(8, "big")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
l0 = ctx.launch_count()
for i in range(n):
    proof, out = c.prove(pub, blob)
print("launches per proof", (ctx.launch_count() - l0) // n, "proof bytes", len(proof))
