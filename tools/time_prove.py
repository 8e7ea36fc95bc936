import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tendermintx_b200 as tmx
root = "tests/golden/celestia"
idx = json.load(open(f"{root}/index.json"))["skip_n128_seed0"]
ctx = tmx.Context(0)
c = tmx.Circuit.build(ctx, tmx.KIND_SKIP, 128, tmx.CelestiaConfig)
f = tmx.InputDataFetcher(f"{root}/skip_n128_seed0")
th = bytes.fromhex(idx["trusted_hash"])
blob = f.get_skip_inputs(128, idx["trusted"], th, idx["target"])
pub = idx["trusted"].to_bytes(8, "big") + th + idx["target"].to_bytes(8, "big")
timing = os.environ.pop("TMX_TIMING", None)
t0 = time.perf_counter()
for i in range(3):
    proof, out = c.prove(pub, blob)
print("warm-up proofs ms", (time.perf_counter() - t0) * 1e3 / 3, "proof bytes", len(proof))
c.verify(proof, pub, out)
os.environ["TMX_TIMING"] = "1"
t0 = time.perf_counter(); c.prove(pub, blob); print("timed prove (with syncs) ms", (time.perf_counter() - t0) * 1e3)
del os.environ["TMX_TIMING"]
for i in range(3):
    l0 = ctx.launch_count()
    t0 = time.perf_counter(); c.prove(pub, blob); print("plain prove ms", (time.perf_counter() - t0) * 1e3, "launches", ctx.launch_count() - l0)
print("phase ms (lde, merkle) per witness table", c.last_phase_ms())
