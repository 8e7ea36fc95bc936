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
for i in range(3):
    c.prove(pub, blob)
os.environ["TMX_TIMING"] = "1"
t0 = time.perf_counter(); c.prove(pub, blob); print("timed prove (with syncs) ms", (time.perf_counter() - t0) * 1e3)
del os.environ["TMX_TIMING"]
t0 = time.perf_counter(); c.prove(pub, blob); print("plain prove ms", (time.perf_counter() - t0) * 1e3)
