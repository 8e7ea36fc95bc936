"""Soak: N proofs through a pool of provers; reports throughput per block of 100 and host / device memory drift."""
import json, os, resource, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tendermintx_b200 as tmx
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "celestia")
idx = json.load(open(f"{root}/index.json"))["skip_n128_seed3"]
f = tmx.InputDataFetcher(f"{root}/skip_n128_seed3")
th = bytes.fromhex(idx["trusted_hash"])
blob = f.get_skip_inputs(128, idx["trusted"], th, idx["target"])
pub = idx["trusted"].to_bytes(8, "big") + th + idx["target"].to_bytes(8, "big")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
pool = tmx.ProverPool(0, tmx.KIND_SKIP, 128, tmx.CelestiaConfig, in_flight=4)
ref = pool.prove_many([(pub, blob)] * 8)[0]
free0, total = torch.cuda.mem_get_info()
rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
for blk in range(N // 100):
    t0 = time.perf_counter()
    res = pool.prove_many([(pub, blob)] * 100)
    dt = time.perf_counter() - t0
    assert all(r == ref for r in res)
    free, _ = torch.cuda.mem_get_info()
    print(f"block {blk}: {dt * 10:.2f} ms/proof, device memory in use {(total - free) / 2**30:.2f} GiB (drift {(free0 - free) / 2**20:.1f} MiB), "
          f"host max RSS {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024:.0f} MiB (drift {(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - rss0) / 1024:.0f} MiB)")
pool.close()
print("done")
