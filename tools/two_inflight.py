"""Throughput experiment: T host threads, each with its own tmx context + circuit on the same GPU, proving concurrently."""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tendermintx_b200 as tmx
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "celestia")
idx = json.load(open(f"{root}/index.json"))["skip_n128_seed0"]
f = tmx.InputDataFetcher(f"{root}/skip_n128_seed0")
th = bytes.fromhex(idx["trusted_hash"])
blob = f.get_skip_inputs(128, idx["trusted"], th, idx["target"])
pub = idx["trusted"].to_bytes(8, "big") + th + idx["target"].to_bytes(8, "big")
for T in (1, 2, 3, 4):
    ctxs = [tmx.Context(0) for _ in range(T)]
    cs = [tmx.Circuit.build(c, tmx.KIND_SKIP, 128, tmx.CelestiaConfig) for c in ctxs]
    for c in cs:
        for _ in range(2):
            c.prove(pub, blob)
    K = 6
    def work(c):
        for _ in range(K):
            c.prove(pub, blob)
    ths = [threading.Thread(target=work, args=(c,)) for c in cs]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print(f"in flight {T}: {T*K} proofs in {dt*1e3:.1f} ms -> {dt*1e3/(T*K):.2f} ms/proof, {T*K/dt*3600:.0f} proofs/h")
    for c in cs: c.close()
    for c in ctxs: c.close()
