"""GPU debug: compare the prover pipeline stage by stage with the oracle (run under gpurun)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
import tendermintx_b200 as tmx

P = 2**64 - 2**32 + 1
c = {x["name"]: x for x in json.load(open("tests/golden/fixture_vectors.json"))["cases"]}["step_10000_n2"]
blob = bytes.fromhex(c["blob"])
tr = oracle.build_traces(blob)
ctx = tmx.Context(0)
O = oracle.lib()
alpha = np.array([123456789123, 987654321987], dtype=np.uint64)
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
host = lambda t: t.cpu().numpy().view(np.uint64)
O.tm_debug_set_shape(ctypes.c_uint32(0), ctypes.c_uint32(c["n_max"]))  # step circuit: the SHA-256 public columns depend on the shape
for t in range(3):
    C, n = tr[t].shape
    lg = n.bit_length() - 1
    lde = np.zeros((C, 2 * n), dtype=np.uint64); qv = np.zeros((2, 2 * n), dtype=np.uint64)
    O.tm_debug_quotient(t, vp(np.ascontiguousarray(tr[t])), ctypes.c_size_t(n), ctypes.c_size_t(C), vp(alpha), vp(lde), vp(qv))
    d_lde = ctx.lde(dev(tr[t]), lg, 1)
    print("table", t, "lde equal", np.array_equal(host(d_lde), lde))
    d_q = torch.zeros((2, 2 * n), dtype=torch.int64, device="cuda")
    rc = tmx.lib().tmx_quotient(ctx.handle, 0, c["n_max"], t, ctypes.c_void_p(d_lde.data_ptr()), lg, vp(alpha), ctypes.c_void_p(d_q.data_ptr()), ctx._stream())
    torch.cuda.synchronize()
    got = host(d_q)
    print("  quotient rc", rc, "equal", np.array_equal(got, qv), "mismatch count", int((got != qv).sum()))
    # inverse NTT of 2 columns vs oracle
    inv = host(ctx.ntt(d_q.clone(), lg + 1, inverse=True))
    want = np.stack([oracle.ntt(qv[i], inverse=True) for i in range(2)])
    print("  intt(2 cols) equal", np.array_equal(inv, want))
    os.makedirs("gpurun_out", exist_ok=True)
    np.save(f"gpurun_out/q_gpu_{t}.npy", got)

