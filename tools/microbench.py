"""Kernel micro-benchmarks (run on the B200 box): LDE and Poseidon Merkle at trace-like shapes.
Prints one line per shape: algorithmic GB/s for K1 (8*n*B*(1+2^r) bytes), permutations/s for K2."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import numpy as np
import torch

import tendermintx_b200 as tmx

P = 2**64 - 2**32 + 1


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    ctx = tmx.Context(0)
    shapes = [(16, 256, 1), (16, 1024, 1), (18, 135, 3), (20, 135, 3), (12, 512, 1)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(x) for x in s.split(",")) for s in sys.argv[1:]]
    for log_n, n_cols, r in shapes:
        n = 1 << log_n
        vals = torch.randint(0, 2**62, (n_cols, n), dtype=torch.int64, device="cuda")
        out = torch.empty((n_cols, n << r), dtype=torch.int64, device="cuda")
        coeffs = torch.empty((n_cols, n), dtype=torch.int64, device="cuda")
        best, mean = timed(lambda: ctx.lde(vals, log_n, r, out=out, coeffs=coeffs))
        alg = 8 * n * n_cols * (1 + (1 << r))
        print(json.dumps({"kernel": "lde", "log_n": log_n, "cols": n_cols, "rate_bits": r, "ms_best": best,
                          "ms_mean": mean, "alg_GBps": alg / best / 1e6}))
        best, mean = timed(lambda: ctx.ntt(vals, log_n))
        print(json.dumps({"kernel": "ntt", "log_n": log_n, "cols": n_cols, "ms_best": best,
                          "alg_GBps": 16 * n * n_cols / best / 1e6}))
        dig = torch.empty((tmx.lib().tmx_merkle_digest_count(log_n + r, 4), 4), dtype=torch.int64, device="cuda")
        best, mean = timed(lambda: ctx.poseidon_merkle(out, log_n + r, 4, digests=dig))
        perms = (n << r) * ((n_cols + 7) // 8) + (n << r)
        print(json.dumps({"kernel": "poseidon_merkle", "log_rows": log_n + r, "cols": n_cols, "ms_best": best,
                          "Mperm_per_s": perms / best / 1e3, "read_GBps": 8 * n_cols * (n << r) / best / 1e6}))
        del vals, out, coeffs, dig
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
