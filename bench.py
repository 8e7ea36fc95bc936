"""bench.py -- skip proofs/hour (CelestiaConfig, VALIDATOR_SET_SIZE_MAX = 128) on N B200s, with the Goldilocks
LDE roofline and the CPU oracle timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete skip proof (witness tables -> commitments -> quotients -> FRI -> proof bytes) of the
synthetic 128-validator celestia chain committed under tests/golden/celestia (seed = rank, so ranks prove
independent statements: weak scaling, no data-path collective; the only collective is the NCCL broadcast of the
circuit artefact at start-up).  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MAX = 128
METRIC = "skip proofs/hour (CelestiaConfig, 128 val)"
UNIT = "proofs/hour"
WORKLOAD = "skip circuit CelestiaConfig VALIDATOR_SET_SIZE_MAX=128 (synthetic celestia chain, 128 signers, seed=rank)"


def load_case(seed):
    with open(os.path.join(ROOT, "tests", "golden", "celestia", "index.json")) as f:
        idx = json.load(f)[f"skip_n128_seed{seed % 8}"]
    return os.path.join(ROOT, "tests", "golden", "celestia", f"skip_n128_seed{seed % 8}"), idx


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def run_reference(args):
    """The CPU implementation of the path on the box's host cores: the repo's deterministic oracle prover (the
    Rust reference cannot be built here: no cargo, un-vendored dependencies), all OpenMP threads, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from oracle import tm_inputs as ti

    fixture, idx = load_case(0)
    src = ti.FixtureSource(fixture)
    th = bytes.fromhex(idx["trusted_hash"])
    blob = ti.skip_inputs(src, N_MAX, idx["trusted"], th, idx["target"])
    pub = ti.skip_public_input(idx["trusted"], th, idx["target"])
    budget_s = 200.0
    times = []
    t_all = time.perf_counter()
    while len(times) < max(1, args.steps) and (not times or time.perf_counter() - t_all + times[-1] < budget_s):
        t0 = time.perf_counter()
        status, proof, out = oracle.prove(pub, blob, "celestia")
        times.append(time.perf_counter() - t0)
        assert status == "OK" and out.hex() == idx["target_hash"]
    mean = sum(times) / len(times)
    v = 3600.0 / mean
    cores = os.cpu_count()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": 0,
        "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "prover": "CPU oracle (restatement of the plonky2-style pipeline), OpenMP"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full proof(s) of the workload (bounded to ~{int(budget_s)} s)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import tendermintx_b200 as tmx

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ctx = tmx.Context(local_rank)
    cfg = tmx.CelestiaConfig

    # ---- build on rank 0, NCCL-broadcast the circuit artefact, load everywhere ----
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, f"main.{rank}.circuit")
        if rank == 0:
            c0 = tmx.Circuit.build(ctx, tmx.KIND_SKIP, N_MAX, cfg)
            c0.save(path)
            c0.close()
            data = np.fromfile(path, dtype=np.uint8)
        else:
            data = None
        if world > 1:
            from tendermintx_b200 import sharding

            blob_bytes = sharding.broadcast_bytes(data.tobytes() if rank == 0 else b"", 0, device="cuda")
            with open(path, "wb") as fh:
                fh.write(blob_bytes)
        h = ctypes.c_void_p()
        rc = tmx.lib().tmx_circuit_load(ctx.handle, path.encode(), ctypes.byref(h))
        assert rc == 0, tmx.lib().tmx_last_error()
        circuit = tmx.Circuit.__new__(tmx.Circuit)
        circuit.ctx, circuit.kind, circuit.n_max, circuit.config, circuit._h = ctx, tmx.KIND_SKIP, N_MAX, cfg, h

    # ---- inputs: host-side assembly from the fixture directory (C++), kept in host memory ----
    fixture, idx = load_case(rank)
    fetcher = tmx.InputDataFetcher(fixture)
    th = fetcher.header_hash(idx["trusted"])
    assert th.hex() == idx["trusted_hash"]
    t0 = time.perf_counter()
    blob = fetcher.get_skip_inputs(N_MAX, idx["trusted"], th, idx["target"])
    assemble_ms = (time.perf_counter() - t0) * 1e3
    pub = idx["trusted"].to_bytes(8, "big") + th + idx["target"].to_bytes(8, "big")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = ctx.torch_stream()
    # ---- warm-up (also the correctness gate: output header and CPU verification) ----
    proof = out = None
    for _ in range(max(args.warmup, 1)):
        proof, out = circuit.prove(pub, blob)
    assert out.hex() == idx["target_hash"], "proved header differs from the fixture's block hash"
    circuit.verify(proof, pub, out)

    # ---- value: K proofs from HBM-resident inputs, CUDA events on the prover's stream, max over ranks ----
    circuit.set_inputs(blob)
    barrier()
    launches0 = ctx.launch_count()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            circuit.prove(pub, None)
        e1.record(stream)
        e1.synchronize()
        dev_ms = e0.elapsed_time(e1)
        barrier()
    launches = ctx.launch_count() - launches0
    # ---- e2e: the public call with HOST buffers (blob H2D, proof bytes D2H inside the timed region) ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proof, out = circuit.prove(pub, blob)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([dev_ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    value = world * args.steps / (dev_ms / 1e3) * 3600.0
    e2e_value = world * args.steps / (e2e_ms / 1e3) * 3600.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the LDE kernel family (K1) on the largest table of this proof ----
    peak, peak_src = measured_peaks()
    dims = tmx.Context.trace_dims(tmx.KIND_SKIP, N_MAX)
    rows, cols = dims[2]
    log_n = rows.bit_length() - 1
    vals = torch.randint(0, 2**62, (cols, rows), dtype=torch.int64, device="cuda")
    lde = torch.empty((cols, rows * 2), dtype=torch.int64, device="cuda")
    coef = torch.empty((cols, rows), dtype=torch.int64, device="cuda")
    for _ in range(3):
        ctx.lde(vals, log_n, 1, out=lde, coeffs=coef)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.lde(vals, log_n, 1, out=lde, coeffs=coef)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    lde_ms = sum(ts) / len(ts)
    alg_bytes = 8 * rows * cols * (1 + 2)
    achieved = alg_bytes / (lde_ms / 1e3) / 1e9
    dig = None
    tm = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dig = ctx.poseidon_merkle(lde, log_n + 1, 4, digests=dig)
        b.record()
        b.synchronize()
        tm.append(a.elapsed_time(b))
    merkle_ms = min(tm)
    perms = (rows * 2) * ((cols + 7) // 8) + rows * 2
    del vals, lde, coef, dig

    # ---- CPU baseline: the oracle prover on this box's host cores, one full proof of the same workload ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle

        t0 = time.perf_counter()
        status, want, _ = oracle.prove(pub, blob, "celestia")
        cpu_s = time.perf_counter() - t0
        same = status == "OK" and np.array_equal(np.frombuffer(proof, dtype=np.uint64), want)
        cpu = {"value": 3600.0 / cpu_s, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "1 full proof of the workload by the CPU oracle prover (OpenMP, all cores)",
               "proof_bytes_equal_gpu": bool(same)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (Goldilocks field, bytes/bits in the witness kernels)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "tables": {"sha256": dims[0], "sha512": dims[1], "ed25519": dims[2]},
                   "l2": "working set per proof (traces + LDEs, about 4 GB) is far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"{world} independent proofs, one per GPU", "host_input_assembly_ms": assemble_ms},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": len(blob), "d2h_bytes_per_step": len(proof) + 224 + N_MAX,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "ntt_pass_kernel family (K1: iNTT + coset LDE of the Ed25519 table, "
                     f"{cols} cols x 2^{log_n}, rate 1/2)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes": alg_bytes, "ms": lde_ms},
        "kernels": {"poseidon_merkle_ms": merkle_ms, "poseidon_Mperm_per_s": perms / merkle_ms / 1e3,
                    "note": "K2 is bound by 32-bit integer multiply issue, not HBM"},
        "proof_bytes": len(proof),
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
