"""`skip` / `step` entry points with the reference's two subcommands [REF bin/skip.rs:1-27, bin/step.rs:1-26,
succinct.json:8-9,15-16]:

    python -m tendermintx_b200.cli skip build [--n-max 128] [--chain-id celestia] [--out build/main.circuit]
    python -m tendermintx_b200.cli skip prove input.json [--fixtures DIR] [--circuit build/main.circuit] [--out output.json]

`input.json`  = {"type": "req_bytes", "data": {"input": "0x<48 or 40 bytes>"}}
`output.json` = {"type": "res_bytes", "data": {"proof": "0x...", "output": "0x<32 bytes>"}}

The off-chain inputs come from a fixture directory in the RPC JSON layout (env TMX_FIXTURE_DIR or --fixtures);
the reference's RPC mode needs a network and is out of scope.
"""
import argparse
import ctypes
import json
import os
import sys

from . import KIND_SKIP, KIND_STEP, SKIP_MAX, Circuit, Context, TendermintConfig, TmxError, _check, lib


def _load_circuit(ctx, path):
    h = ctypes.c_void_p()
    _check(lib().tmx_circuit_load(ctx.handle, path.encode(), ctypes.byref(h)))
    c = Circuit.__new__(Circuit)
    c.ctx, c._h = ctx, h
    return c


def main(argv=None):
    ap = argparse.ArgumentParser(prog="tendermintx_b200.cli")
    ap.add_argument("circuit", choices=["skip", "step"])
    sub = ap.add_subparsers(dest="cmd", required=True)
    b = sub.add_parser("build")
    b.add_argument("--n-max", type=int, default=100)  # REF circuits/consts.rs:4
    b.add_argument("--chain-id", default="celestia")
    b.add_argument("--skip-max", type=int, default=SKIP_MAX)
    b.add_argument("--out", default="build/main.circuit")
    b.add_argument("--device", type=int, default=0)
    p = sub.add_parser("prove")
    p.add_argument("input")
    p.add_argument("--fixtures", default=os.environ.get("TMX_FIXTURE_DIR"))
    p.add_argument("--circuit-file", default="build/main.circuit")
    p.add_argument("--out", default="output.json")
    p.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    kind = KIND_SKIP if args.circuit == "skip" else KIND_STEP
    try:
        ctx = Context(args.device)
        if args.cmd == "build":
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            c = Circuit.build(ctx, kind, args.n_max, TendermintConfig(args.chain_id, args.skip_max))
            c.save(args.out)
            print(f"wrote {args.out} (digest {[hex(x) for x in c.digest()]})")
            return 0
        if not args.fixtures:
            ap.error("prove needs --fixtures DIR (or TMX_FIXTURE_DIR)")
        with open(args.input) as f:
            req = json.load(f)
        if req.get("type") != "req_bytes":
            raise TmxError(1, "input.json: expected type req_bytes")
        pub = bytes.fromhex(req["data"]["input"].removeprefix("0x"))
        c = _load_circuit(ctx, args.circuit_file)
        proof, out = c.prove_fixture(pub, args.fixtures)
        c.verify(proof, pub, out)
        with open(args.out, "w") as f:
            json.dump({"type": "res_bytes", "data": {"proof": "0x" + proof.hex(), "output": "0x" + out.hex()}}, f)
        print(f"wrote {args.out}: output 0x{out.hex()}")
        return 0
    except TmxError as e:
        print(f"error: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
