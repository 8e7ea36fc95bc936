"""Native `skip` / `step` entry points (tendermintx_b200/bin, built from csrc/bin/entrypoint.cpp over the C ABI): the
reference's bin/skip.rs / bin/step.rs with plonky2x's `build` and `prove input.json` subcommands [REF succinct.json:8-9,15-16]."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "tendermintx_b200", "bin")


def test_binaries_exist_and_fail_loudly_without_a_gpu(tmp_path):
    import torch

    for name in ("skip", "step"):
        assert os.access(os.path.join(BIN, name), os.X_OK), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([os.path.join(BIN, "skip")], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([os.path.join(BIN, "skip"), "build", "--circuit-file", str(tmp_path / "main.circuit")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_skip_build_then_prove_input_json(ctx, tmp_path):
    """`skip build` -> ./build/main.circuit; `skip prove input.json` -> output.json; same proof bytes as the Python mirror."""
    import tendermintx_b200 as tmx

    with open(os.path.join(HERE, "golden", "celestia", "index.json")) as f:
        idx = json.load(f)["skip_n16_seed0"]
    fixtures = os.path.join(HERE, "golden", "celestia", "skip_n16_seed0")
    pub = idx["trusted"].to_bytes(8, "big") + bytes.fromhex(idx["trusted_hash"]) + idx["target"].to_bytes(8, "big")
    circuit_file = str(tmp_path / "build" / "main.circuit")
    r = subprocess.run([os.path.join(BIN, "skip"), "build", "--n-max", "16", "--circuit-file", circuit_file], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(circuit_file)
    inp, outp = str(tmp_path / "input.json"), str(tmp_path / "output.json")
    with open(inp, "w") as f:
        json.dump({"type": "req_bytes", "data": {"input": "0x" + pub.hex()}}, f)
    r = subprocess.run([os.path.join(BIN, "skip"), "prove", inp, "--fixtures", fixtures, "--circuit-file", circuit_file, "--out", outp],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(outp) as f:
        res = json.load(f)
    assert res["type"] == "res_bytes" and res["data"]["output"] == "0x" + idx["target_hash"]
    circuit = tmx.Circuit.build(ctx, tmx.KIND_SKIP, 16, tmx.CelestiaConfig)
    proof, out = circuit.prove_fixture(pub, fixtures)
    assert res["data"]["proof"] == "0x" + proof.hex()
    circuit.close()
    # a request the circuit cannot satisfy exits non-zero, as the reference's prover would panic
    with open(inp, "w") as f:
        json.dump({"type": "req_bytes", "data": {"input": "0x" + (pub[:40] + (idx["trusted"] + 1).to_bytes(8, "big")).hex()}}, f)
    r = subprocess.run([os.path.join(BIN, "skip"), "prove", inp, "--fixtures", fixtures, "--circuit-file", circuit_file, "--out", outp],
                       capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.gpu
def test_step_build_then_prove_input_json(ctx, tmp_path):
    with open(os.path.join(HERE, "golden", "celestia", "index.json")) as f:
        idx = json.load(f)["step_n128_seed0"]
    fixtures = os.path.join(HERE, "golden", "celestia", "step_n128_seed0")
    pub = idx["trusted"].to_bytes(8, "big") + bytes.fromhex(idx["trusted_hash"])
    circuit_file = str(tmp_path / "main.circuit")
    assert subprocess.run([os.path.join(BIN, "step"), "build", "--n-max", "128", "--circuit-file", circuit_file]).returncode == 0
    inp, outp = str(tmp_path / "input.json"), str(tmp_path / "output.json")
    with open(inp, "w") as f:
        json.dump({"type": "req_bytes", "data": {"input": "0x" + pub.hex()}}, f)
    r = subprocess.run([os.path.join(BIN, "step"), "prove", inp, "--fixtures", fixtures, "--circuit-file", circuit_file, "--out", outp],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(outp) as f:
        assert json.load(f)["data"]["output"] == "0x" + idx["target_hash"]
