"""CPU-side proof tests: the oracle prover is deterministic, its proofs verify under BOTH verifiers (oracle C and
the product's C++ verifier, which needs no GPU), and tampering / wrong statements are rejected."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def proved(oracle):
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        c = {x["name"]: x for x in json.load(f)["cases"]}["step_10000_n2"]
    pub, blob = bytes.fromhex(c["input"]), bytes.fromhex(c["blob"])
    status, proof, out = oracle.prove(pub, blob, "mocha-4")
    assert status == "OK" and out.hex() == c["expected_output"]
    return c, pub, blob, proof, out


def test_oracle_proof_verifies_and_is_deterministic(oracle, proved):
    c, pub, blob, proof, out = proved
    assert oracle.verify_proof(proof, pub, "mocha-4", 0, 2, out) == 0
    _, proof2, _ = oracle.prove(pub, blob, "mocha-4")
    assert np.array_equal(proof, proof2)


def test_product_verifier_accepts_oracle_proof(proved):
    import tendermintx_b200 as tmx

    c, pub, blob, proof, out = proved
    tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, proof.tobytes(), pub, out)


def test_both_verifiers_reject_tampering(oracle, proved):
    import tendermintx_b200 as tmx

    c, pub, blob, proof, out = proved
    rng = np.random.default_rng(0)
    for pos in [8, 100, 5000] + [int(x) for x in rng.integers(8, proof.size, 6)]:
        bad = proof.copy()
        bad[pos] ^= 1
        assert oracle.verify_proof(bad, pub, "mocha-4", 0, 2, out) != 0, pos
        with pytest.raises(tmx.TmxError):
            tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, bad.tobytes(), pub, out)
    # wrong public input / output / circuit parameters
    bad_pub = bytearray(pub)
    bad_pub[3] ^= 1
    assert oracle.verify_proof(proof, bytes(bad_pub), "mocha-4", 0, 2, out) != 0
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, proof.tobytes(), bytes(bad_pub), out)
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(tmx.KIND_STEP, 2, tmx.CelestiaConfig, proof.tobytes(), pub, out)
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, proof.tobytes()[:-8], pub, out)


def test_unsatisfied_statement_is_not_provable(oracle, proved):
    c, pub, blob, proof, out = proved
    b = bytearray(blob)
    b[920 + 32 + 3] ^= 0x40
    status, p, _ = oracle.prove(pub, bytes(b), "mocha-4")
    assert status == "SIGNATURE" and p is None
