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
    # one flipped bit anywhere: header, caps of both rounds and bus totals (the first few hundred words), then openings,
    # FRI caps, final polynomials, proof-of-work witnesses and query data of the five tables
    for pos in list(range(0, 8)) + [8, 100, 200, 300, 5000] + [int(x) for x in rng.integers(8, proof.size, 24)]:
        bad = proof.copy()
        bad[pos] ^= 1
        assert oracle.verify_proof(bad, pub, "mocha-4", 0, 2, out) != 0, pos
        with pytest.raises(tmx.TmxError):
            tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, bad.tobytes(), pub, out)
    # a word that is not a canonical field element (x + p with x < 2^32 - 1 names the same element), a trailing extra word
    small = int(np.argmax(proof[8:] < 2**32 - 1)) + 8
    assert proof[small] < 2**32 - 1
    bad = proof.copy()
    bad[small] += np.uint64(2**64 - 2**32 + 1)
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, bad.tobytes(), pub, out)
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, proof.tobytes() + bytes(8), pub, out)
    assert oracle.verify_proof(np.concatenate([proof, np.zeros(1, dtype=np.uint64)]), pub, "mocha-4", 0, 2, out) != 0
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


@pytest.mark.parametrize("table", [0, 1, 2, 4])
def test_cheating_prover_with_one_bad_cell_is_rejected(oracle, proved, table):
    """Soundness end to end: a prover that commits to a witness with ONE wrong cell (after witness generation, so no
    assertion stops it) still emits a well-formed proof; the constraint identity at the out-of-domain point no longer
    holds and both verifiers reject it."""
    import ctypes

    import tendermintx_b200 as tmx

    c, pub, blob, proof, out = proved
    rng = np.random.default_rng(7 + table)
    for _ in range(3):
        col, row = int(rng.integers(0, 2000)), int(rng.integers(0, 1 << 20))
        oracle.lib().tm_debug_corrupt_next_proof(ctypes.c_int(table), ctypes.c_size_t(col), ctypes.c_size_t(row))
        status, bad, out2 = oracle.prove(pub, blob, "mocha-4")
        assert status == "OK" and out2 == out and bad.size == proof.size and not np.array_equal(bad, proof)
        assert oracle.verify_proof(bad, pub, "mocha-4", 0, 2, out) != 0, (table, col, row)
        with pytest.raises(tmx.TmxError):
            tmx.verify_proof(tmx.KIND_STEP, 2, tmx.Mocha4Config, bad.tobytes(), pub, out)
    # and the hook is one-shot: the next proof is the honest one again
    assert np.array_equal(oracle.prove(pub, blob, "mocha-4")[1], proof)
