"""Cheating provers.  Each test plays a prover that wants a FALSE statement (or an invalid witness) accepted: it skips the
statement pre-check (the host-side assertions that stop an honest prover), takes the witness tables the generators produce
from its doctored inputs -- or patches cells afterwards -- and runs the complete protocol honestly from there, so it emits a
well-formed proof.  Both verifiers (the oracle's C verifier and the product's C++ verifier, `tmx_verify`) must reject every one
of them, and accept the honest proof of the same statement.  This is what makes `tmx_verify == OK` mean `verify_skip` /
`verify_step` holds [REF circuits/builder/verify.rs:469-563]: the checks live in the proof (bus, range tables, logic table,
public-input terms), not in the prover's host code."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
P = 2**64 - 2**32 + 1
T_SHA256, T_SHA512, T_ED, T_LOGIC, T_RANGE = range(5)


def _case(name):
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        c = {x["name"]: x for x in json.load(f)["cases"]}[name]
    return c, bytes.fromhex(c["input"]), bytes.fromhex(c["blob"]), (1 if c["kind"] == "skip" else 0)


def _both_reject(oracle, proof, pub, kind, n_max, out, chain="mocha-4"):
    import tendermintx_b200 as tmx

    rc = oracle.verify_proof(proof, pub, chain, kind, n_max, out)
    assert rc != 0, "the oracle verifier accepted a cheating proof"
    cfg = tmx.Mocha4Config if chain == "mocha-4" else tmx.CelestiaConfig
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(kind, n_max, cfg, proof.tobytes(), pub, out)
    return rc


def _both_accept(oracle, proof, pub, kind, n_max, out):
    import tendermintx_b200 as tmx

    assert oracle.verify_proof(proof, pub, "mocha-4", kind, n_max, out) == 0
    tmx.verify_proof(kind, n_max, tmx.Mocha4Config, proof.tobytes(), pub, out)


@pytest.fixture(scope="module")
def honest(oracle):
    c, pub, blob, kind = _case("skip_3000_3100_n4")
    status, proof, out = oracle.prove(pub, blob, "mocha-4")
    assert status == "OK" and out.hex() == c["expected_output"]
    _both_accept(oracle, proof, pub, kind, c["n_max"], out)
    return c, pub, blob, kind, proof, out


def _cheat(oracle, pub, blob, patches=(), expect_status="OK"):
    oracle.cheat_next_proof(skip_precheck=True, patches=patches)
    status, proof, out = oracle.prove(pub, blob, "mocha-4")
    assert status == expect_status
    return proof, out


def test_prover_that_skips_the_statement_check_and_claims_another_output(oracle, honest):
    """out32 = a header of the prover's choosing; every table honest for the real one."""
    c, pub, blob, kind, proof, out = honest
    b = bytearray(blob)
    b[32 + 5] ^= 0x01  # tmx_offchain_head.header (the proven target header)
    bad, bad_out = _cheat(oracle, pub, bytes(b))
    assert bad_out != out
    assert _both_reject(oracle, bad, pub, kind, c["n_max"], bad_out) == 200  # unbalanced bus: header-proof roots != claimed output


def test_honest_tables_for_another_trusted_header(oracle, honest):
    """The public input names a different trusted header than the one the validator-hash proof leads to."""
    c, pub, blob, kind, proof, out = honest
    p2 = bytearray(pub)
    p2[8 + 3] ^= 0x80
    bad, bad_out = _cheat(oracle, bytes(p2), blob)
    assert _both_reject(oracle, bad, bytes(p2), kind, c["n_max"], bad_out) == 200


def test_honest_tables_for_another_validator_set(oracle, honest):
    """Witness tables of a different (internally consistent) statement: the step fixture's inputs cannot be passed off under the
    public input of another height."""
    c, pub, blob, kind, proof, out = honest
    c2, pub2, blob2, kind2 = _case("skip_10000_10500_n4")
    bad, bad_out = _cheat(oracle, pub, blob2)  # tables of 10000 -> 10500 under the public input 3000 -> 3100
    _both_reject(oracle, bad, pub, kind, c["n_max"], bad_out)


def test_limb_above_16_bits_with_compensating_carry(oracle, honest):
    """The classic attack on an unchecked multiplication gadget: w = wlo + 2^16 whi, so (wlo + 2^16, whi - 1) satisfies every
    polynomial constraint of the Ed25519 table; only the range check of wlo notices."""
    c, pub, blob, kind, proof, out = honest
    ED_MUL, STRIDE, WLO, WHI = 100, 63, 33, 48
    tabs = oracle.all_traces(blob, "mocha-4", public_input=pub)
    ed = tabs[T_ED]
    row = 5
    g = next(g for g in range(14) if ed[ED_MUL + g * STRIDE + WHI + 3, row] >= 1)
    col_lo, col_hi = ED_MUL + g * STRIDE + WLO + 3, ED_MUL + g * STRIDE + WHI + 3
    bad, bad_out = _cheat(oracle, pub, blob, patches=[(T_ED, col_lo, row, 1 << 16), (T_ED, col_hi, row, P - 1)])
    assert _both_reject(oracle, bad, pub, kind, c["n_max"], out) == 200  # the lookup of wlo >= 2^16 has no provider


def test_bad_signature(oracle, honest):
    """One signature is invalid (a bit of R flipped); the prover does not care and proves anyway."""
    c, pub, blob, kind, proof, out = honest
    b = bytearray(blob)
    n_signed_before = sum(b[920 + 240 * i + 236] for i in range(4))
    assert b[920 + 236] == 1
    b[920 + 32 + 3] ^= 0x40  # sig_r of validator 0
    oracle.cheat_next_proof(skip_precheck=True)
    status, bad, bad_out = oracle.prove(pub, bytes(b), "mocha-4")
    # either the flipped R is not even a curve point (no Ed25519 table can be built: nothing to verify), or the proof is rejected
    if status == "OK":
        _both_reject(oracle, bad, pub, kind, c["n_max"], bad_out)
    else:
        assert status == "SIGNATURE" and n_signed_before >= 1


def test_unsigned_validator_counted_as_signed(oracle, honest):
    """A validator that did not sign is flagged as signed to reach the 2/3 threshold: its slot then has to verify a real
    signature, which it does not have."""
    c, pub, blob, kind, proof, out = honest
    c2, pub2, blob2, kind2 = _case("step_10500_n4_with_dummy")
    b = bytearray(blob2)
    unsigned = [i for i in range(4) if b[920 + 240 * i + 236] == 0]
    assert unsigned
    b[920 + 240 * unsigned[0] + 236] = 1
    oracle.cheat_next_proof(skip_precheck=True)
    status, bad, bad_out = oracle.prove(pub2, bytes(b), "mocha-4")
    if status == "OK":
        _both_reject(oracle, bad, pub2, kind2, c2["n_max"], bad_out)
    else:
        assert status == "SIGNATURE"


def test_trusted_validator_flagged_without_a_matching_signer(oracle, honest):
    """Raising the trusted voting power: set the `signed on target` flag of a trusted validator in the logic table (and its
    running sum) without a signed target validator of that key."""
    c, pub, blob, kind, proof, out = honest
    lt, st = oracle.logic_trace(pub, blob, "mocha-4")
    assert st == 0
    circ = oracle.circuit(kind, c["n_max"], "mocha-4")
    sel_h1 = circ.table_data(T_LOGIC)[1][50]  # LGK_SEL + LT_H1
    rows = np.nonzero(sel_h1)[0]
    H1_FLAG = 178
    unflagged = [int(r) for r in rows[:c["n_max"]] if lt[H1_FLAG, r] == 0]
    if not unflagged:
        pytest.skip("every trusted validator of the fixture signed the target block")
    bad, bad_out = _cheat(oracle, pub, blob, patches=[(T_LOGIC, H1_FLAG, unflagged[0], 1)])
    _both_reject(oracle, bad, pub, kind, c["n_max"], out)


def test_padding_slot_of_the_trusted_set_cannot_carry_voting_power(oracle, honest):
    """The trusted validators hash binds only the ENABLED slots.  A prover fills a padding slot of the trusted set with the key
    of a signed target validator and a huge voting power, flags it as 'signed on target' and keeps every running sum
    consistent: the 1/3 threshold would be met by data nobody committed to.  (The reference sums `voting_power` over all
    flagged slots, enabled or not [REF verify.rs:392-431, voting.rs:69-86]; here a flag needs an enabled row.)"""
    import struct

    c, pub, blob, kind = _case("skip_10000_10500_n4")
    n = c["n_max"]
    nb_trusted = struct.unpack_from("<I", blob, 16)[0]
    assert nb_trusted < n
    signer = next(i for i in range(n) if blob[920 + 240 * i + 236])
    slot, power = nb_trusted, 10**12
    b = bytearray(blob)
    tf = 920 + 240 * n + 48 * slot
    b[tf:tf + 32] = blob[920 + 240 * signer:920 + 240 * signer + 32]
    struct.pack_into("<QI", b, tf + 32, power, 37 + 6)  # 10^12 is a six-byte varint
    blob2 = bytes(b)
    status, proof, out = oracle.prove(pub, blob2, "mocha-4")  # padding data is free: the honest proof still goes through
    assert status == "OK"
    _both_accept(oracle, proof, pub, kind, n, out)
    lt, st = oracle.logic_trace(pub, blob2, "mocha-4")
    assert st == 0
    circ = oracle.circuit(kind, n, "mocha-4")
    rows = [int(r) for r in np.nonzero(circ.table_data(T_LOGIC)[1][50])[0]]  # H1 rows: trusted leaves, then target leaves
    H1_FLAG, H1_SUM, H1_SUM4, H1_DF, H1_DF2, H1_MK = 178, 197, 201, 202, 206, 400 + 16
    assert lt[H1_FLAG, rows[slot]] == 0

    def limbs(col, row):
        return sum(int(lt[col + k, row]) << (16 * k) for k in range(4))

    patches = [(T_LOGIC, H1_FLAG, rows[slot], 1), (T_LOGIC, H1_MK, rows[n + signer], int(lt[H1_MK, rows[n + signer]]) + 1)]
    for i in range(slot, n):
        v = limbs(H1_SUM, rows[i]) + power
        patches += [(T_LOGIC, H1_SUM + k, rows[i], (v >> (16 * k)) & 0xFFFF) for k in range(4)]
        patches.append((T_LOGIC, H1_SUM4, rows[i], 4 * ((v >> 48) & 0xFFFF)))
    d = limbs(H1_DF, rows[n - 1]) + 3 * power
    patches += [(T_LOGIC, H1_DF + k, rows[n - 1], (d >> (16 * k)) & 0xFFFF) for k in range(4)]
    patches.append((T_LOGIC, H1_DF2, rows[n - 1], 2 * ((d >> 48) & 0xFFFF)))
    bad, bad_out = _cheat(oracle, pub, blob2, patches=patches)
    _both_reject(oracle, bad, pub, kind, n, bad_out)


def test_fewer_enabled_validators_than_the_header_commits_to(oracle, honest):
    """Dropping a non-signer from the total voting power (nb_enabled - 1) changes the validator-set root."""
    c, pub, blob, kind, proof, out = honest
    b = bytearray(blob)
    nb = int.from_bytes(b[12:16], "little")
    b[12:16] = (nb - 1).to_bytes(4, "little")
    bad, bad_out = _cheat(oracle, pub, bytes(b))
    assert _both_reject(oracle, bad, pub, kind, c["n_max"], bad_out) == 200


def test_wrong_height_and_wrong_chain(oracle, honest):
    import tendermintx_b200 as tmx

    c, pub, blob, kind, proof, out = honest
    # the public target height differs from the height in the header
    p2 = pub[:40] + (int.from_bytes(pub[40:], "big") + 1).to_bytes(8, "big")
    bad, bad_out = _cheat(oracle, p2, blob)
    _both_reject(oracle, bad, p2, kind, c["n_max"], bad_out)
    # an honest proof for mocha-4 is not a proof for a circuit of another chain
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(kind, c["n_max"], tmx.CelestiaConfig, proof.tobytes(), pub, out)


def test_skip_distance_is_checked_by_the_verifier(oracle, honest):
    """verify_skip_distance depends on public data only: the verifier evaluates it itself [REF verify.rs:508-526]."""
    import tendermintx_b200 as tmx

    c, pub, blob, kind, proof, out = honest
    with pytest.raises(tmx.TmxError):
        tmx.verify_proof(kind, c["n_max"], tmx.TendermintConfig(b"mocha-4", skip_max=50), proof.tobytes(), pub, out)
