"""GPU parity of the bus protocol's kernels through the C ABI vs the CPU oracle, bit-exact: range-lookup histogram
(tmx_bus_count), second commitment round (tmx_bus_aux: helper columns + running sum), constraint quotient (tmx_quotient: compiled
AIR templates on the GPU vs the oracle's interpreter over the artefact's DAG), the SHA-512 table alone (tmx_sha512_trace) and one
FRI fold (tmx_fri_fold) -- so that a regression in one of them is localised instead of showing up only as differing proof bytes."""
import json
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
P = 2**64 - 2**32 + 1
BETA, GAMMA = (0x1122334455667788, 0x0102030405060708), (0x0F0E0D0C0B0A0908, 0x7766554433221100)
ALPHA = (0x0123456789ABCDEF, 0x0FEDCBA987654321)


def _cases():
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def _host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("name", ["skip_3000_3100_n4", "step_10500_n4_with_dummy", "skip_10000_10500_n32"])
def test_bus_rounds_and_quotient_equal_oracle(ctx, oracle, name):
    import tendermintx_b200 as tmx

    c = _cases()[name]
    blob = bytes.fromhex(c["blob"])
    kind, n_max = struct.unpack_from("<II", blob, 4)
    circ = oracle.circuit(kind, n_max, "mocha-4")
    shapes = circ.table_shapes()
    tabs = oracle.all_traces(blob, "mocha-4", public_input=bytes.fromhex(c["input"]))
    circuit = tmx.Circuit.build(ctx, kind, n_max, tmx.Mocha4Config)
    assert list(circuit.digest()) == [int(x) for x in circ.digest()]
    H16, H11, H8 = 1 << 16, (1 << 16) + (1 << 11), (1 << 16) + (1 << 11) + (1 << 8)
    hist_total = np.zeros(H8 + 2, dtype=np.uint64)
    for table, t in enumerate(tabs):
        if t is None:
            assert circuit.table_shape(table)[0] == 0
            continue
        rows, cols, kc, a_cols = circuit.table_shape(table)
        assert (cols, rows) == t.shape and kc == shapes[table][3] and a_cols == 2 * (shapes[table][6] + 1)
        d_t = _dev(t)
        if table != oracle.T_RANGE:
            hist, bad = circuit.bus_count(table, d_t)
            assert not bad
            hist_total += hist.cpu().numpy().astype(np.uint64)
        want_aux, want_total = oracle.aux_trace(circ, table, t, BETA, GAMMA)
        aux, total = circuit.bus_aux(table, d_t, BETA, GAMMA)
        assert tuple(int(x) for x in want_total) == total, (name, table)
        got_aux = _host(aux)
        assert np.array_equal(got_aux, want_aux), (name, table, np.argwhere(got_aux != want_aux)[:3])
        lde_m, lde_a, want_q = oracle.quotient(circ, table, t, want_aux, want_total, BETA, GAMMA, ALPHA)
        q = circuit.quotient(table, _dev(lde_m), _dev(lde_a), want_total, BETA, GAMMA, ALPHA)
        assert want_q.any(axis=1).all()
        assert np.array_equal(_host(q), want_q), (name, table)
    rg = tabs[oracle.T_RANGE]
    assert np.array_equal(rg[0], hist_total[:H16])
    assert np.array_equal(rg[1][:1 << 11], hist_total[H16:H11])
    assert np.array_equal(rg[2][:1 << 8], hist_total[H11:H8])
    assert np.array_equal(rg[3][:2], hist_total[H8:])
    circuit.close()


def test_sha512_table_alone_equals_oracle(ctx, oracle):
    c = _cases()["skip_10000_10500_n4"]
    blob = bytes.fromhex(c["blob"])
    want = oracle.build_traces(blob)[1]
    got = _host(ctx.sha512_trace(blob, 1, 4))
    assert np.array_equal(got, want)


def test_fri_fold_equals_coefficient_folding(ctx, oracle):
    """One arity-16 fold in evaluation space (the product's kernel) against the definition the oracle prover uses: fold the
    COEFFICIENTS (c'_i = sum_j beta^j c_{16 i + j}) and re-evaluate on the coset shift^16 * <w^16>."""
    import torch

    rng = np.random.default_rng(5)
    log_cosets, shift = 6, 7
    n = 16 << log_cosets
    lg = log_cosets + 4
    coef = rng.integers(0, P, size=(2, n), dtype=np.uint64, endpoint=False)  # two component polynomials of an extension-valued polynomial
    beta = (int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64)))

    def ext_mul(a, b):
        return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)

    def evals_on_coset(c, sh):  # natural order values of both components on sh * <w_len>
        out = []
        for comp in c:
            s, scaled = 1, []
            for x in comp:
                scaled.append(int(x) * s % P)
                s = s * sh % P
            out.append(oracle.ntt(np.array(scaled, dtype=np.uint64)))
        return out

    def bitrev(i, bits):
        return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0

    ev = evals_on_coset(coef, shift)
    leaves = np.zeros((n, 2), dtype=np.uint64)
    for p in range(n):
        leaves[p] = (ev[0][bitrev(p, lg)], ev[1][bitrev(p, lg)])
    got = _host(ctx.fri_fold(_dev(leaves), log_cosets, shift, beta))
    folded = [[0] * (n // 16), [0] * (n // 16)]
    for i in range(n // 16):
        acc = (0, 0)
        for j in range(15, -1, -1):
            acc = ext_mul(acc, beta)
            acc = ((acc[0] + int(coef[0][16 * i + j])) % P, (acc[1] + int(coef[1][16 * i + j])) % P)
        folded[0][i], folded[1][i] = acc
    ev2 = evals_on_coset(folded, pow(shift, 16, P))
    want = np.zeros((n // 16, 2), dtype=np.uint64)
    for p in range(n // 16):
        want[p] = (ev2[0][bitrev(p, log_cosets)], ev2[1][bitrev(p, log_cosets)])
    assert np.array_equal(got, want)
