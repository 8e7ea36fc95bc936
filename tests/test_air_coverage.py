"""Constraint coverage of the tables on the trace domain, evaluated by the oracle's interpreter over the constraint DAG of the
build artefact (the SAME data the product's kernels and verifier compile): (1) an honest pair of first- and second-round
traces satisfies every constraint on every row, cyclically, bus helper columns and running sum included; (2) corrupting ANY
single cell of ANY first-round column is caught by at least one constraint of the two rows that read it -- no column of any
table is unconstrained inside its table.  Whether the constraints are SUFFICIENT for the statement is a different property:
tests/test_cheating_provers.py attacks that."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
P = 2**64 - 2**32 + 1
NAMES = ["sha256", "sha512", "ed25519", "logic", "range"]
BETA, GAMMA = (0x1122334455667788, 0x0102030405060708), (0x0F0E0D0C0B0A0908, 0x7766554433221100)


@pytest.fixture(scope="module", params=["skip_3000_3100_n4", "step_10500_n4_with_dummy"])
def traces(oracle, request):
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        c = {c["name"]: c for c in json.load(f)["cases"]}[request.param]
    blob = bytes.fromhex(c["blob"])
    kind = 1 if c["kind"] == "skip" else 0
    circ = oracle.circuit(kind, c["n_max"], "mocha-4")
    tabs = oracle.all_traces(blob, "mocha-4", public_input=bytes.fromhex(c["input"]))
    aux = [oracle.aux_trace(circ, t, tabs[t], BETA, GAMMA) if tabs[t] is not None else None for t in range(oracle.N_TABLES)]
    return circ, tabs, aux


@pytest.mark.parametrize("table", [0, 1, 2, 3, 4])
def test_honest_trace_satisfies_every_row(oracle, traces, table):
    circ, tabs, aux = traces
    t, (a, total) = tabs[table], aux[table]
    out = oracle.constraints_at_rows(circ, table, t, a, total, BETA, GAMMA, np.arange(t.shape[1]))
    bad = np.nonzero(out.any(axis=1))[0]
    assert bad.size == 0, (NAMES[table], bad[:10])


@pytest.mark.parametrize("table", [0, 1, 2])
def test_single_cell_corruption_is_caught_in_every_column(oracle, traces, table):
    circ, tabs, aux = traces
    t = tabs[table].copy()
    a, total = aux[table]
    C, n = t.shape
    rng = np.random.default_rng(100 + table)
    missed = []
    for c in range(C):
        caught = False
        for r in rng.integers(0, n, 3):  # three rows per column: selectors switch some relations off on some rows
            r = int(r)
            old = t[c, r]
            t[c, r] = (int(old) + 1) % P
            out = oracle.constraints_at_rows(circ, table, t, a, total, BETA, GAMMA, [(r - 1) % n, r])
            t[c, r] = old
            if out.any():
                caught = True
                break
        if not caught:
            missed.append(c)
    assert missed == [], (NAMES[table], missed)
