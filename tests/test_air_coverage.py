"""Constraint coverage of the three AIRs (oracle/air.inc = tendermintx_b200/csrc/air.cuh, same emission order; their
equality is what the proof-byte parity tests pin).  On the trace domain: (1) an honest trace satisfies every
constraint on every row, cyclically; (2) corrupting ANY single cell of ANY column is caught by at least one constraint
of the two rows that read it -- no column of any table is unconstrained.  (What is NOT enforced yet is the linkage BETWEEN
tables and to the public inputs; that is a different property.)"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
P = 2**64 - 2**32 + 1
NAMES = ["sha256", "sha512", "ed25519"]


@pytest.fixture(scope="module", params=["skip_3000_3100_n4", "step_10500_n4_with_dummy"])
def traces(oracle, request):
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        c = {c["name"]: c for c in json.load(f)["cases"]}[request.param]
    blob = bytes.fromhex(c["blob"])
    return oracle.build_traces(blob), 1 if c["kind"] == "skip" else 0, c["n_max"]


@pytest.mark.parametrize("table", [0, 1, 2])
def test_honest_trace_satisfies_every_row(oracle, traces, table):
    tabs, kind, n_max = traces
    t = tabs[table]
    out = oracle.constraints_at_rows(table, t, np.arange(t.shape[1]), kind, n_max)
    bad = np.nonzero(out.any(axis=1))[0]
    assert bad.size == 0, (NAMES[table], bad[:10])


@pytest.mark.parametrize("table", [0, 1, 2])
def test_single_cell_corruption_is_caught_in_every_column(oracle, traces, table):
    tabs, kind, n_max = traces
    t = tabs[table].copy()
    C, n = t.shape
    rng = np.random.default_rng(100 + table)
    missed = []
    for c in range(C):
        caught = False
        for r in rng.integers(0, n, 3):  # three rows per column: selectors switch some relations off on some rows
            r = int(r)
            old = t[c, r]
            t[c, r] = (int(old) + 1) % P
            out = oracle.constraints_at_rows(table, t, [(r - 1) % n, r], kind, n_max)
            t[c, r] = old
            if out.any():
                caught = True
                break
        if not caught:
            missed.append(c)
    assert missed == [], (NAMES[table], missed)
