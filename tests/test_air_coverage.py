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


def test_logic_table_cells_are_bound_against_an_adaptive_prover(oracle):
    """The logic table needs a stronger check than the one-bad-cell test above, because an adaptive prover recomputes the range
    table's multiplicities and every helper column after changing a cell.  tools/audit_logic_free_cells.py does that for all
    424 cells of one row per (row type, parameter vector): a cell is FREE when the AIR still holds and the bus still balances.
    The committed result (tests/golden/logic_free_cells.json; every free cell is one the row type does not use, or a byte
    beyond the message length) is re-checked here on a sample: bound cells must stay bound, free cells are listed."""
    with open(os.path.join(HERE, "golden", "logic_free_cells.json")) as f:
        audit = json.load(f)
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        c = {c["name"]: c for c in json.load(f)["cases"]}[audit["case"]]
    pub, blob, kind = bytes.fromhex(c["input"]), bytes.fromhex(c["blob"]), 1
    circ = oracle.circuit(kind, c["n_max"], audit["chain"])
    lt, st = oracle.logic_trace(pub, blob, audit["chain"])
    assert st == 0
    n = lt.shape[1]

    def totals(trace):
        tabs = oracle.all_traces(blob, audit["chain"], logic_trace=trace)
        a3, t3 = oracle.aux_trace(circ, 3, trace, BETA, GAMMA)
        a4, t4 = oracle.aux_trace(circ, 4, tabs[4], BETA, GAMMA)
        return a3, t3, [(int(t3[i]) + int(t4[i])) % P for i in range(2)]

    _, _, honest = totals(lt)
    rng = np.random.default_rng(7)
    checked = 0
    for r, info in audit["rows"].items():
        r, free = int(r), set(info["free"])
        bound = [col for col in range(lt.shape[0]) if col not in free]
        for col in rng.choice(bound, 2, replace=False):
            verdicts = []
            for delta in (1, P - 1):
                t = lt.copy()
                t[col, r] = (int(t[col, r]) + delta) % P
                try:
                    a3, t3, tot = totals(t)
                except ValueError:
                    verdicts.append("range")
                    continue
                ok = not oracle.constraints_at_rows(circ, 3, t, a3, t3, BETA, GAMMA, [(r - 1) % n, r]).any() and tot == honest
                verdicts.append("free" if ok else "bound")
                break
            assert "free" not in verdicts, (info["type"], r, int(col))
            checked += 1
    assert checked >= 50
