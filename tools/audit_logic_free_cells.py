"""Audit: which cells of the logic table can an ADAPTIVE prover change without anyone noticing?

For one representative row of every distinct (row type, parameter vector) of the circuit and every one of its 424 cells: change
the cell (+1, or -1 when +1 leaves the cell's range table), rebuild everything a prover would rebuild (the range table's
multiplicities, the helper columns and running sums of both tables), and ask (1) do the AIR constraints of the rows that read
the cell still hold and (2) is the bus still balanced (the totals of the logic and the range table against the honest ones; the
other tables and the verifier's public terms are untouched).  A cell that passes both is FREE: nothing in the proof binds it.
Free cells must be cells the row type does not use (or witness freedom that cannot change the statement); the report groups
them by row type so that this can be read off against logic.cuh.

Usage: python tools/audit_logic_free_cells.py [fixture case] > report   (CPU only; about 17 minutes for the default case)
The committed result is tests/golden/logic_free_cells.json (tests/test_air_coverage.py re-checks a sample of it)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle

P = 2**64 - 2**32 + 1
BETA, GAMMA = (0x1122334455667788, 0x0102030405060708), (0x0F0E0D0C0B0A0908, 0x7766554433221100)
T_LOGIC, T_RANGE = 3, 4
TYPES = ["H1", "H2", "SIG", "SC", "XS", "MUL", "EDIO", "GLOB", "CFE"]
name = sys.argv[1] if len(sys.argv) > 1 else "skip_10000_10500_n4"
c = {x["name"]: x for x in json.load(open(os.path.join(ROOT, "tests", "golden", "fixture_vectors.json")))["cases"]}[name]
pub, blob, kind = bytes.fromhex(c["input"]), bytes.fromhex(c["blob"]), 1 if c["kind"] == "skip" else 0
circ = oracle.circuit(kind, c["n_max"], "mocha-4")
K = circ.table_data(T_LOGIC)[1]  # constant columns [LGK_COLS, rows]
lt, st = oracle.logic_trace(pub, blob, "mocha-4")
assert st == 0
tabs = oracle.all_traces(blob, "mocha-4", logic_trace=lt)
a3, t3 = oracle.aux_trace(circ, T_LOGIC, lt, BETA, GAMMA)
a4, t4 = oracle.aux_trace(circ, T_RANGE, tabs[T_RANGE], BETA, GAMMA)
n = lt.shape[1]
assert not oracle.constraints_at_rows(circ, T_LOGIC, lt, a3, t3, BETA, GAMMA, np.arange(n)).any()
honest = [(int(t3[i]) + int(t4[i])) % P for i in range(2)]
sel = K[50:59]
used = [r for r in range(n) if sel[:, r].any()]
seen, reps = set(), []
for r in used:
    sig = (tuple(int(x) for x in sel[:, r]), tuple(int(x != 0) for x in K[59:, r]), tuple(int(x) for x in K[0:50, r]))
    if sig not in seen:
        seen.add(sig)
        reps.append(r)
print(f"{name}: {len(used)} used rows, {len(reps)} distinct row signatures, {lt.shape[0]} cells each", flush=True)
t0 = time.time()
for r in reps:
    ty = TYPES[int(np.argmax(sel[:, r]))]
    free, ranged = [], 0
    for col in range(lt.shape[0]):
        verdict = None
        for delta in (1, P - 1):
            t = lt.copy()
            t[col, r] = (int(t[col, r]) + delta) % P
            try:
                tb = oracle.all_traces(blob, "mocha-4", logic_trace=t)
            except ValueError:
                verdict = "range"
                continue
            b3, u3 = oracle.aux_trace(circ, T_LOGIC, t, BETA, GAMMA)
            b4, u4 = oracle.aux_trace(circ, T_RANGE, tb[T_RANGE], BETA, GAMMA)
            ok_air = not oracle.constraints_at_rows(circ, T_LOGIC, t, b3, u3, BETA, GAMMA, [(r - 1) % n, r]).any()
            ok_bus = [(int(u3[i]) + int(u4[i])) % P for i in range(2)] == honest
            verdict = "free" if ok_air and ok_bus else "bound"
            break
        if verdict == "free":
            free.append(col)
        ranged += verdict == "range"
    params = [i for i in range(K.shape[0] - 59) if K[59 + i, r]]
    runs, i = [], 0
    while i < len(free):
        j = i
        while j + 1 < len(free) and free[j + 1] == free[j] + 1:
            j += 1
        runs.append(f"{free[i]}" if i == j else f"{free[i]}-{free[j]}")
        i = j + 1
    print(f"row {r:4d} {ty:5s} params {params}: {lt.shape[0] - len(free)} bound cells, free: {' '.join(runs)}; {ranged} cells held only by their range", flush=True)
print(f"done in {time.time() - t0:.0f} s")
