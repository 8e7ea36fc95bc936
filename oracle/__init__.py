"""ctypes wrapper over oracle/_build/liboracle.so -- the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, by bench.py's ``cpu_baseline`` / ``--impl reference`` legs and
by ``__graft_entry__.smoke()``; never by the product package ``tendermintx_b200``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TMX_ORACLE_LIB: run the same tests against another build of the same sources (the ASan / UBSan one, `make san`)
LIB_PATH = os.environ.get("TMX_ORACLE_LIB") or os.path.join(_HERE, "_build", "liboracle.so")
P = 2**64 - 2**32 + 1

_LIB = None


def build(force=False):
    if os.environ.get("TMX_ORACLE_LIB"):
        return LIB_PATH
    if force or not os.path.exists(LIB_PATH) or _stale():
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def _stale():
    t = os.path.getmtime(LIB_PATH)
    for f in os.listdir(_HERE):
        if f.endswith((".c", ".h", ".inc")) and os.path.getmtime(os.path.join(_HERE, f)) > t:
            return True
    return False


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(LIB_PATH)
        _LIB.poseidon_round_constants.restype = ctypes.POINTER(ctypes.c_uint64)
    return _LIB


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def round_constants():
    rc = lib().poseidon_round_constants()
    return np.array([rc[i] for i in range(360)], dtype=np.uint64)


def poseidon_permute(state):
    s, p = _u64(np.array(state, dtype=np.uint64).copy())
    lib().poseidon_permute(p)
    return s


def hash_no_pad(x):
    x, p = _u64(x)
    out = np.zeros(4, dtype=np.uint64)
    lib().poseidon_hash_no_pad(p, ctypes.c_size_t(x.size), out.ctypes.data_as(ctypes.c_void_p))
    return out


def two_to_one(l, r):
    l, pl = _u64(l)
    r, pr = _u64(r)
    out = np.zeros(4, dtype=np.uint64)
    lib().poseidon_two_to_one(pl, pr, out.ctypes.data_as(ctypes.c_void_p))
    return out


def ntt(a, inverse=False):
    a, p = _u64(np.array(a, dtype=np.uint64).copy())
    (lib().ntt_inverse if inverse else lib().ntt_forward)(p, ctypes.c_size_t(a.size))
    return a


def naive_dft(a):
    a, p = _u64(a)
    out = np.zeros_like(a)
    lib().ntt_naive_dft(p, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.size))
    return out


def lde_batch(values, rate_bits, want_coeffs=False):
    """values: [n_cols, n] uint64 (row c = column c).  Returns [n_cols, n << rate_bits] in bit-reversed
    row order (plonky2 PolynomialBatch leaf order), and optionally the plain coefficients."""
    values, pv = _u64(values)
    n_cols, n = values.shape
    out = np.zeros((n_cols, n << rate_bits), dtype=np.uint64)
    coeffs = np.zeros((n_cols, n), dtype=np.uint64) if want_coeffs else None
    lib().ntt_lde_batch(pv, ctypes.c_size_t(n_cols), ctypes.c_size_t(n), ctypes.c_uint(rate_bits),
                        out.ctypes.data_as(ctypes.c_void_p),
                        coeffs.ctypes.data_as(ctypes.c_void_p) if want_coeffs else ctypes.c_void_p(0))
    return (out, coeffs) if want_coeffs else out


class _MerkleTree(ctypes.Structure):
    _fields_ = [("n_leaves", ctypes.c_size_t), ("leaf_len", ctypes.c_size_t), ("cap_height", ctypes.c_uint),
                ("n_levels", ctypes.c_uint), ("digests", ctypes.POINTER(ctypes.c_uint64)),
                ("cap", ctypes.POINTER(ctypes.c_uint64))]


def commit_columns(cols, cap_height):
    """cols: [n_cols, n_rows] uint64.  Returns all digests [(count), 4] with levels concatenated (leaf level
    first, cap level last) -- the same layout tmx_poseidon_merkle writes."""
    cols, pc = _u64(cols)
    n_cols, n_rows = cols.shape
    t = _MerkleTree()
    lib().commit_columns(ctypes.byref(t), pc, ctypes.c_size_t(n_cols), ctypes.c_size_t(n_rows), ctypes.c_uint(cap_height))
    total = sum(n_rows >> l for l in range(t.n_levels))
    out = np.ctypeslib.as_array(t.digests, shape=(total * 4,)).copy().reshape(total, 4)
    lib().merkle_free(ctypes.byref(t))
    return out


# ---------------------------------------------------------------- witness side (oracle_w.h)
CHECK_NAMES = ["OK", "SKIP_DISTANCE", "TRUSTED_HEADER_PROOF", "TRUSTED_VALHASH", "TRUSTED_THRESHOLD", "SIGNATURE",
               "VALHASH", "VALHASH_PROOF", "THRESHOLD", "SIGN_BYTES", "CHAIN_ID", "HEIGHT", "LAST_BLOCK_ID",
               "NEXT_VALHASH", "VOTING_OVERFLOW", "VARINT_SIGN", "ROUND_SIGN", "INPUT"]


def _buf(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(bytes(b)) if len(b) else (ctypes.c_uint8 * 1)()


def sha256(msg):
    out = (ctypes.c_uint8 * 32)()
    lib().sha256(_buf(msg), ctypes.c_size_t(len(msg)), out)
    return bytes(out)


def sha512(msg):
    out = (ctypes.c_uint8 * 64)()
    lib().sha512(_buf(msg), ctypes.c_size_t(len(msg)), out)
    return bytes(out)


def ed25519_verify(pk, sig, msg):
    return bool(lib().ed25519_verify(_buf(pk), _buf(sig), _buf(msg), ctypes.c_size_t(len(msg))))


def sc_reduce512(x):
    out = (ctypes.c_uint8 * 32)()
    lib().sc_reduce512(out, _buf(x))
    return bytes(out)


def marshal_int64_varint(v):
    out = (ctypes.c_uint8 * 9)()
    lib().tm_marshal_int64_varint(ctypes.c_uint64(v), out)
    return bytes(out)


def marshal_validator(pk, power):
    out = (ctypes.c_uint8 * 46)()
    lib().tm_marshal_validator.restype = ctypes.c_size_t
    n = lib().tm_marshal_validator(_buf(pk), ctypes.c_uint64(power), out)
    return bytes(out), n


def root_from_hashed_leaves(leaves, nb_enabled):
    out = (ctypes.c_uint8 * 32)()
    lib().tm_root_from_hashed_leaves(_buf(b"".join(leaves)), ctypes.c_size_t(len(leaves)), ctypes.c_size_t(nb_enabled), out)
    return bytes(out)


def voting_threshold(power, in_group, nb_enabled, num, den):
    n = len(power)
    p = (ctypes.c_uint64 * n)(*power)
    g = (ctypes.c_uint8 * n)(*[1 if x else 0 for x in in_group])
    res = ctypes.c_int(0)
    rc = lib().tm_voting_threshold(p, g, ctypes.c_size_t(n), ctypes.c_size_t(nb_enabled), ctypes.c_uint64(num),
                                   ctypes.c_uint64(den), ctypes.byref(res))
    return rc, bool(res.value)


def verify_circuit(public_input, blob, chain_id, skip_max=100800):
    """verify_skip / verify_step predicate.  Returns (check_name, output32 or None)."""
    out = (ctypes.c_uint8 * 32)()
    cid = chain_id.encode() if isinstance(chain_id, str) else chain_id
    rc = lib().tm_verify_circuit(_buf(public_input), ctypes.c_size_t(len(public_input)), _buf(blob),
                                 ctypes.c_size_t(len(blob)), _buf(cid), ctypes.c_size_t(len(cid)),
                                 ctypes.c_uint64(skip_max), out)
    return CHECK_NAMES[rc], (bytes(out) if rc == 0 else None)


class _Trace(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_size_t), ("n_cols", ctypes.c_size_t), ("data", ctypes.POINTER(ctypes.c_uint64))]


N_TABLES = 5
T_SHA256, T_SHA512, T_ED, T_LOGIC, T_RANGE = range(5)


def trace_dims(kind, n_max):
    d = (ctypes.c_size_t * 6)()
    lib().tm_trace_dims(ctypes.c_uint32(kind), ctypes.c_uint32(n_max), d)
    return [(d[0], d[1]), (d[2], d[3]), (d[4], d[5])]


def build_traces(blob):
    """Returns [sha256, sha512, ed25519] tables as uint64 arrays of shape [n_cols, n_rows] (column-major)."""
    tr = (_Trace * 3)()
    rc = lib().tm_build_traces(_buf(blob), ctypes.c_size_t(len(blob)), tr)
    if rc != 0:
        raise ValueError(f"tm_build_traces: {CHECK_NAMES[rc]}")
    out = []
    for t in tr:
        a = np.ctypeslib.as_array(t.data, shape=(t.n_cols * t.n_rows,)).copy().reshape(t.n_cols, t.n_rows)
        out.append(a)
    lib().tm_free_traces(tr)
    return out


# ---------------------------------------------------------------- circuits (build artefacts) and proofs
_PRODUCT_LIB = None
_CIRCUITS = {}


def _product_lib():
    """libtmx.so, loaded directly (not through the product's Python package): the oracle takes the circuit DEFINITION -- the
    build artefact, pure data -- from the product's host-side builder and nothing else."""
    global _PRODUCT_LIB
    if _PRODUCT_LIB is None:
        path = os.path.join(os.path.dirname(_HERE), "tendermintx_b200", "libtmx.so")
        _PRODUCT_LIB = ctypes.CDLL(path)
        _PRODUCT_LIB.tmx_circuit_artefact.restype = ctypes.c_size_t
        _PRODUCT_LIB.tmx_circuit_artefact.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64,
                                                      ctypes.c_void_p, ctypes.c_size_t]
        _PRODUCT_LIB.tmx_logic_trace.restype = ctypes.c_size_t
        _PRODUCT_LIB.tmx_logic_trace.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                                 ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]
    return _PRODUCT_LIB


def artefact_words(kind, n_max, chain_id, skip_max=100800):
    cid = chain_id.encode() if isinstance(chain_id, str) else chain_id
    n = _product_lib().tmx_circuit_artefact(kind, n_max, cid, len(cid), skip_max, None, 0)
    if n == 0:
        raise RuntimeError("tmx_circuit_artefact failed")
    w = np.zeros(n, dtype=np.uint64)
    _product_lib().tmx_circuit_artefact(kind, n_max, cid, len(cid), skip_max, w.ctypes.data_as(ctypes.c_void_p), n)
    return w


class Circuit:
    """A parsed build artefact.  Parsing recomputes the digest and re-commits the constant columns with the oracle's own code."""

    def __init__(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        rc = ctypes.c_int(0)
        lib().tm_circuit_load.restype = ctypes.c_void_p
        self.handle = lib().tm_circuit_load(words.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(words.size), ctypes.byref(rc))
        if not self.handle:
            raise ValueError(f"tm_circuit_load: code {rc.value}")
        self.words = words

    def digest(self):
        out = np.zeros(4, dtype=np.uint64)
        lib().tm_circuit_digest(ctypes.c_void_p(self.handle), out.ctypes.data_as(ctypes.c_void_p))
        return out

    def table_shapes(self):
        """[(present, log_n, n_main, n_const, n_per, period, n_helpers, n_constraints)] per table, read from the words."""
        w = self.words
        pos, out = 20, []
        for _ in range(N_TABLES):
            present = int(w[pos]); pos += 1
            if not present:
                out.append((0,) * 8)
                continue
            f = [int(x) for x in w[pos:pos + 10]]; pos += 10
            out.append((1, f[0], f[1], f[2], f[3], f[4], f[5], f[6]))
            pos += f[3] * f[4] + (f[2] << f[0]) + f[7] + 2 * f[8] + f[9]
        return out


    def table_data(self, table):
        """(periodic [n_per, period], constants [n_const, n]) of a table, read from the words."""
        w = self.words
        pos = 20
        for t in range(N_TABLES):
            present = int(w[pos]); pos += 1
            if not present:
                if t == table:
                    return None
                continue
            f = [int(x) for x in w[pos:pos + 10]]; pos += 10
            n_per, n_cst = f[3] * f[4], f[2] << f[0]
            if t == table:
                per = np.array(w[pos:pos + n_per]).reshape(f[3], f[4]) if f[3] else np.zeros((0, f[4]), dtype=np.uint64)
                cst = np.array(w[pos + n_per:pos + n_per + n_cst]).reshape(f[2], 1 << f[0])
                return np.ascontiguousarray(per), np.ascontiguousarray(cst)
            pos += n_per + n_cst + f[7] + 2 * f[8] + f[9]
        return None


def circuit(kind, n_max, chain_id, skip_max=100800):
    key = (kind, n_max, chain_id if isinstance(chain_id, str) else chain_id.decode(), skip_max)
    if key not in _CIRCUITS:
        _CIRCUITS[key] = Circuit(artefact_words(kind, n_max, chain_id, skip_max))
    return _CIRCUITS[key]


def logic_trace(public_input, blob, chain_id, force=True):
    """First-round trace of the logic table, [cols, rows]: the witness of the plain-gate gadgets.  It is produced by the
    product's HOST code (the table is small and has no kernel); the oracle proves with it and checks it against the
    constraint DAG like any other table.  Returns (trace or None when the circuit has no logic table, status)."""
    kind, n_max = _blob_shape(blob)
    cid = chain_id.encode() if isinstance(chain_id, str) else chain_id
    st = ctypes.c_int(0)
    cells = _product_lib().tmx_logic_trace(kind, n_max, cid, len(cid), bytes(public_input), bytes(blob), int(force), None, 0, ctypes.byref(st))
    if cells == 0:
        return None, 0
    sh = circuit(kind, n_max, chain_id).table_shapes()[T_LOGIC]
    out = np.zeros((sh[2], 1 << sh[1]), dtype=np.uint64)
    assert out.size == cells
    _product_lib().tmx_logic_trace(kind, n_max, cid, len(cid), bytes(public_input), bytes(blob), int(force),
                                   out.ctypes.data_as(ctypes.c_void_p), cells, ctypes.byref(st))
    return out, st.value


def _blob_shape(blob):
    kind, n_max = np.frombuffer(bytes(blob[4:12]), dtype=np.uint32)
    return int(kind), int(n_max)


def cheat_next_proof(skip_precheck=False, patches=()):
    """Test hooks for the NEXT prove() on this thread: skip the statement pre-check, and / or add deltas to trace cells after all
    witness generation.  patches: iterable of (table, col, row, delta)."""
    if skip_precheck:
        lib().tm_debug_skip_precheck_next_proof()
    patches = list(patches)
    if patches:
        n = len(patches)
        tabs = (ctypes.c_int * n)(*[p[0] for p in patches])
        cols = (ctypes.c_size_t * n)(*[p[1] for p in patches])
        rows = (ctypes.c_size_t * n)(*[p[2] for p in patches])
        deltas = (ctypes.c_uint64 * n)(*[p[3] % P for p in patches])
        lib().tm_debug_patch_next_proof(ctypes.c_int(n), tabs, cols, rows, deltas)


def prove(public_input, blob, chain_id, skip_max=100800, logic_trace=None):
    """CPU oracle prover.  Returns (status, proof as uint64 array or None, output32 or None)."""
    kind, n_max = _blob_shape(blob)
    c = circuit(kind, n_max, chain_id, skip_max)
    p = ctypes.POINTER(ctypes.c_uint64)()
    n = ctypes.c_size_t(0)
    out = (ctypes.c_uint8 * 32)()
    lt = None
    if logic_trace is None:
        logic_trace, _ = globals()["logic_trace"](public_input, blob, chain_id)
    if logic_trace is not None:
        lt = np.ascontiguousarray(logic_trace, dtype=np.uint64)
    rc = lib().tm_prove(ctypes.c_void_p(c.handle), _buf(public_input), ctypes.c_size_t(len(public_input)), _buf(blob),
                        ctypes.c_size_t(len(blob)), lt.ctypes.data_as(ctypes.c_void_p) if lt is not None else ctypes.c_void_p(0),
                        ctypes.byref(p), ctypes.byref(n), out)
    if rc != 0:
        return CHECK_NAMES[rc], None, None
    proof = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    lib().tm_proof_free(p)
    return "OK", proof, bytes(out)


def verify_proof(proof, public_input, chain_id, kind, n_max, output32, skip_max=100800):
    """Returns 0 when the proof verifies; a non-zero diagnostic code otherwise."""
    c = circuit(kind, n_max, chain_id, skip_max)
    proof = np.ascontiguousarray(proof, dtype=np.uint64)
    return lib().tm_verify_proof(ctypes.c_void_p(c.handle), proof.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(proof.size),
                                 _buf(public_input), ctypes.c_size_t(len(public_input)), _buf(output32))


def all_traces(blob, chain_id, skip_max=100800, logic_trace=None, public_input=None):
    """First-round traces of all tables (None where a table is absent), [n_cols, n_rows] each."""
    kind, n_max = _blob_shape(blob)
    if logic_trace is None and public_input is not None:
        logic_trace, _ = globals()["logic_trace"](public_input, blob, chain_id)
    c = circuit(kind, n_max, chain_id, skip_max)
    shapes = c.table_shapes()
    arrs, ptrs = [], (ctypes.c_void_p * N_TABLES)()
    for t, sh in enumerate(shapes):
        if sh[0]:
            a = np.zeros((sh[2], 1 << sh[1]), dtype=np.uint64)
            arrs.append(a)
            ptrs[t] = a.ctypes.data_as(ctypes.c_void_p)
        else:
            arrs.append(None)
            ptrs[t] = None
    lt = np.ascontiguousarray(logic_trace, dtype=np.uint64) if logic_trace is not None else None
    rc = lib().tm_debug_traces(ctypes.c_void_p(c.handle), _buf(blob), ctypes.c_size_t(len(blob)),
                               lt.ctypes.data_as(ctypes.c_void_p) if lt is not None else ctypes.c_void_p(0), ptrs)
    if rc != 0:
        raise ValueError(f"tm_debug_traces: {CHECK_NAMES[rc]}")
    return arrs


def aux_trace(circ, table, trace, beta, gamma):
    """Second-round trace [2 (H + 1), n] and the table's bus total for given challenges (pairs of u64)."""
    sh = circ.table_shapes()[table]
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    aux = np.zeros((2 * (sh[6] + 1), 1 << sh[1]), dtype=np.uint64)
    total = np.zeros(2, dtype=np.uint64)
    b, g = np.array(beta, dtype=np.uint64), np.array(gamma, dtype=np.uint64)
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    lib().tm_debug_aux(ctypes.c_void_p(circ.handle), ctypes.c_int(table), vp(trace), vp(b), vp(g), vp(aux), vp(total))
    return aux, total


def quotient(circ, table, trace, aux, total, beta, gamma, alpha):
    """LDEs of both traces (bit-reversed rows) and the quotient values on the LDE coset, natural order [2, m]."""
    sh = circ.table_shapes()[table]
    m = 2 << sh[1]
    trace, aux = np.ascontiguousarray(trace, dtype=np.uint64), np.ascontiguousarray(aux, dtype=np.uint64)
    lde_m, lde_a, qv = np.zeros((sh[2], m), dtype=np.uint64), np.zeros((aux.shape[0], m), dtype=np.uint64), np.zeros((2, m), dtype=np.uint64)
    arr = lambda x: np.array(x, dtype=np.uint64)
    t, b, g, a = arr(total), arr(beta), arr(gamma), arr(alpha)
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    lib().tm_debug_quotient(ctypes.c_void_p(circ.handle), ctypes.c_int(table), vp(trace), vp(aux), vp(t), vp(b), vp(g), vp(a), vp(lde_m),
                            vp(lde_a), vp(qv))
    return lde_m, lde_a, qv


def constraints_at_rows(circ, table, trace, aux, total, beta, gamma, rows, alpha=(0x123456789ABCDEF, 0xFEDCBA987654321)):
    """Folded constraint values (two challenges) of the transitions row -> row + 1 (cyclic) on the trace domain, second-round
    columns included; [len(rows), 2] uint64, all zero for a satisfying pair of traces."""
    trace, aux = np.ascontiguousarray(trace, dtype=np.uint64), np.ascontiguousarray(aux, dtype=np.uint64)
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    arr = lambda x: np.array(x, dtype=np.uint64)
    t, b, g, a = arr(total), arr(beta), arr(gamma), arr(alpha)
    out = np.zeros((rows.size, 2), dtype=np.uint64)
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    lib().tm_debug_constraints_at_rows(ctypes.c_void_p(circ.handle), ctypes.c_int(table), vp(trace), vp(aux), vp(t), vp(b), vp(g), vp(a),
                                       vp(rows), ctypes.c_size_t(rows.size), vp(out))
    return out
