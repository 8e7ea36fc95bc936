"""ctypes wrapper over oracle/_build/liboracle.so -- the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, by bench.py's ``cpu_baseline`` / ``--impl reference`` legs and
by ``__graft_entry__.smoke()``; never by the product package ``tendermintx_b200``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
P = 2**64 - 2**32 + 1

_LIB = None


def build(force=False):
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
