"""CPU-side parity of the witness kernels' per-thread logic: tendermintx_b200/csrc/witness_jobs.cuh + witness.cuh are
header code shared by the CUDA kernels and a host build (tools/hostsim.cpp, TEST TOOL: compiled here with g++ into
tests/_build, never part of libtmx.so).  The tables it fills must equal the oracle's cell for cell, so a logic error in
the SHA-256 / SHA-512 / Ed25519 row expansion shows up in the authoring container, without a GPU."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hostsim():
    # TMX_HOSTSIM_SAN=1 (tools/oracle_sanitize.sh): the kernels' per-thread logic under ASan + UBSan on the host
    san = os.environ.get("TMX_HOSTSIM_SAN") == "1"
    out = os.path.join(HERE, "_build", "libhostsim_san.so" if san else "libhostsim.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tools", "hostsim.cpp")
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if san else ["-O2"]
    subprocess.check_call(["g++", *flags, "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


def _cases():
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}


@pytest.mark.parametrize("name", ["step_10000_n2", "skip_3000_3100_n4", "step_10500_n4_with_dummy"])
def test_kernel_row_logic_equals_oracle_tables(hostsim, oracle, name):
    c = _cases()[name]
    blob = bytes.fromhex(c["blob"])
    want = oracle.build_traces(blob)
    kind = 1 if c["kind"] == "skip" else 0
    dims = oracle.trace_dims(kind, c["n_max"])
    got = [np.zeros((cols, rows), dtype=np.uint64) for rows, cols in dims]
    aux = np.zeros(4096, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = hostsim.hostsim_build_traces(blob, p(got[0]), ctypes.c_size_t(dims[0][0]), p(got[1]), ctypes.c_size_t(dims[1][0]),
                                      p(got[2]), ctypes.c_size_t(dims[2][0]), p(aux))
    assert rc == 0
    for t, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape
        bad = np.argwhere(g != w)
        assert bad.size == 0, f"table {t}: first differing (column, row) {bad[0]} of {len(bad)}"


def test_synthetic_chain_with_absent_signers(hostsim, oracle):
    from oracle import tm_inputs as ti

    src, t, g = ti.synthetic_source(seed=22, n_validators=6, absent_frac=0.3)
    th = ti.header_hash(src.signed_header(t)["header"])
    blob = ti.skip_inputs(src, 8, t, th, g)
    want = oracle.build_traces(blob)
    dims = oracle.trace_dims(1, 8)
    got = [np.zeros((cols, rows), dtype=np.uint64) for rows, cols in dims]
    aux = np.zeros(4096, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert hostsim.hostsim_build_traces(blob, p(got[0]), ctypes.c_size_t(dims[0][0]), p(got[1]), ctypes.c_size_t(dims[1][0]),
                                        p(got[2]), ctypes.c_size_t(dims[2][0]), p(aux)) == 0
    for g_, w in zip(got, want):
        assert np.array_equal(g_, w)


CHAIN = "mocha-4"
BETA, GAMMA = (0x1122334455667788, 0x0102030405060708), (0x0F0E0D0C0B0A0908, 0x7766554433221100)
ALPHA = (0x0123456789ABCDEF, 0x0FEDCBA987654321)


def _u64(x):
    return np.array(x, dtype=np.uint64)


@pytest.mark.parametrize("name", ["skip_3000_3100_n4", "step_10500_n4_with_dummy"])
def test_product_bus_and_air_on_the_lde_coset_equal_oracle(hostsim, oracle, name):
    """The product's per-thread kernel bodies compiled for the host (csrc/stark_rows.cuh: range-lookup histogram, helper
    columns + running sum of the bus, constraint quotient through the compiled AIR templates) against the oracle, which
    INTERPRETS the constraint DAG of the build artefact with its own bus / quotient code: the range table's multiplicities,
    every second-round column and the quotient values at every point of the LDE coset must agree.  Coset points are generic,
    so every term of every constraint contributes a non-zero value."""
    c = _cases()[name]
    blob = bytes.fromhex(c["blob"])
    kind = 1 if c["kind"] == "skip" else 0
    circ = oracle.circuit(kind, c["n_max"], CHAIN)
    shapes = circ.table_shapes()
    tabs = oracle.all_traces(blob, CHAIN, public_input=bytes.fromhex(c["input"]))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    beta, gamma, alpha = _u64(BETA), _u64(GAMMA), _u64(ALPHA)
    hostsim.hostsim_set_plan(p(np.ascontiguousarray(circ.table_data(oracle.T_LOGIC)[1])), ctypes.c_size_t(1 << shapes[oracle.T_LOGIC][1]))
    H16, H11, H8 = 1 << 16, (1 << 16) + (1 << 11), (1 << 16) + (1 << 11) + (1 << 8)
    hist = np.zeros(H8 + 2 + 1, dtype=np.uint32)
    for table, t in enumerate(tabs):
        if t is None:
            continue
        _, log_n, C, Kc, n_per, P, H, _ = shapes[table]
        n = 1 << log_n
        per, cst = circ.table_data(table)
        t = np.ascontiguousarray(t)
        args = (ctypes.c_uint32(kind), ctypes.c_uint32(c["n_max"]), ctypes.c_int(table))
        if table != oracle.T_RANGE:
            hostsim.hostsim_bus_count(*args, p(t), p(cst), p(per), ctypes.c_size_t(n), ctypes.c_size_t(P), p(hist))
        # second-round trace
        want_aux, want_total = oracle.aux_trace(circ, table, t, BETA, GAMMA)
        got_aux, got_total = np.zeros_like(want_aux), np.zeros(2, dtype=np.uint64)
        hostsim.hostsim_bus_aux(*args, ctypes.c_int(H), p(t), p(cst), p(per), ctypes.c_size_t(n), ctypes.c_size_t(P), p(beta), p(gamma),
                                p(got_aux), p(got_total))
        assert np.array_equal(got_total, want_total), (name, table)
        assert np.array_equal(got_aux, want_aux), (name, table, np.argwhere(got_aux != want_aux)[:3])
        # quotient
        lde_m, lde_a, want = oracle.quotient(circ, table, t, want_aux, want_total, BETA, GAMMA, ALPHA)
        lde_k = oracle.lde_batch(cst, 1) if Kc else np.zeros((1, 2 * n), dtype=np.uint64)
        got = np.zeros((2, 2 * n), dtype=np.uint64)
        hostsim.hostsim_quotient(*args, p(lde_m), p(lde_k), p(lde_a), p(per), ctypes.c_int(n_per), ctypes.c_size_t(P), ctypes.c_size_t(n),
                                 p(want_total), p(beta), p(gamma), p(alpha), p(got))
        assert want.any(axis=1).all(), table
        assert np.array_equal(got, want), (name, table, int((got != want).sum()))
    # the range table's multiplicity columns are the histogram
    rg = tabs[oracle.T_RANGE]
    assert hist[-1] == 0
    assert np.array_equal(rg[0], hist[:H16].astype(np.uint64))
    assert np.array_equal(rg[1][:1 << 11], hist[H16:H11].astype(np.uint64)) and not rg[1][1 << 11:].any()
    assert np.array_equal(rg[2][:1 << 8], hist[H11:H8].astype(np.uint64)) and not rg[2][1 << 8:].any()
    assert np.array_equal(rg[3][:2], hist[H8:H8 + 2].astype(np.uint64)) and not rg[3][2:].any()
