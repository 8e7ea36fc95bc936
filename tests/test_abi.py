"""CPU-side checks of the drop-in boundary: libtmx.so loads and exports every symbol include/tmx.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tmx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tmx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tendermintx_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tmx.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_no_device_fails_loudly():
    import torch
    import pytest
    import tendermintx_b200 as tmx

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tmx.TmxError) as e:
        tmx.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the product package may include, import, link or load it."""
    pkg = os.path.join(ROOT, "tendermintx_b200")
    bad = re.compile(r"liboracle|^\s*(import|from)\s+oracle\b|#\s*include\s*[\"<][^\">]*oracle[^\">]*[\">]|-loracle|dlopen", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                m = bad.search(txt)
                assert m is None, (os.path.join(dirpath, f), m.group(0))


def test_device_poseidon_fast_path_matches_plain_and_oracle_on_host():
    """The kernels' permutation arithmetic (multiplier-free linear layer, unreduced lanes) compiled for the host
    equals the plain formulation and the oracle on random, boundary and all-ones states."""
    import ctypes

    import numpy as np

    import oracle
    import tendermintx_b200 as tmx

    P = 2**64 - 2**32 + 1
    rng = np.random.default_rng(7)
    states = [rng.integers(0, P, size=12, dtype=np.uint64) for _ in range(64)]
    states += [np.zeros(12, dtype=np.uint64), np.full(12, P - 1, dtype=np.uint64), np.arange(12, dtype=np.uint64),
               np.full(12, 2**32 - 1, dtype=np.uint64), np.full(12, 2**32, dtype=np.uint64)]
    a = np.stack(states).copy()
    b = a.copy()
    lib = tmx.lib()
    assert lib.tmx_host_poseidon_permute(a.ctypes.data_as(ctypes.c_void_p), len(states), 0) == 0
    assert lib.tmx_host_poseidon_permute(b.ctypes.data_as(ctypes.c_void_p), len(states), 1) == 0
    assert np.array_equal(a, b)
    c = np.stack(states).copy()
    assert lib.tmx_host_poseidon_permute(c.ctypes.data_as(ctypes.c_void_p), len(states), 2) == 0  # host transcript formulation
    assert np.array_equal(a, c)
    for i, s in enumerate(states):
        assert np.array_equal(b[i], oracle.poseidon_permute(s))


def test_build_artefact_is_reproduced_by_the_oracle(oracle):
    """One definition, two evaluators: the product's verifier evaluates the AIR templates compiled for the extension field, the
    build artefact carries the same constraints as a DAG (obtained by running the templates on symbolic values) which the oracle
    interprets.  A proof made by the oracle from the DAG must satisfy the compiled constraints (tests/test_prove_cpu.py); here
    the artefact itself is checked: it parses, its digest is reproduced by the oracle's own hashing and its constant-column
    caps by the oracle's own NTT / Poseidon code (oracle.Circuit raises otherwise), and shapes agree with the product's ABI."""
    import tendermintx_b200 as tmx

    for kind, n_max, chain in [(0, 2, "mocha-4"), (1, 4, "mocha-4"), (1, 16, "celestia")]:
        circ = oracle.circuit(kind, n_max, chain)
        shapes = circ.table_shapes()
        dims = tmx.Context.trace_dims(kind, n_max)
        for t in range(3):
            assert shapes[t][0] == 1 and (1 << shapes[t][1], shapes[t][2]) == tuple(dims[t])
        assert shapes[oracle.T_RANGE][:3] == (1, 16, 4) and shapes[oracle.T_LOGIC][0] == 1
        # every table declares at least one bus interaction and the Ed25519 table range-checks all 14 x 63 gadget cells
        assert shapes[oracle.T_ED][6] == 14 * 63 // 2 + 2
