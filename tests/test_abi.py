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


def test_factored_ed25519_constraint_fold_equals_literal_fold_on_host():
    """K5 fast path: the factored evaluation of the Ed25519 table's constraint combination is the same field element
    as the literal Horner fold of air_ed25519(), on random cells (the identity is polynomial, not witness-dependent),
    on small 16-bit cells, and for 0 / 1 / random values of the three periodic columns."""
    import ctypes

    import numpy as np

    import tendermintx_b200 as tmx

    P = 2**64 - 2**32 + 1
    ED_COLS = 945
    rng = np.random.default_rng(11)
    lib = tmx.lib()
    for trial in range(6):
        hi = P if trial % 2 == 0 else 1 << 16
        l = rng.integers(0, hi, size=ED_COLS, dtype=np.uint64)
        n = rng.integers(0, hi, size=ED_COLS, dtype=np.uint64)
        per = [np.array([0, 1, 0], dtype=np.uint64), np.array([1, 0, 1], dtype=np.uint64),
               rng.integers(0, P, size=3, dtype=np.uint64)][trial % 3]  # {not_block_end, first row of [s]B, first row of [h]A}
        alpha = rng.integers(0, P, size=2, dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        assert lib.tmx_host_air_ed25519(vp(l), vp(n), vp(per), vp(alpha), vp(out)) == 0
        assert out[0] == out[2] and out[1] == out[3], (trial, out)
        assert out[0] != 0
