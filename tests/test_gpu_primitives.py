"""GPU parity: K1 (NTT / LDE) and K2 (Poseidon Merkle) through the C ABI vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

P = 2**64 - 2**32 + 1
pytestmark = pytest.mark.gpu


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def _host(t):
    return t.cpu().numpy().view(np.uint64)


def _rand(rng, shape):
    a = rng.integers(0, P, size=shape, dtype=np.uint64)
    # sprinkle edge values
    flat = a.reshape(-1)
    flat[:: max(1, flat.size // 7)] = P - 1
    flat[1:: max(1, flat.size // 5)] = 0
    return a


def test_poseidon_kats_gpu(ctx, oracle):
    from test_oracle_primitives import KAT

    states = np.array([k[0] for k in KAT.values()], dtype=np.uint64)
    got = _host(ctx.poseidon_permute(_dev(states)))
    for row, (_, out) in zip(got, KAT.values()):
        assert [f"{int(x):016x}" for x in row] == out.split()
    rng = np.random.default_rng(0)
    s = _rand(rng, (1000, 12))
    got = _host(ctx.poseidon_permute(_dev(s)))
    for i in range(0, 1000, 37):
        assert np.array_equal(got[i], oracle.poseidon_permute(s[i]))


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 16, 17])
def test_ntt_matches_oracle(ctx, oracle, log_n):
    rng = np.random.default_rng(log_n)
    n_cols = 3 if log_n > 12 else 19
    a = _rand(rng, (n_cols, 1 << log_n))
    fwd = _host(ctx.ntt(_dev(a), log_n))
    for c in range(n_cols):
        assert np.array_equal(fwd[c], oracle.ntt(a[c])), (log_n, c)
    back = _host(ctx.ntt(_dev(fwd), log_n, inverse=True))
    assert np.array_equal(back, a)


def test_ntt8_kat_gpu(ctx):
    from test_oracle_primitives import NTT8

    got = _host(ctx.ntt(_dev(np.arange(8, dtype=np.uint64).reshape(1, 8)), 3))
    assert [int(x) for x in got[0]] == NTT8


@pytest.mark.parametrize("log_n,rate_bits,n_cols", [(3, 1, 5), (6, 1, 20), (10, 1, 7), (12, 1, 33), (12, 3, 4), (14, 2, 3), (16, 1, 5)])
def test_lde_matches_oracle(ctx, oracle, log_n, rate_bits, n_cols):
    import torch

    rng = np.random.default_rng(100 + log_n)
    vals = _rand(rng, (n_cols, 1 << log_n))
    want, coeffs = oracle.lde_batch(vals, rate_bits, want_coeffs=True)
    d_coeffs = torch.empty((n_cols, 1 << log_n), dtype=torch.int64, device="cuda")
    got = _host(ctx.lde(_dev(vals), log_n, rate_bits, coeffs=d_coeffs))
    assert np.array_equal(got, want)
    # coefficient output is coset-scaled: c_i * 7^i
    scaled = _host(d_coeffs)
    pw = 1
    for i in (0, 1, 2, (1 << log_n) - 1):
        pw = pow(7, i, P)
        assert int(scaled[0, i]) == int(coeffs[0, i]) * pw % P
    # without the optional coefficient buffer
    got2 = _host(ctx.lde(_dev(vals), log_n, rate_bits))
    assert np.array_equal(got2, want)


@pytest.mark.parametrize("log_rows,n_cols,cap", [(3, 5, 1), (6, 3, 4), (6, 4, 0), (10, 8, 4), (10, 9, 4), (12, 135, 4), (13, 21, 4), (4, 12, 7)])
def test_poseidon_merkle_matches_oracle(ctx, oracle, log_rows, n_cols, cap):
    rng = np.random.default_rng(200 + log_rows + n_cols)
    cols = _rand(rng, (n_cols, 1 << log_rows))
    want = oracle.commit_columns(cols, cap)
    got = _host(ctx.poseidon_merkle(_dev(cols), log_rows, cap))
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_lde_linearity_full_size(ctx):
    """Size-independent property at a BASELINE-scale shape: LDE(a) + LDE(b) == LDE(a + b) (mod p)."""
    import torch

    rng = np.random.default_rng(7)
    log_n, n_cols = 16, 64
    a = rng.integers(0, P, size=(n_cols, 1 << log_n), dtype=np.uint64)
    b = rng.integers(0, P, size=(n_cols, 1 << log_n), dtype=np.uint64)
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    la, lb, ls = (_host(ctx.lde(_dev(x), log_n, 1)) for x in (a, b, s))
    assert np.array_equal(((la.astype(object) + lb.astype(object)) % P).astype(np.uint64), ls)
