"""Pins the CPU oracle's proving primitives to plonky2's known answers."""
import numpy as np

P = 2**64 - 2**32 + 1

KAT = {
    "zeros": ([0] * 12, "3c18a9786cb0b359 c4055e3364a246c3 7953db0ab48808f4 c71603f33a1144ca d7709673896996dc 46a84e87642f44ed d032648251ee0b3c 1c687363b207df62 df8565563e8045fe 40f5b37ff4254dae d070f637b431067c 1792b1c4342109d7"),
    "range": (list(range(12)), "d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2 414754bfebd051f0 cb1f8980294a023f 6eb2a9e4d54a9d0f 1902bc3af467e056 f045d5eafdc6021f e4150f77caaa3be5 c9bfd01d39b50cce 5c0a27fcb0e1459b"),
    "neg_one": ([P - 1] * 12, "be0085cfc57a8357 d95af71847d05c09 cf55a13d33c1c953 95803a74f4530e82 fcd99eb30a135df1 e095905e913a3029 de0392461b42919b 7d3260e24e81d031 10d3d0465d9deaa0 a87571083dfc2a47 e18263681e9958f8 e28e96f1ae5e60d3"),
}
NTT8 = [28, 18445622567621360637, 18445618169507741693, 1130298020461564, 18446744069414584317,
        18445613771394122749, 1125899906842620, 1121501793223676]


def test_round_constants(oracle):
    rc = oracle.round_constants()
    first = [0xB585F766F2144405, 0x7746A55F43921AD7, 0xB2FB0D31CEE799B4, 0x0F6760A4803427D7,
             0xE10D666650F4E012, 0x8CAE14CB07D09BF1]
    assert [int(x) for x in rc[:6]] == first
    assert int(rc.max()) < 0xFFFEEAC900011537


def test_poseidon_kats(oracle):
    for name, (inp, out) in KAT.items():
        got = oracle.poseidon_permute(inp)
        assert [f"{int(x):016x}" for x in got] == out.split(), name


def test_goldilocks_constants():
    g = pow(7, (P - 1) >> 32, P)
    assert g == 1753635133440165772
    assert pow(g, 1 << 31, P) == P - 1
    assert pow(8, P - 2, P) == 16140901060737761281


def test_ntt8_kat(oracle):
    assert [int(x) for x in oracle.ntt(list(range(8)))] == NTT8


def test_ntt_vs_naive_and_roundtrip(oracle):
    rng = np.random.default_rng(1)
    for lg in (1, 2, 5, 9):
        a = rng.integers(0, P, size=1 << lg, dtype=np.uint64)
        f = oracle.ntt(a)
        assert np.array_equal(f, oracle.naive_dft(a))
        assert np.array_equal(oracle.ntt(f, inverse=True), a)


def _bitrev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def test_lde_is_evaluation_on_coset(oracle):
    rng = np.random.default_rng(2)
    n, r = 16, 1
    vals = rng.integers(0, P, size=(3, n), dtype=np.uint64)
    lde, coeffs = oracle.lde_batch(vals, r, want_coeffs=True)
    m = n << r
    w = pow(7, (P - 1) // m, P)
    for c in range(3):
        co = [int(x) for x in coeffs[c]]
        for j in range(m):
            x = 7 * pow(w, _bitrev(j, 5), P) % P
            assert int(lde[c, j]) == sum(ci * pow(x, i, P) for i, ci in enumerate(co)) % P
    # even natural-order points are the original values scaled back: P(w_n^i) on the subgroup are the inputs
    wn = pow(7, (P - 1) // n, P)
    co = [int(x) for x in coeffs[0]]
    assert [sum(ci * pow(pow(wn, i, P), k, P) for k, ci in enumerate(co)) % P for i in range(n)] == [int(v) for v in vals[0]]


def test_hash_modes_and_merkle(oracle):
    x = np.arange(1, 20, dtype=np.uint64)
    s = np.zeros(12, dtype=np.uint64)
    s[:8] = x[:8]
    s = oracle.poseidon_permute(s)
    s[:8] = x[8:16]
    s = oracle.poseidon_permute(s)
    s[:3] = x[16:19]
    s = oracle.poseidon_permute(s)
    assert np.array_equal(oracle.hash_no_pad(x), s[:4])
    l, r = np.arange(4, dtype=np.uint64), np.arange(4, 8, dtype=np.uint64)
    st = np.zeros(12, dtype=np.uint64)
    st[:4], st[4:8] = l, r
    assert np.array_equal(oracle.two_to_one(l, r), oracle.poseidon_permute(st)[:4])
    # merkle: 8 rows x 5 cols, cap height 1 -> levels 8,4,2
    cols = np.arange(40, dtype=np.uint64).reshape(5, 8)
    d = oracle.commit_columns(cols, 1)
    assert d.shape == (14, 4)
    leaves = [oracle.hash_no_pad(cols[:, j]) for j in range(8)]
    assert all(np.array_equal(d[j], leaves[j]) for j in range(8))
    assert np.array_equal(d[8], oracle.two_to_one(leaves[0], leaves[1]))
    assert np.array_equal(d[12], oracle.two_to_one(d[8], d[9]))
    # <= 4 columns: hash_or_noop copies
    d2 = oracle.commit_columns(cols[:3], 3)
    assert np.array_equal(d2[5], np.array([5, 13, 21, 0], dtype=np.uint64))
