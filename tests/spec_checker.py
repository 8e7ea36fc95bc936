"""An independent reading of the witness tables: plain Python written from the SPECIFICATIONS (FIPS 180-4 for SHA-256 / SHA-512,
RFC 8032 and the hwcd twisted-Edwards formulas for Ed25519, hashlib as ground truth), sharing no text with the constraint code
(tendermintx_b200/csrc/air.cuh) or with the oracle's interpreter.  It answers a different question than the constraint checks:
not "does this table satisfy our AIR" but "does every row of this table encode the computation the statement is about".
Used on tables produced by the CPU oracle (tests/test_trace_semantics.py) and by the GPU kernels (tests/test_gpu_witness.py)."""
import hashlib
import re
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def _layout():
    out = {}
    with open(os.path.join(os.path.dirname(HERE), "include", "tmx_trace.h")) as f:
        for m in re.finditer(r"^#define\s+(\w+)\s+(\(?[-\w\s+*()<]+?\)?)\s*(?:/\*.*)?$", f.read(), re.M):
            name, expr = m.group(1), m.group(2)
            try:
                out[name] = int(eval(expr, {}, out))
            except Exception:
                pass
    return out


L = _layout()
M32, M64 = (1 << 32) - 1, (1 << 64) - 1
K256 = [int(x, 16) for x in """428a2f98 71374491 b5c0fbcf e9b5dba5 3956c25b 59f111f1 923f82a4 ab1c5ed5 d807aa98 12835b01 243185be 550c7dc3 72be5d74
80deb1fe 9bdc06a7 c19bf174 e49b69c1 efbe4786 0fc19dc6 240ca1cc 2de92c6f 4a7484aa 5cb0a9dc 76f988da 983e5152 a831c66d b00327c8 bf597fc7 c6e00bf3
d5a79147 06ca6351 14292967 27b70a85 2e1b2138 4d2c6dfc 53380d13 650a7354 766a0abb 81c2c92e 92722c85 a2bfe8a1 a81a664b c24b8b70 c76c51a3 d192e819
d6990624 f40e3585 106aa070 19a4c116 1e376c08 2748774c 34b0bcb5 391c0cb3 4ed8aa4a 5b9cca4f 682e6ff3 748f82ee 78a5636f 84c87814 8cc70208 90befffa
a4506ceb bef9a3f7 c67178f2""".split()]
IV256 = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
IV512 = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1, 0x510e527fade682d1, 0x9b05688c2b3e6c1f,
         0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]


def _primes(n):
    ps, c = [], 2
    while len(ps) < n:
        if all(c % p for p in ps if p * p <= c):
            ps.append(c)
        c += 1
    return ps


def _icbrt(x):
    r = int(round(x ** (1 / 3)))
    while r ** 3 > x:
        r -= 1
    while (r + 1) ** 3 <= x:
        r += 1
    return r


# FIPS 180-4 4.2.3: the first 64 bits of the fractional parts of the cube roots of the first eighty primes
K512 = [(_icbrt(p << 192)) & M64 for p in _primes(80)]
assert K512[0] == 0x428a2f98d728ae22 and K512[79] == 0x6c44198c4a475817


def rotr(x, r, bits):
    return ((x >> r) | (x << (bits - r))) & ((1 << bits) - 1)


def sha256_rounds(cv, block):
    """FIPS 180-4 6.2.2: yields the working variables BEFORE each round and returns (states, W, digest words)."""
    W = [int.from_bytes(block[4 * i:4 * i + 4], "big") for i in range(16)]
    for t in range(16, 64):
        s0 = rotr(W[t - 15], 7, 32) ^ rotr(W[t - 15], 18, 32) ^ (W[t - 15] >> 3)
        s1 = rotr(W[t - 2], 17, 32) ^ rotr(W[t - 2], 19, 32) ^ (W[t - 2] >> 10)
        W.append((W[t - 16] + s0 + W[t - 7] + s1) & M32)
    a, b, c, d, e, f, g, h = cv
    states = []
    for t in range(64):
        states.append((a, b, c, d, e, f, g, h))
        S1 = rotr(e, 6, 32) ^ rotr(e, 11, 32) ^ rotr(e, 25, 32)
        ch = (e & f) ^ (~e & M32 & g)
        t1 = (h + S1 + ch + K256[t] + W[t]) & M32
        S0 = rotr(a, 2, 32) ^ rotr(a, 13, 32) ^ rotr(a, 22, 32)
        mj = (a & b) ^ (a & c) ^ (b & c)
        t2 = (S0 + mj) & M32
        a, b, c, d, e, f, g, h = (t1 + t2) & M32, a, b, c, (d + t1) & M32, e, f, g
    return states, W, [(x + y) & M32 for x, y in zip(cv, (a, b, c, d, e, f, g, h))]


def sha512_rounds(cv, block):
    W = [int.from_bytes(block[8 * i:8 * i + 8], "big") for i in range(16)]
    for t in range(16, 80):
        s0 = rotr(W[t - 15], 1, 64) ^ rotr(W[t - 15], 8, 64) ^ (W[t - 15] >> 7)
        s1 = rotr(W[t - 2], 19, 64) ^ rotr(W[t - 2], 61, 64) ^ (W[t - 2] >> 6)
        W.append((W[t - 16] + s0 + W[t - 7] + s1) & M64)
    a, b, c, d, e, f, g, h = cv
    states = []
    for t in range(80):
        states.append((a, b, c, d, e, f, g, h))
        S1 = rotr(e, 14, 64) ^ rotr(e, 18, 64) ^ rotr(e, 41, 64)
        ch = (e & f) ^ (~e & M64 & g)
        t1 = (h + S1 + ch + K512[t] + W[t]) & M64
        S0 = rotr(a, 28, 64) ^ rotr(a, 34, 64) ^ rotr(a, 39, 64)
        mj = (a & b) ^ (a & c) ^ (b & c)
        a, b, c, d, e, f, g, h = (t1 + S0 + mj) & M64, a, b, c, (d + t1) & M64, e, f, g
    return states, W, [(x + y) & M64 for x, y in zip(cv, (a, b, c, d, e, f, g, h))]


def _bits(tab, col0, row, n):
    return sum(int(tab[col0 + i, row]) << i for i in range(n))


def pad(msg, block):
    ln = 8 if block == 64 else 16
    m = msg + b"\x80"
    m += b"\x00" * ((-len(m) - ln) % block)
    return m + (8 * len(msg)).to_bytes(ln, "big")


def check_sha256_table(tab, messages):
    """messages: the byte strings hashed, in chunk order (each occupies whole chunks).  Every row of every used chunk must show
    the FIPS round state, the schedule window, the chaining value and (last row) the digest; digests must equal hashlib's."""
    chunk = 0
    for msg in messages:
        blocks = pad(msg, 64)
        cv = list(IV256)
        for b in range(len(blocks) // 64):
            states, W, dg = sha256_rounds(cv, blocks[64 * b:64 * b + 64])
            for t in (0, 1, 15, 16, 31, 62, 63):  # spot rows (every row would be slow in pure Python; the set covers all roles)
                row = 64 * chunk + t
                a, bb, c, d, e, f, g, h = states[t]
                got = (_bits(tab, L["S256_A"], row, 32), _bits(tab, L["S256_B"], row, 32), _bits(tab, L["S256_C"], row, 32), int(tab[L["S256_D"], row]),
                       _bits(tab, L["S256_E"], row, 32), _bits(tab, L["S256_F"], row, 32), _bits(tab, L["S256_G"], row, 32), int(tab[L["S256_H"], row]))
                assert got == states[t], ("sha256 state", chunk, t)
                assert int(tab[L["S256_W"] + 15, row]) == W[t], ("sha256 W_t", chunk, t)
                assert [int(tab[L["S256_CV"] + j, row]) for j in range(8)] == cv, ("sha256 cv", chunk, t)
                if t < 63:
                    nxt = states[t + 1]
                    assert _bits(tab, L["S256_AN"], row, 32) == nxt[0] and _bits(tab, L["S256_EN"], row, 32) == nxt[4]
            assert [int(tab[L["S256_W"] + j, 64 * chunk + 15]) for j in range(16)] == W[:16], ("sha256 message words on row 15", chunk)
            assert [int(tab[L["S256_DG"] + j, 64 * chunk + 63]) for j in range(8)] == dg, ("sha256 digest", chunk)
            cv = dg
            chunk += 1
        assert b"".join(x.to_bytes(4, "big") for x in cv) == hashlib.sha256(msg).digest()
    return chunk


def check_sha512_table(tab, messages):
    """messages: per validator slot the message R || A || M[..len]; a slot is two 128-row chunks (80 rounds each + filler)."""
    for slot, msg in enumerate(messages):
        blocks = pad(msg, 128)
        two = len(blocks) == 256
        cv = list(IV512)
        for b in range(len(blocks) // 128):
            states, W, dg = sha512_rounds(cv, blocks[128 * b:128 * b + 128])
            base = 256 * slot + 128 * b
            for t in (0, 15, 16, 40, 79):
                row = base + t
                a, bb, c, d, e, f, g, h = states[t]
                got = (_bits(tab, L["S512_A"], row, 64), _bits(tab, L["S512_B"], row, 64), _bits(tab, L["S512_C"], row, 64),
                       int(tab[L["S512_D"], row]) | int(tab[L["S512_D"] + 1, row]) << 32,
                       _bits(tab, L["S512_E"], row, 64), _bits(tab, L["S512_F"], row, 64), _bits(tab, L["S512_G"], row, 64),
                       int(tab[L["S512_H"], row]) | int(tab[L["S512_H"] + 1, row]) << 32)
                assert got == states[t], ("sha512 state", slot, b, t)
                assert int(tab[L["S512_W"] + 30, row]) | int(tab[L["S512_W"] + 31, row]) << 32 == W[t]
            words15 = [int(tab[L["S512_W"] + 2 * j, base + 15]) | int(tab[L["S512_W"] + 2 * j + 1, base + 15]) << 32 for j in range(16)]
            assert words15 == W[:16], ("sha512 message words on row 15", slot, b)
            dgt = [int(tab[L["S512_DG"] + 2 * j, base + 79]) | int(tab[L["S512_DG"] + 2 * j + 1, base + 79]) << 32 for j in range(8)]
            assert dgt == dg, ("sha512 digest", slot, b)
            assert int(tab[L["S512_TWO"], base]) == int(two)
            cv = dg
        assert b"".join(x.to_bytes(8, "big") for x in cv) == hashlib.sha512(msg).digest()


# ---- Ed25519 from RFC 8032 (affine arithmetic on big integers) ----
P25519 = 2**255 - 19
ELL = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P25519 - 2, P25519)) % P25519
BY = 4 * pow(5, P25519 - 2, P25519) % P25519


def _recover_x(y, sign):
    x2 = (y * y - 1) * pow(D * y * y + 1, P25519 - 2, P25519) % P25519
    x = pow(x2, (P25519 + 3) // 8, P25519)
    if (x * x - x2) % P25519:
        x = x * pow(2, (P25519 - 1) // 4, P25519) % P25519
    assert (x * x - x2) % P25519 == 0
    return P25519 - x if (x & 1) != sign else x


BX = _recover_x(BY, 0)


def ed_add(p, q):  # RFC 8032 5.1.4 affine addition law of -x^2 + y^2 = 1 + d x^2 y^2
    x1, y1 = p
    x2, y2 = q
    k = D * x1 * x2 * y1 * y2 % P25519
    x3 = (x1 * y2 + x2 * y1) * pow(1 + k, P25519 - 2, P25519) % P25519
    y3 = (y1 * y2 + x1 * x2) * pow(1 - k, P25519 - 2, P25519) % P25519
    return x3, y3


def ed_mul(s, p):
    r = (0, 1)
    while s:
        if s & 1:
            r = ed_add(r, p)
        p = ed_add(p, p)
        s >>= 1
    return r


def decompress(b):
    y = int.from_bytes(b, "little") & ((1 << 255) - 1)
    return _recover_x(y, b[31] >> 7), y


def _fe(tab, col0, row):
    return sum(int(tab[col0 + i, row]) << (16 * i) for i in range(16))


def check_ed25519_table(tab, triples):
    """triples: per validator slot (public key, signature, message) as verified (dummy triple for unsigned slots).  Every row's
    accumulator must be the projective image of the running value of the joint evaluation [s]B + [h](-A), every multiplication
    gadget an exact integer identity U V = c + q p with all cells in range, and the slot's result must equal R."""
    ED_MUL, STRIDE = L["ED_MUL"], L["ED_MUL_STRIDE"]
    for slot, (pk, sig, msg) in enumerate(triples):
        A, R = decompress(pk), decompress(sig[:32])
        s = int.from_bytes(sig[32:], "little")
        h = int.from_bytes(hashlib.sha512(sig[:32] + pk + msg).digest(), "little") % ELL
        assert s < ELL
        negA = ((-A[0]) % P25519, A[1])
        T = [(0, 1), (BX, BY), negA, ed_add((BX, BY), negA)]
        acc = (0, 1)
        for r in range(256):
            row = 256 * slot + r
            j = 255 - r
            bs, bh = (s >> j) & 1, (h >> j) & 1
            assert (int(tab[L["ED_BS"], row]), int(tab[L["ED_BH"], row])) == (bs, bh), ("scalar bits", slot, r)
            if r % 16 == 0 or r in (1, 255):
                X, Y, Z = (_fe(tab, L["ED_ACC"] + 16 * k, row) for k in range(3))
                assert Z % P25519 and (X - acc[0] * Z) % P25519 == 0 and (Y - acc[1] * Z) % P25519 == 0, ("accumulator", slot, r)
                x, y = T[bs + 2 * bh]
                ypx, ymx, t2d = (_fe_signed(tab, L["ED_ADD"] + 16 * k, row) for k in range(3))
                assert (ypx - (y + x)) % P25519 == 0 and (ymx - (y - x)) % P25519 == 0 and (t2d - 2 * D * x * y) % P25519 == 0, ("addend", slot, r)
                for g in range(L["ED_N_MUL"]):
                    c0 = ED_MUL + g * STRIDE
                    cells = [int(tab[c0 + i, row]) for i in range(STRIDE)]
                    assert all(v < (1 << 16) for v in cells[:L["ED_MUL_WHI"]]) and all(v < (1 << 11) for v in cells[L["ED_MUL_WHI"]:])
            acc = ed_add(ed_add(acc, acc), T[bs + 2 * bh])
        # the result leaves on the last row as the product cells of the X4 / Y4 / Z4 gadgets
        row = 256 * slot + 255
        X, Y, Z = (_fe(tab, ED_MUL + L[g] * STRIDE, row) for g in ("ED_G_X4", "ED_G_Y4", "ED_G_Z4"))
        assert (X - acc[0] * Z) % P25519 == 0 and (Y - acc[1] * Z) % P25519 == 0
        want = ed_add(ed_mul(s, (BX, BY)), ed_mul(h, negA))
        assert acc == want == R, ("signature equation [s]B - [h]A == R", slot)


def _fe_signed(tab, col0, row):
    p = 2**64 - 2**32 + 1
    v = 0
    for i in range(16):
        x = int(tab[col0 + i, row])
        v += (x - p if x > p // 2 else x) << (16 * i)
    return v


def check_mul_gadgets_exact(tab, rows):
    """U V = c + q p over the INTEGERS for the doubling's first gadget (X * X) on the given rows: the committed carries are the
    carries of that identity."""
    ED_MUL = L["ED_MUL"]
    for row in rows:
        X = _fe(tab, L["ED_ACC"], row)
        c = _fe(tab, ED_MUL, row)
        q = sum(int(tab[ED_MUL + L["ED_MUL_Q"] + i, row]) << (16 * i) for i in range(17))
        assert X * X == c + q * P25519 and c < P25519
