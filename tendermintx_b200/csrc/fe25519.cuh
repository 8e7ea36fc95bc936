// Curve25519 field arithmetic (5 x 51-bit limbs, products in 128 bits), extended Edwards points, the
// double-and-add step whose 17 multiplications the Ed25519 table witnesses, scalar reduction mod l, and
// the 16-bit-limb multiplication gadget witness (c, q, carries).  Host+device so the per-thread logic of
// the kernels can be exercised on a CPU by tests/ (tools/hostsim.cu); the product only runs it on the GPU.
//
// Replaces the Ed25519 witness generation inside plonky2x's `curta_eddsa_verify_sigs_conditional`
// [REF circuits/builder/verify.rs:248-259].
#pragma once
#include <cstdint>
#include "gl.cuh"
#include "../../include/tmx_trace.h"

namespace tmx {

typedef unsigned __int128 u128;

struct fe51 {
    uint64_t v[5];
};
constexpr uint64_t M51 = (1ULL << 51) - 1;

TMX_HD fe51 fe_zero() { return fe51{{0, 0, 0, 0, 0}}; }
TMX_HD fe51 fe_one() { return fe51{{1, 0, 0, 0, 0}}; }

TMX_HD fe51 fe_carry(fe51 a) {
    uint64_t c;
    c = a.v[0] >> 51; a.v[0] &= M51; a.v[1] += c;
    c = a.v[1] >> 51; a.v[1] &= M51; a.v[2] += c;
    c = a.v[2] >> 51; a.v[2] &= M51; a.v[3] += c;
    c = a.v[3] >> 51; a.v[3] &= M51; a.v[4] += c;
    c = a.v[4] >> 51; a.v[4] &= M51; a.v[0] += 19 * c;
    c = a.v[0] >> 51; a.v[0] &= M51; a.v[1] += c;
    return a;
}
TMX_HD fe51 fe_add(const fe51& a, const fe51& b) {
    fe51 r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.v[i] = a.v[i] + b.v[i];
    return fe_carry(r);
}
// a - b with a, b carried (limbs < 2^52): add 4p first
TMX_HD fe51 fe_sub(const fe51& a, const fe51& b) {
    fe51 r;
    r.v[0] = a.v[0] + 0x1FFFFFFFFFFFB4ULL - b.v[0];
#pragma unroll
    for (int i = 1; i < 5; i++) r.v[i] = a.v[i] + 0x1FFFFFFFFFFFFCULL - b.v[i];
    return fe_carry(r);
}
TMX_HD fe51 fe_mul(const fe51& a, const fe51& b) {
    const uint64_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4];
    const uint64_t b0 = b.v[0], b1 = b.v[1], b2 = b.v[2], b3 = b.v[3], b4 = b.v[4];
    const uint64_t b1_19 = 19 * b1, b2_19 = 19 * b2, b3_19 = 19 * b3, b4_19 = 19 * b4;
    u128 r0 = (u128)a0 * b0 + (u128)a1 * b4_19 + (u128)a2 * b3_19 + (u128)a3 * b2_19 + (u128)a4 * b1_19;
    u128 r1 = (u128)a0 * b1 + (u128)a1 * b0 + (u128)a2 * b4_19 + (u128)a3 * b3_19 + (u128)a4 * b2_19;
    u128 r2 = (u128)a0 * b2 + (u128)a1 * b1 + (u128)a2 * b0 + (u128)a3 * b4_19 + (u128)a4 * b3_19;
    u128 r3 = (u128)a0 * b3 + (u128)a1 * b2 + (u128)a2 * b1 + (u128)a3 * b0 + (u128)a4 * b4_19;
    u128 r4 = (u128)a0 * b4 + (u128)a1 * b3 + (u128)a2 * b2 + (u128)a3 * b1 + (u128)a4 * b0;
    fe51 r;
    uint64_t c;
    c = (uint64_t)(r0 >> 51); r.v[0] = (uint64_t)r0 & M51; r1 += c;
    c = (uint64_t)(r1 >> 51); r.v[1] = (uint64_t)r1 & M51; r2 += c;
    c = (uint64_t)(r2 >> 51); r.v[2] = (uint64_t)r2 & M51; r3 += c;
    c = (uint64_t)(r3 >> 51); r.v[3] = (uint64_t)r3 & M51; r4 += c;
    c = (uint64_t)(r4 >> 51); r.v[4] = (uint64_t)r4 & M51;
    r.v[0] += 19 * c;
    c = r.v[0] >> 51; r.v[0] &= M51; r.v[1] += c;
    return r;
}
TMX_HD fe51 fe_sq(const fe51& a) { return fe_mul(a, a); }

// canonical value as four little-endian 64-bit words
struct fe256 {
    uint64_t w[4];
};
TMX_HD fe256 fe_freeze(fe51 a) {
    a = fe_carry(a);
    // two strict passes: every limb < 2^51 and the value < 2^255 afterwards
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            a.v[i] += c;
            c = a.v[i] >> 51;
            a.v[i] &= M51;
        }
        a.v[0] += 19 * c;
    }
    // subtract p if a >= p: a + 19 overflows bit 255 exactly then, and (a + 19) - 2^255 = a - p
    uint64_t t[5];
    uint64_t c = 19;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        t[i] = a.v[i] + c;
        c = t[i] >> 51;
        t[i] &= M51;
    }
    if (c) {
#pragma unroll
        for (int i = 0; i < 5; i++) a.v[i] = t[i];
    }
    fe256 r;
    r.w[0] = a.v[0] | (a.v[1] << 51);
    r.w[1] = (a.v[1] >> 13) | (a.v[2] << 38);
    r.w[2] = (a.v[2] >> 26) | (a.v[3] << 25);
    r.w[3] = (a.v[3] >> 39) | (a.v[4] << 12);
    return r;
}
TMX_HD fe51 fe_from256(const fe256& x) {
    fe51 r;
    r.v[0] = x.w[0] & M51;
    r.v[1] = ((x.w[0] >> 51) | (x.w[1] << 13)) & M51;
    r.v[2] = ((x.w[1] >> 38) | (x.w[2] << 26)) & M51;
    r.v[3] = ((x.w[2] >> 25) | (x.w[3] << 39)) & M51;
    r.v[4] = (x.w[3] >> 12) & M51;  // drops bit 255
    return r;
}
TMX_HD fe256 fe256_from_bytes(const uint8_t* b) {
    fe256 r;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) w |= (uint64_t)b[8 * i + j] << (8 * j);
        r.w[i] = w;
    }
    return r;
}
TMX_HD bool fe256_eq(const fe256& a, const fe256& b) {
    return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3];
}
TMX_HD bool fe256_is_zero(const fe256& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }
// is the 255-bit value (bit 255 cleared) < p ?
TMX_HD bool fe256_lt_p(const fe256& a) {
    const uint64_t top = a.w[3] & 0x7FFFFFFFFFFFFFFFULL;
    if (top != 0x7FFFFFFFFFFFFFFFULL) return true;
    if (a.w[2] != ~0ULL || a.w[1] != ~0ULL) return true;
    return a.w[0] < 0xFFFFFFFFFFFFFFEDULL;
}

TMX_HD fe51 fe_const_d() {
    return fe51{{0x34DCA135978A3ULL, 0x1A8283B156EBDULL, 0x5E7A26001C029ULL, 0x739C663A03CBBULL, 0x52036CEE2B6FFULL}};
}
TMX_HD fe51 fe_const_2d() {
    return fe51{{0x69B9426B2F159ULL, 0x35050762ADD7AULL, 0x3CF44C0038052ULL, 0x6738CC7407977ULL, 0x2406D9DC56DFFULL}};
}
TMX_HD fe51 fe_const_sqrtm1() {
    return fe51{{0x61B274A0EA0B0ULL, 0x0D5A5FC8F189DULL, 0x7EF5E9CBD0C60ULL, 0x78595A6804C9EULL, 0x2B8324804FC1DULL}};
}

// a^(2^252 - 3)
TMX_HD fe51 fe_pow22523(const fe51& a) {
    fe51 acc = fe_one();
#pragma unroll 1
    for (int i = 251; i >= 0; i--) {
        acc = fe_sq(acc);
        if (i != 1) acc = fe_mul(acc, a);
    }
    return acc;
}

// a^(p - 2), p - 2 = 2^255 - 21: all exponent bits set except bits 2 and 4
TMX_HD fe51 fe_invert(const fe51& a) {
    fe51 acc = fe_one();
#pragma unroll 1
    for (int i = 254; i >= 0; i--) {
        acc = fe_sq(acc);
        if (i != 2 && i != 4) acc = fe_mul(acc, a);
    }
    return acc;
}

struct ge51 {
    fe51 X, Y, Z, T;
};
TMX_HD ge51 ge_identity51() { return ge51{fe_zero(), fe_one(), fe_one(), fe_zero()}; }
TMX_HD ge51 ge_base51() {
    ge51 b;
    b.X = fe51{{0x62D608F25D51AULL, 0x412A4B4F6592AULL, 0x75B7171A4B31DULL, 0x1FF60527118FEULL, 0x216936D3CD6E5ULL}};
    b.Y = fe51{{0x6666666666658ULL, 0x4CCCCCCCCCCCCULL, 0x1999999999999ULL, 0x3333333333333ULL, 0x6666666666666ULL}};
    b.Z = fe_one();
    b.T = fe_mul(b.X, b.Y);
    return b;
}

// RFC 8032 decoding with canonical-y requirement.  Returns false if not a curve point.
TMX_HD bool ge_decompress51(const uint8_t enc[32], ge51* out) {
    fe256 yb = fe256_from_bytes(enc);
    const int sign = enc[31] >> 7;
    if (!fe256_lt_p(yb)) return false;
    fe51 y = fe_from256(yb);
    fe51 y2 = fe_sq(y);
    fe51 u = fe_sub(y2, fe_one());
    fe51 v = fe_add(fe_mul(y2, fe_const_d()), fe_one());
    fe51 v3 = fe_mul(fe_sq(v), v);
    fe51 v7 = fe_mul(fe_sq(v3), v);
    fe51 x = fe_mul(fe_mul(u, v3), fe_pow22523(fe_mul(u, v7)));
    fe256 chk = fe_freeze(fe_mul(fe_sq(x), v));
    fe256 uf = fe_freeze(u);
    if (!fe256_eq(chk, uf)) {
        fe256 nu = fe_freeze(fe_sub(fe_zero(), u));
        if (!fe256_eq(chk, nu)) return false;
        x = fe_mul(x, fe_const_sqrtm1());
    }
    fe256 xf = fe_freeze(x);
    if (fe256_is_zero(xf) && sign) return false;
    if ((int)(xf.w[0] & 1) != sign) x = fe_sub(fe_zero(), x);
    out->X = x;
    out->Y = y;
    out->Z = fe_one();
    out->T = fe_mul(x, y);
    return true;
}

// add-2008-hwcd-3 and dbl-2008-hwcd for a = -1, exactly the formulas of DESIGN.md "Ed25519 table"
TMX_HD ge51 ge_add51(const ge51& p, const ge51& q) {
    fe51 A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
    fe51 B = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
    fe51 C = fe_mul(fe_mul(p.T, q.T), fe_const_2d());
    fe51 Dh = fe_mul(p.Z, q.Z);
    fe51 D = fe_add(Dh, Dh);
    fe51 E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    ge51 r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.T = fe_mul(E, H);
    r.Z = fe_mul(F, G);
    return r;
}
TMX_HD ge51 ge_dbl51(const ge51& p) {
    fe51 A = fe_sq(p.X), B = fe_sq(p.Y), Cz = fe_sq(p.Z);
    fe51 S = fe_sq(fe_add(p.X, p.Y));
    fe51 E = fe_sub(fe_sub(S, A), B);
    fe51 G = fe_sub(B, A);
    fe51 F = fe_sub(G, fe_add(Cz, Cz));
    fe51 H = fe_sub(fe_zero(), fe_add(A, B));
    ge51 r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.T = fe_mul(E, H);
    r.Z = fe_mul(F, G);
    return r;
}
TMX_HD bool ge_equal51(const ge51& a, const ge51& b) {
    return fe256_eq(fe_freeze(fe_mul(a.X, b.Z)), fe_freeze(fe_mul(b.X, a.Z))) &&
           fe256_eq(fe_freeze(fe_mul(a.Y, b.Z)), fe_freeze(fe_mul(b.Y, a.Z)));
}

// canonical point packed as 16 little-endian 64-bit words (X, Y, Z, T)
struct ge_packed {
    fe256 X, Y, Z, T;
};
TMX_HD ge_packed ge_pack(const ge51& p) { return ge_packed{fe_freeze(p.X), fe_freeze(p.Y), fe_freeze(p.Z), fe_freeze(p.T)}; }

// ---- scalar arithmetic mod l = 2^252 + 27742317777372353535851937790883648493 ----
TMX_HD bool sc_lt_l(const uint64_t s[4]) {
    const uint64_t L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};
    for (int i = 3; i >= 0; i--) {
        if (s[i] < L[i]) return true;
        if (s[i] > L[i]) return false;
    }
    return false;
}
// 512-bit little-endian digest mod l, bit-serial (runs once per validator)
TMX_HD void sc_reduce512(const uint8_t in[64], uint64_t out[4]) {
    const uint64_t L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};
    uint64_t r[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int bit = 511; bit >= 0; bit--) {
        r[3] = (r[3] << 1) | (r[2] >> 63);
        r[2] = (r[2] << 1) | (r[1] >> 63);
        r[1] = (r[1] << 1) | (r[0] >> 63);
        r[0] = (r[0] << 1) | ((in[bit >> 3] >> (bit & 7)) & 1);
        if (!sc_lt_l(r)) {
            uint64_t borrow = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint64_t a = r[i], b = L[i];
                uint64_t d = a - b - borrow;
                borrow = (a < b) || (a == b && borrow) ? 1 : 0;
                r[i] = d;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = r[i];
}

// ---- multiplication gadget witness on 16-bit limbs ----
// U, V: signed limb vectors (|limb| < 2^19); writes the 63 cells of one gadget (c[16], q[17], and the 15 carries out of the
// odd limbs, offset by ED_W_OFFSET and split into a 16-bit and an 11-bit part) with stride `stride` starting at `cells`;
// returns c limbs in c_out.
TMX_HD void mul_gadget_cells(const int32_t U[16], const int32_t V[16], gl* cells, size_t stride, int32_t c_out[16], const int32_t* W = nullptr) {
    int64_t t[31];
#pragma unroll
    for (int k = 0; k < 31; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) t[i + j] += (int64_t)U[i] * V[j];
    if (W)  // additive operand of the logic table's gadgets: U * V + W = c + q * p
        for (int k = 0; k < 16; k++) t[k] += W[k];
    // canonical residue: digits of N, fold 2^256 = 38, conditional subtraction of p
    int64_t d[36];
    int64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 36; k++) {
        int64_t s = (k < 31 ? t[k] : 0) + carry;
        carry = s >> 16;
        d[k] = s & 0xFFFF;
    }
    // fold the high digits down with 2^256 = 38 (mod p); four rounds always suffice for N < 2^520
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        carry = 0;
#pragma unroll
        for (int k = 0; k < 36; k++) {
            int64_t s = carry + (k < 16 ? d[k] : 0) + (k < 20 ? 38 * d[16 + k] : 0);
            carry = s >> 16;
            d[k] = s & 0xFFFF;
        }
    }
    // now value < 2^256 (digits >= 16 are zero); subtract p while >= p (at most twice)
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        bool ge = true;
        for (int k = 15; k >= 0; k--) {
            const int64_t pk = k == 0 ? 0xFFED : (k == 15 ? 0x7FFF : 0xFFFF);
            if (d[k] > pk) break;
            if (d[k] < pk) { ge = false; break; }
        }
        if (!ge) break;
        int64_t b = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int64_t pk = k == 0 ? 0xFFED : (k == 15 ? 0x7FFF : 0xFFFF);
            int64_t s = d[k] - pk + b;
            b = s >> 16;
            d[k] = s & 0xFFFF;
        }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        c_out[k] = (int32_t)d[k];
        cells[(size_t)k * stride] = (gl)d[k];
    }
    // exact low-to-high division by p: q limbs and the carries of the identity t - c - q*p = 0
    int64_t q[17];
    carry = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        int64_t s = (k < 31 ? t[k] : 0) - (k < 16 ? d[k] : 0) + carry;
#pragma unroll
        for (int i = 0; i < 17; i++) {
            if (i < k && k - i < 16) {
                const int64_t pk = (k - i) == 15 ? 0x7FFF : 0xFFFF;  // k - i >= 1 here
                s -= q[i] * pk;
            }
        }
        if (k < 17) {
            const int64_t qk = ((s & 0xFFFF) * 0x35E5) & 0xFFFF;
            q[k] = qk;
            s -= qk * 0xFFED;
            cells[(size_t)(ED_MUL_Q + k) * stride] = (gl)qk;
        }
        carry = s >> 16;
        if ((k & 1) && k < 31) {
            const int64_t w = carry + ED_W_OFFSET;
            cells[(size_t)(ED_MUL_WLO + (k >> 1)) * stride] = (gl)(w & 0xFFFF);
            cells[(size_t)(ED_MUL_WHI + (k >> 1)) * stride] = (gl)(w >> 16);
        }
    }
}

}  // namespace tmx
