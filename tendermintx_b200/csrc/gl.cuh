// Goldilocks field (p = 2^64 - 2^32 + 1) and its quadratic extension F_p[X]/(X^2-7) for host and
// device code of the product library.  Values are kept canonical (< p) between operations so every
// buffer is bit-comparable with the CPU oracle.
//
// Replaces (on the GPU) plonky2_field 0.2.0 GoldilocksField / QuadraticExtension, which the reference
// reaches through plonky2x::prelude [REF circuits/builder/validator.rs:264, circuits/skip.rs:138-139].
#pragma once
#include <cstdint>
#include <cstddef>

#if defined(__CUDACC__)
#define TMX_HD __host__ __device__ __forceinline__
#define TMX_D __device__ __forceinline__
#else
#define TMX_HD inline
#define TMX_D inline
#endif

namespace tmx {

typedef uint64_t gl;
constexpr gl GL_P = 0xFFFFFFFF00000001ULL;
constexpr gl GL_EPS = 0xFFFFFFFFULL;  // 2^64 mod p
constexpr gl GL_GEN = 7ULL;           // multiplicative generator, also the LDE coset shift
constexpr gl GL_ROOT_2_32 = 1753635133440165772ULL;

// Host versions are branch-free on purpose (transcript, verifier: data-dependent branches on random field elements
// mispredict half the time).  Device versions use the carry flag: the compare / select formulation costs about
// twice the instructions there (ncu on the NTT: SEL + ISETP were a quarter of all issued instructions).
TMX_HD gl gl_add_carry(gl a, gl b, gl* carry) {  // a + b mod 2^64, *carry = 0 / 1
#if defined(__CUDA_ARCH__)
    gl s, c;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(s), "=l"(c) : "l"(a), "l"(b));
    *carry = c;
    return s;
#else
    const gl s = a + b;
    *carry = (gl)(s < a);
    return s;
#endif
}
TMX_HD gl gl_sub_borrow_mask(gl a, gl b, gl* mask) {  // a - b mod 2^64, *mask = borrow ? ~0 : 0
#if defined(__CUDA_ARCH__)
    gl d, m;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u64 %1, 0, 0;" : "=l"(d), "=l"(m) : "l"(a), "l"(b));
    *mask = m;
    return d;
#else
    *mask = (gl)0 - (gl)(a < b);
    return a - b;
#endif
}

TMX_HD gl gl_canon(gl a) {  // [0, 2^64) -> [0, p)
#if defined(__CUDA_ARCH__)
    gl t;
    // a - p = a + (2^32 - 1) mod 2^64 carries exactly when a >= p; the carry goes straight into a predicate
    asm("{\n\t.reg .u32 al, ah, tl, th, c;\n\t.reg .pred q;\n\tmov.b64 {al, ah}, %1;\n\t"
        "add.cc.u32 tl, al, 0xffffffff;\n\taddc.cc.u32 th, ah, 0;\n\taddc.u32 c, 0, 0;\n\t"
        "setp.ne.u32 q, c, 0;\n\tselp.u32 tl, tl, al, q;\n\tselp.u32 th, th, ah, q;\n\tmov.b64 %0, {tl, th};\n\t}"
        : "=l"(t) : "l"(a));
    return t;
#else
    return a - (GL_P & (0 - (gl)(a >= GL_P)));
#endif
}

TMX_HD gl gl_sub(gl a, gl b) {  // a, b < p (b = p is also fine: used by gl_add)
#if defined(__CUDA_ARCH__)
    gl d;
    // on borrow add p, i.e. subtract 2^32 - 1 from the wrapped difference: the borrow word (0 / 0xffffffff) IS that
    // constant.  No second borrow: a - b + 2^64 >= 2^64 - p + 1 = 2^32.
    asm("{\n\t.reg .u32 al, ah, bl, bh, dl, dh, m;\n\tmov.b64 {al, ah}, %1;\n\tmov.b64 {bl, bh}, %2;\n\t"
        "sub.cc.u32 dl, al, bl;\n\tsubc.cc.u32 dh, ah, bh;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 dl, dl, m;\n\tsubc.u32 dh, dh, 0;\n\tmov.b64 %0, {dl, dh};\n\t}"
        : "=l"(d) : "l"(a), "l"(b));
    return d;
#else
    return (a - b) + (GL_P & (0 - (gl)(a < b)));
#endif
}
TMX_HD gl gl_add(gl a, gl b) {
#if defined(__CUDA_ARCH__)
    return gl_sub(a, GL_P - b);  // a - (p - b): seven instructions against ten for add / compare / select
#else
    const gl s = a + b;
    const gl over = (gl)(s < a) | (gl)(s >= GL_P);  // wrapped past 2^64, or landed in [p, 2^64)
    return s - (GL_P & (0 - over));                 // s - p (mod 2^64) is right in both cases (a, b < p)
#endif
}
TMX_HD gl gl_neg(gl a) { return a ? GL_P - a : 0; }

// x = lo + 2^64*hi with 2^64 = 2^32-1, 2^96 = -1  (mod p)
TMX_HD gl gl_reduce128(gl lo, gl hi) {
    const gl hh = hi >> 32, hl = hi & GL_EPS;
    gl t0 = lo - hh;
    t0 -= GL_EPS & (0 - (gl)(lo < hh));
    const gl t1 = hl * GL_EPS;
    gl t2 = t0 + t1;
    t2 += GL_EPS & (0 - (gl)(t2 < t1));
    return gl_canon(t2);
}

// ---- lazily reduced arithmetic: values in [0, 2^64), congruent mod p (used inside Poseidon and the Horner chains
// of the quotient kernel; canonicalise with gl_canon before a value leaves the kernel) ----
// 64 x 64 -> 128: one mul.lo / mul.hi pair lets ptxas share the partial products (three IMAD.WIDE, one
// IMAD.WIDE.X and four carry instructions); a * b next to __umul64hi(a, b) in C costs five wide and two narrow
// multiplies, and a hand-written schoolbook on 32-bit halves pays for zero-extended register pairs.
TMX_HD void gl_mul128(gl a, gl b, gl* lo, gl* hi) {
#if defined(__CUDA_ARCH__)
    asm("mul.lo.u64 %0, %2, %3;\n\tmul.hi.u64 %1, %2, %3;" : "=l"(*lo), "=l"(*hi) : "l"(a), "l"(b));
#else
    const unsigned __int128 m = (unsigned __int128)a * b;
    *lo = (gl)m;
    *hi = (gl)(m >> 64);
#endif
}
TMX_HD gl gl_reduce128_nc(gl lo, gl hi) {  // result in [0, 2^64), congruent mod p
    gl m, c;
    const gl hh = hi >> 32, hl = hi & GL_EPS;
    gl t0 = gl_sub_borrow_mask(lo, hh, &m);
    t0 -= (m & GL_EPS);
    const gl t1 = (hl << 32) - hl;
    const gl t2 = gl_add_carry(t0, t1, &c);
    return t2 + ((0 - c) & GL_EPS);
}
TMX_HD gl gl_mul_nc(gl a, gl b) {
    gl lo, hi;
    gl_mul128(a, b, &lo, &hi);
    return gl_reduce128_nc(lo, hi);
}
// acc * a + x with one reduction (the 128-bit product cannot overflow when x is added: hi <= 2^64 - 2)
TMX_HD gl gl_mac_nc(gl acc, gl a, gl x) {
    gl lo, hi, c;
    gl_mul128(acc, a, &lo, &hi);
    lo = gl_add_carry(lo, x, &c);
    return gl_reduce128_nc(lo, hi + c);
}

// 192-bit accumulator for long dot products: sum of up to 2^63 products of two values in [0, 2^64), reduced once.
// (A canonical multiply-add costs ~35 instructions, a product plus a three-word add 14.)
struct gl_acc192 {
    gl w0, w1, w2;
};
TMX_HD gl_acc192 gl_acc_zero() {
    gl_acc192 a;
    a.w0 = a.w1 = a.w2 = 0;
    return a;
}
TMX_HD void gl_acc_mac(gl_acc192& acc, gl a, gl b) {
    gl lo, hi;
    gl_mul128(a, b, &lo, &hi);
#if defined(__CUDA_ARCH__)
    asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, 0;" : "+l"(acc.w0), "+l"(acc.w1), "+l"(acc.w2) : "l"(lo), "l"(hi));
#else
    gl c;
    acc.w0 = gl_add_carry(acc.w0, lo, &c);
    gl c2, c3;
    acc.w1 = gl_add_carry(acc.w1, c, &c2);
    acc.w1 = gl_add_carry(acc.w1, hi, &c3);
    acc.w2 += c2 + c3;
#endif
}
// canonical value of w0 + 2^64 w1 + 2^128 w2 with w2 < 2^32 (2^128 = -2^32 mod p)
TMX_HD gl gl_acc_reduce(const gl_acc192& acc) {
    const gl r = gl_canon(gl_reduce128_nc(acc.w0, acc.w1));
    const gl t = gl_canon(gl_reduce128_nc(acc.w2 << 32, 0));
    return gl_sub(r, t);
}

TMX_HD gl gl_mul(gl a, gl b) {
#if defined(__CUDA_ARCH__)
    return gl_canon(gl_mul_nc(a, b));
#else
    unsigned __int128 m = (unsigned __int128)a * b;
    return gl_reduce128((gl)m, (gl)(m >> 64));
#endif
}
TMX_HD gl gl_sqr(gl a) { return gl_mul(a, a); }
// 7 a, canonical (a < 2^64): the part above 2^64 is below 7
TMX_HD gl gl_mul7(gl a) {
    gl lo, hi, c;
    gl_mul128(a, 7, &lo, &hi);
    const gl t = gl_add_carry(lo, (hi << 32) - hi, &c);
    return gl_canon(t + ((0 - c) & GL_EPS));
}

TMX_HD gl gl_pow(gl b, uint64_t e) {
    gl r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_sqr(b);
        e >>= 1;
    }
    return r;
}
TMX_HD gl gl_inv(gl a) { return gl_pow(a, GL_P - 2); }
TMX_HD gl gl_root_of_unity(unsigned k) {
    gl r = GL_ROOT_2_32;
    for (unsigned i = k; i < 32; i++) r = gl_sqr(r);
    return r;
}

struct gl2 {
    gl a0, a1;
};
TMX_HD gl2 gl2_make(gl a0, gl a1) {
    gl2 r;
    r.a0 = a0;
    r.a1 = a1;
    return r;
}
TMX_HD gl2 gl2_from(gl a) { return gl2_make(a, 0); }
TMX_HD gl2 gl2_add(gl2 a, gl2 b) { return gl2_make(gl_add(a.a0, b.a0), gl_add(a.a1, b.a1)); }
TMX_HD gl2 gl2_sub(gl2 a, gl2 b) { return gl2_make(gl_sub(a.a0, b.a0), gl_sub(a.a1, b.a1)); }
TMX_HD gl2 gl2_neg(gl2 a) { return gl2_make(gl_neg(a.a0), gl_neg(a.a1)); }
TMX_HD gl2 gl2_mul(gl2 a, gl2 b) {
    gl c0 = gl_add(gl_mul(a.a0, b.a0), gl_mul(7, gl_mul(a.a1, b.a1)));
    gl c1 = gl_add(gl_mul(a.a0, b.a1), gl_mul(a.a1, b.a0));
    return gl2_make(c0, c1);
}
TMX_HD gl2 gl2_scale(gl2 a, gl s) { return gl2_make(gl_mul(a.a0, s), gl_mul(a.a1, s)); }
TMX_HD gl2 gl2_inv(gl2 a) {
    gl n = gl_sub(gl_sqr(a.a0), gl_mul(7, gl_sqr(a.a1)));
    gl ni = gl_inv(n);
    return gl2_make(gl_mul(a.a0, ni), gl_mul(gl_neg(a.a1), ni));
}
TMX_HD gl2 gl2_pow(gl2 b, uint64_t e) {
    gl2 r = gl2_from(1);
    while (e) {
        if (e & 1) r = gl2_mul(r, b);
        b = gl2_mul(b, b);
        e >>= 1;
    }
    return r;
}
TMX_HD bool gl2_eq(gl2 a, gl2 b) { return a.a0 == b.a0 && a.a1 == b.a1; }

TMX_HD unsigned ilog2(size_t n) {
    unsigned k = 0;
    while (((size_t)1 << k) < n) k++;
    return k;
}
TMX_HD uint32_t bitrev32(uint32_t x, unsigned bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

}  // namespace tmx
