/*
 * oracle/gl.h -- Goldilocks field and its quadratic extension, CPU restatement.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product (tendermintx_b200/); only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it,
 * as the checker.
 *
 * What it restates: plonky2_field 0.2.0 @ 4f8e6315 (GoldilocksField,
 * QuadraticExtension), reached from the reference through
 * plonky2x::prelude::{GoldilocksField, PlonkParameters}
 * [REF circuits/builder/validator.rs:264, circuits/skip.rs:138-139].
 * The dependency source is NOT under /root/reference (Cargo.lock:2957-2982),
 * so the published algorithm is restated and pinned by the known-answer
 * vectors in SURVEY.md Appendix C (tests/test_oracle_primitives.py).
 *
 *   p = 2^64 - 2^32 + 1, multiplicative generator 7,
 *   2^32-th root of unity 1753635133440165772, extension F_p[X]/(X^2 - 7).
 */
#ifndef TMX_ORACLE_GL_H
#define TMX_ORACLE_GL_H

#include <stdint.h>
#include <stddef.h>

typedef uint64_t gl_t;
typedef unsigned __int128 u128;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */
#define GL_GENERATOR 7ULL
#define GL_ROOT_2_32 1753635133440165772ULL
#define GL_W 7ULL /* extension non-residue */

static inline gl_t gl_canon(gl_t a) { return a - (GL_P & (0 - (uint64_t)(a >= GL_P))); }

/* branch-free: data-dependent branches mispredict half the time on random field elements */
static inline gl_t gl_add(gl_t a, gl_t b) {
    uint64_t s = a + b;
    uint64_t over = (uint64_t)(s < a) | (uint64_t)(s >= GL_P); /* wrapped past 2^64, or landed in [p, 2^64) */
    return s - (GL_P & (0 - over));                           /* s - p (mod 2^64) is right in both cases */
}

static inline gl_t gl_sub(gl_t a, gl_t b) {
    uint64_t d = a - b;
    return d + (GL_P & (0 - (uint64_t)(a < b)));
}

static inline gl_t gl_neg(gl_t a) { return a ? GL_P - a : 0; }

/* 2^64 = 2^32 - 1 and 2^96 = -1 (mod p): x = lo + 2^64*(hl + 2^32*hh) = lo - hh + hl*(2^32-1) */
static inline gl_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t0 = lo - hh;
    t0 -= GL_EPS & (0 - (uint64_t)(lo < hh)); /* borrow: add p back, i.e. subtract 2^32-1 mod 2^64 */
    uint64_t t1 = hl * GL_EPS;
    uint64_t t2 = t0 + t1;
    t2 += GL_EPS & (0 - (uint64_t)(t2 < t1)); /* carry: 2^64 = 2^32-1 */
    return gl_canon(t2);
}

static inline gl_t gl_mul(gl_t a, gl_t b) { return gl_reduce128((u128)a * b); }

static inline gl_t gl_sqr(gl_t a) { return gl_mul(a, a); }

static inline gl_t gl_pow(gl_t b, uint64_t e) {
    gl_t r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_sqr(b);
        e >>= 1;
    }
    return r;
}

static inline gl_t gl_inv(gl_t a) { return gl_pow(a, GL_P - 2); }

/* primitive 2^k-th root of unity, k <= 32 */
static inline gl_t gl_root_of_unity(unsigned k) {
    gl_t r = GL_ROOT_2_32;
    for (unsigned i = k; i < 32; i++) r = gl_sqr(r);
    return r;
}

/* ---- quadratic extension: a0 + a1*X, X^2 = 7 ---- */
typedef struct {
    gl_t a0, a1;
} gl2_t;

static inline gl2_t gl2_make(gl_t a0, gl_t a1) {
    gl2_t r = {a0, a1};
    return r;
}
static inline gl2_t gl2_from(gl_t a) { return gl2_make(a, 0); }
static inline gl2_t gl2_add(gl2_t a, gl2_t b) { return gl2_make(gl_add(a.a0, b.a0), gl_add(a.a1, b.a1)); }
static inline gl2_t gl2_sub(gl2_t a, gl2_t b) { return gl2_make(gl_sub(a.a0, b.a0), gl_sub(a.a1, b.a1)); }
static inline gl2_t gl2_neg(gl2_t a) { return gl2_make(gl_neg(a.a0), gl_neg(a.a1)); }
static inline gl2_t gl2_mul(gl2_t a, gl2_t b) {
    gl_t c0 = gl_add(gl_mul(a.a0, b.a0), gl_mul(GL_W, gl_mul(a.a1, b.a1)));
    gl_t c1 = gl_add(gl_mul(a.a0, b.a1), gl_mul(a.a1, b.a0));
    return gl2_make(c0, c1);
}
static inline gl2_t gl2_scale(gl2_t a, gl_t s) { return gl2_make(gl_mul(a.a0, s), gl_mul(a.a1, s)); }
static inline gl2_t gl2_inv(gl2_t a) {
    /* 1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2) */
    gl_t n = gl_sub(gl_sqr(a.a0), gl_mul(GL_W, gl_sqr(a.a1)));
    gl_t ni = gl_inv(n);
    return gl2_make(gl_mul(a.a0, ni), gl_mul(gl_neg(a.a1), ni));
}
static inline gl2_t gl2_pow(gl2_t b, uint64_t e) {
    gl2_t r = gl2_from(1);
    while (e) {
        if (e & 1) r = gl2_mul(r, b);
        b = gl2_mul(b, b);
        e >>= 1;
    }
    return r;
}
static inline int gl2_eq(gl2_t a, gl2_t b) { return a.a0 == b.a0 && a.a1 == b.a1; }

static inline unsigned tmx_log2(size_t n) {
    unsigned k = 0;
    while (((size_t)1 << k) < n) k++;
    return k;
}
static inline size_t tmx_bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

#endif
