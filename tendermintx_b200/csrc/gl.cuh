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

// Branch-free on purpose: on the host (transcript, verifier) data-dependent branches on random field elements
// mispredict half the time; on the device the compiler turns these into selects either way.
TMX_HD gl gl_canon(gl a) { return a - (GL_P & (0 - (gl)(a >= GL_P))); }

TMX_HD gl gl_add(gl a, gl b) {
    const gl s = a + b;
    const gl over = (gl)(s < a) | (gl)(s >= GL_P);  // wrapped past 2^64, or landed in [p, 2^64)
    return s - (GL_P & (0 - over));                 // s - p (mod 2^64) is right in both cases (a, b < p)
}
TMX_HD gl gl_sub(gl a, gl b) { return (a - b) + (GL_P & (0 - (gl)(a < b))); }
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

TMX_HD gl gl_mul(gl a, gl b) {
#if defined(__CUDA_ARCH__)
    return gl_reduce128(a * b, __umul64hi(a, b));
#else
    unsigned __int128 m = (unsigned __int128)a * b;
    return gl_reduce128((gl)m, (gl)(m >> 64));
#endif
}
TMX_HD gl gl_sqr(gl a) { return gl_mul(a, a); }

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
