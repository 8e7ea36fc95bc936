// Poseidon-12 over Goldilocks (width 12, rate 8, x^7, 4+22+4 rounds) for host and device, plus the
// sponge modes plonky2 uses for Merkle leaves / nodes and the duplex challenger.
//
// Replaces plonky2 0.2.0 hash/poseidon.rs + hash/hashing.rs + iop/challenger.rs, which the reference
// reaches through `builder.build()` / `circuit.prove()` [REF circuits/skip.rs:173,214].  Round constants
// are regenerated at start-up (ChaCha8, rand-0.8 seed_from_u64(0), gen_range(0..p)) -- see
// poseidon_generate_constants() in host.cpp -- and uploaded once per translation unit.
#pragma once
#include "gl.cuh"

namespace tmx {

constexpr int POSEIDON_WIDTH = 12;
constexpr int POSEIDON_ROUNDS = 30;
constexpr int POSEIDON_HALF_FULL = 4;
constexpr int POSEIDON_PARTIAL = 22;

// host copy (filled by poseidon_generate_constants, host.cpp)
extern gl h_poseidon_rc[POSEIDON_ROUNDS * POSEIDON_WIDTH];
void poseidon_generate_constants();

#if defined(__CUDACC__)
// 30 rounds of constants plus one all-zero round (the fast path always adds "the next round's" constants)
static __constant__ gl d_poseidon_rc[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH];
static inline cudaError_t poseidon_upload_constants_tu() {
    poseidon_generate_constants();
    gl padded[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH] = {0};
    for (int i = 0; i < POSEIDON_ROUNDS * POSEIDON_WIDTH; i++) padded[i] = h_poseidon_rc[i];
    return cudaMemcpyToSymbol(d_poseidon_rc, padded, sizeof(padded));
}
#endif

TMX_HD gl poseidon_rc(int i) {
#if defined(__CUDA_ARCH__)
    return d_poseidon_rc[i];
#else
    return h_poseidon_rc[i];
#endif
}

TMX_HD gl poseidon_sbox(gl x) {
    gl x2 = gl_sqr(x), x3 = gl_mul(x2, x), x4 = gl_sqr(x2);
    return gl_mul(x3, x4);
}

// circulant MDS: out[r] = sum_i s[(i+r)%12]*C[i] + 8*s[0] (r==0); constants < 2^6 so the two 32-bit halves
// of every lane are accumulated separately in 64 bits and recombined once.
TMX_HD void poseidon_mds(gl s[12]) {
    const uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (uint64_t)lo[(i + r) % 12] * C[i];
            ah += (uint64_t)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) {
            al += (uint64_t)lo[0] * 8u;
            ah += (uint64_t)hi[0] * 8u;
        }
        // value = al + ah*2^32  (ah < 2^40)
        uint64_t l64 = al + (ah << 32);
        uint64_t h64 = (ah >> 32) + (l64 < al ? 1u : 0u);
        s[r] = gl_reduce128(l64, h64);
    }
}

// Reference formulation (host; also the semantic definition): add constants, S-box, MDS, all canonical.
TMX_HD void poseidon_permute_plain(gl s[12]) {
    int rc = 0;
#pragma unroll 1
    for (int r = 0; r < POSEIDON_ROUNDS; r++) {
        const bool full = r < POSEIDON_HALF_FULL || r >= POSEIDON_HALF_FULL + POSEIDON_PARTIAL;
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], poseidon_rc(rc + i));
        rc += 12;
        if (full) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(s[i]);
        } else
            s[0] = poseidon_sbox(s[0]);
        poseidon_mds(s);
    }
}

#if defined(__CUDACC__)
// ---- device fast path -------------------------------------------------------------------------------------------
// Same permutation, fewer instructions (the kernels are bound by integer issue, not memory):
//   * lanes live in [0, 2^64) (not reduced below p) inside the permutation and are canonicalised once at the end;
//   * the round constants of round r+1 seed the MDS accumulators of round r, so no separate modular additions;
//   * the MDS sums the 32-bit halves of the lanes in 64-bit accumulators (constants < 2^6) and folds the result with
//     2^64 = 2^32 - 1 using the carry flag instead of compare / select sequences.
__device__ __forceinline__ gl gl_reduce128_nc(gl lo, gl hi) {  // result in [0, 2^64), congruent mod p
    gl t0, t2, m;
    const gl hh = hi >> 32, hl = hi & GL_EPS;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u64 %1, 0, 0;" : "=l"(t0), "=l"(m) : "l"(lo), "l"(hh));  // m = borrow ? ~0 : 0
    t0 -= (m & GL_EPS);
    const gl t1 = (hl << 32) - hl;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(t2), "=l"(m) : "l"(t0), "l"(t1));  // m = carry
    return t2 + ((0 - m) & GL_EPS);
}
__device__ __forceinline__ gl gl_mul_nc(gl a, gl b) { return gl_reduce128_nc(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ gl poseidon_sbox_nc(gl x) {
    const gl x2 = gl_mul_nc(x, x), x3 = gl_mul_nc(x2, x), x4 = gl_mul_nc(x2, x2);
    return gl_mul_nc(x3, x4);
}
// out[r] = sum_i s[(i+r)%12] C[i] + 8 s[0] [r == 0] + rc[rc_base + r]
__device__ __forceinline__ void poseidon_mds_rc(gl s[12], int rc_base) {
    const uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        const gl k = d_poseidon_rc[rc_base + r];
        uint64_t al = (uint32_t)k, ah = k >> 32;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (uint64_t)lo[(i + r) % 12] * C[i];
            ah += (uint64_t)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) {
            al += (uint64_t)lo[0] * 8u;
            ah += (uint64_t)hi[0] * 8u;
        }
        // value = al + ah * 2^32 with al, ah < 2^41:  al + (ah_low32 << 32) + (ah >> 32) * (2^32 - 1)
        const uint64_t t = al + (ah >> 32) * GL_EPS;  // < 2^42
        uint64_t v, c;
        asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(v), "=l"(c) : "l"(t), "l"(ah << 32));
        s[r] = v + ((0 - c) & GL_EPS);
    }
}
// One loop over the 30 rounds; the 11-lane S-box block only runs in the 8 full rounds.  Keeping a single copy of
// the MDS / S-box code matters: the previous three-loop version was instruction-cache bound (ncu: no_instruction
// was the dominant stall with ~59 KB of straight-line code).
__device__ __forceinline__ void poseidon_permute_dev(gl s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], d_poseidon_rc[i]);  // round 0 constants (inputs are canonical)
#pragma unroll 1
    for (int r = 0; r < POSEIDON_ROUNDS; r++) {
        s[0] = poseidon_sbox_nc(s[0]);
        if (r < POSEIDON_HALF_FULL || r >= POSEIDON_HALF_FULL + POSEIDON_PARTIAL) {
#pragma unroll
            for (int i = 1; i < 12; i++) s[i] = poseidon_sbox_nc(s[i]);
        }
        poseidon_mds_rc(s, 12 * (r + 1));  // + constants of round r + 1 (zeros after the last round)
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}
#endif

TMX_HD void poseidon_permute(gl s[12]) {
#if defined(__CUDA_ARCH__)
    poseidon_permute_dev(s);
#else
    poseidon_permute_plain(s);
#endif
}

TMX_HD void poseidon_two_to_one(const gl l[4], const gl r[4], gl out[4]) {
    gl s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        s[i] = l[i];
        s[4 + i] = r[i];
        s[8 + i] = 0;
    }
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = s[i];
}

// hash_or_noop of a strided row: element i is p[i*stride]
TMX_HD void poseidon_hash_row(const gl* p, size_t stride, size_t n, gl out[4]) {
    if (n <= 4) {
        for (size_t i = 0; i < 4; i++) out[i] = i < n ? p[i * stride] : 0;
        return;
    }
    gl s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (size_t off = 0; off < n; off += 8) {
        if (off + 8 <= n) {
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = p[(off + i) * stride];
        } else {
            for (size_t i = 0; off + i < n; i++) s[i] = p[(off + i) * stride];
        }
        poseidon_permute(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = s[i];
}

// Duplex challenger (host side; tiny data).  plonky2 iop/challenger.rs semantics: observe() buffers up to
// RATE inputs and invalidates pending outputs; get() duplexes if needed and pops from the END of state[0..8].
struct Challenger {
    gl state[12];
    gl in[8];
    gl out[8];
    int n_in, n_out;
    Challenger() : n_in(0), n_out(0) {
        for (int i = 0; i < 12; i++) state[i] = 0;
    }
    void duplex() {
        for (int i = 0; i < n_in; i++) state[i] = in[i];
        n_in = 0;
        poseidon_permute(state);
        for (int i = 0; i < 8; i++) out[i] = state[i];
        n_out = 8;
    }
    void observe(gl x) {
        n_out = 0;
        in[n_in++] = x;
        if (n_in == 8) duplex();
    }
    void observe(const gl* x, size_t n) {
        for (size_t i = 0; i < n; i++) observe(x[i]);
    }
    void observe_ext(gl2 x) {
        observe(x.a0);
        observe(x.a1);
    }
    gl get() {
        if (n_in != 0 || n_out == 0) duplex();
        return out[--n_out];
    }
    gl2 get_ext() {
        gl a = get();
        gl b = get();
        return gl2_make(a, b);
    }
};

}  // namespace tmx
