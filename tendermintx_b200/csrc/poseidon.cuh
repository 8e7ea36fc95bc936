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
static __constant__ gl d_poseidon_rc[POSEIDON_ROUNDS * POSEIDON_WIDTH];
static inline cudaError_t poseidon_upload_constants_tu() {
    poseidon_generate_constants();
    return cudaMemcpyToSymbol(d_poseidon_rc, h_poseidon_rc, sizeof(h_poseidon_rc));
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

TMX_HD void poseidon_permute(gl s[12]) {
    int rc = 0;
#pragma unroll 1
    for (int r = 0; r < POSEIDON_HALF_FULL; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add(s[i], poseidon_rc(rc + i)));
        rc += 12;
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 0; r < POSEIDON_PARTIAL; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], poseidon_rc(rc + i));
        rc += 12;
        s[0] = poseidon_sbox(s[0]);
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 0; r < POSEIDON_HALF_FULL; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add(s[i], poseidon_rc(rc + i)));
        rc += 12;
        poseidon_mds(s);
    }
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
