// Poseidon-12 over Goldilocks (width 12, rate 8, x^7, 4+22+4 rounds) for host and device, plus the
// sponge modes plonky2 uses for Merkle leaves / nodes and the duplex challenger.
//
// Replaces plonky2 0.2.0 hash/poseidon.rs + hash/hashing.rs + iop/challenger.rs, which the reference
// reaches through `builder.build()` / `circuit.prove()` [REF circuits/skip.rs:173,214].  Round constants
// are regenerated at start-up (ChaCha8, rand-0.8 seed_from_u64(0), gen_range(0..p)) -- see
// poseidon_generate_constants() in context.cu -- and uploaded once per translation unit.
#pragma once
#include "gl.cuh"

namespace tmx {

constexpr int POSEIDON_WIDTH = 12;
constexpr int POSEIDON_ROUNDS = 30;
constexpr int POSEIDON_HALF_FULL = 4;
constexpr int POSEIDON_PARTIAL = 22;

// host copies (filled by poseidon_generate_constants, context.cu)
extern gl h_poseidon_rc[POSEIDON_ROUNDS * POSEIDON_WIDTH];
// Constants of the fast path, 30 rounds plus one all-zero round (it always adds "the next round's" constants).  Lanes
// 1..11 pass through the 22 partial rounds linearly, so their constants are pushed forward through the linear layers:
// a partial round adds ONE constant (lane 0), and what was deferred arrives with the constants of the first full round
// after them:  h_{r+1} = c_{r+1} + M (h_r with lane 0 cleared),  rows 5..25 keep only lane 0 of h_r, row 26 is all of h_26.
extern gl h_poseidon_rc_fast[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH];
void poseidon_generate_constants();

#if defined(__CUDACC__)
static __constant__ gl d_poseidon_rc[POSEIDON_ROUNDS * POSEIDON_WIDTH];
static __constant__ gl d_poseidon_rc_fast[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH];
static inline cudaError_t poseidon_upload_constants_tu() {
    poseidon_generate_constants();
    cudaError_t e = cudaMemcpyToSymbol(d_poseidon_rc_fast, h_poseidon_rc_fast, sizeof(h_poseidon_rc_fast));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(d_poseidon_rc, h_poseidon_rc, sizeof(h_poseidon_rc));
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cudaStreamLegacy);  // staged copies: wait for the DMA (the provers' streams are non-blocking)
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

// ---- fast path (device; also compiles for the host so CPU tests can pin it against the plain formulation) -------
// Same permutation, fewer instructions:
//   * lanes live in [0, 2^64) (not reduced below p) inside the permutation and are canonicalised once at the end;
//   * the round constants of round r+1 are added inside the linear layer of round r;
//   * reductions fold with 2^64 = 2^32 - 1 using the carry flag instead of compare / select sequences;
//   * the linear layer uses no integer multiplies at all (see below).
TMX_HD gl gl_sqr_nc(gl a) { return gl_mul_nc(a, a); }
TMX_HD gl poseidon_sbox_nc(gl x) {
    const gl x2 = gl_sqr_nc(x), x3 = gl_mul_nc(x2, x), x4 = gl_sqr_nc(x2);
    return gl_mul_nc(x3, x4);
}
TMX_HD gl poseidon_rc_fast(int i) {
#if defined(__CUDA_ARCH__)
    return d_poseidon_rc_fast[i];
#else
    return h_poseidon_rc_fast[i];
#endif
}
// ---- multiplier-free MDS ----------------------------------------------------------------------------------------
// ncu on the IMAD formulation above: the integer-multiply pipe (fmaheavy) is 89 % busy while the ALU pipe idles at
// 42 %, and IMAD.WIDE issues at about 1.2 warp-instructions/clk/SM against 2.0 for IADD3/LEA/LOP3/SHF
// (tools/lab/pipe_bench.cu).  So the linear layer moves off the multiplier: every lane is cut into 22/22/20-bit
// pieces, each piece vector goes through a 12-point cyclic convolution done by the CRT split of Z[x]/(x^12 - 1)
// over y = x^3 (y = 1: cyclic, y = -1: negacyclic, y = i: twisted by i), where all frequency-domain constants of
// this particular circulant are powers of two ([16,32,16], [-1,-8,2], [(2,1),(-4,-1),(16,-1)] after folding the
// 1/4 of the inverse transform), i.e. shifts and adds only.  Arithmetic is wrapping 32-bit and exact because every
// output piece is < 264 * 2^22 < 2^31.  (plonky2's CPU code uses the same algebraic split on 32-bit halves with
// 64-bit intermediates; the piece width here is chosen so that nothing leaves 32-bit registers.)
template <class T>
TMX_HD void mds_conv12_pieces(const T s[12], T out[12]) {
    T u0[3], u2[3], ur[3], ui[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const T e = s[b] + s[6 + b], o = s[3 + b] + s[9 + b];
        u0[b] = e + o;
        u2[b] = e - o;
        ur[b] = s[b] - s[6 + b];
        ui[b] = s[3 + b] - s[9 + b];
    }
    // y = 1 (cyclic), constants 16 * [1, 2, 1]; the factor 16 is applied when the frequencies are merged
    const T t = u0[0] + u0[1] + u0[2];
    const T a0[3] = {t + u0[2], t + u0[0], t + u0[1]};
    // y = -1 (negacyclic), constants [-1, -8, 2]
    const T a2[3] = {(u2[2] << 3) - u2[0] - (u2[1] << 1), (T)0 - (u2[0] << 3) - u2[1] - (u2[2] << 1),
                            (u2[0] << 1) - (u2[1] << 3) - u2[2]};
    // y = i, constants k0 = 2 + i, k1 = -4 - i, k2 = 16 - i; product of (kr + i ki) with (ur + i ui)
    //   z0 = k0 u0 + i (k2 u1 + k1 u2),  z1 = k1 u0 + k0 u1 + i k2 u2,  z2 = k2 u0 + k1 u1 + k0 u2
    const T k0r[3] = {(ur[0] << 1) - ui[0], (ur[1] << 1) - ui[1], (ur[2] << 1) - ui[2]};
    const T k0i[3] = {(ui[0] << 1) + ur[0], (ui[1] << 1) + ur[1], (ui[2] << 1) + ur[2]};
    const T k1r[3] = {ui[0] - (ur[0] << 2), ui[1] - (ur[1] << 2), ui[2] - (ur[2] << 2)};
    const T k1i[3] = {(T)0 - (ui[0] << 2) - ur[0], (T)0 - (ui[1] << 2) - ur[1], (T)0 - (ui[2] << 2) - ur[2]};
    const T k2r[3] = {(ur[0] << 4) + ui[0], (ur[1] << 4) + ui[1], (ur[2] << 4) + ui[2]};
    const T k2i[3] = {(ui[0] << 4) - ur[0], (ui[1] << 4) - ur[1], (ui[2] << 4) - ur[2]};
    const T zr[3] = {k0r[0] - k2i[1] - k1i[2], k1r[0] + k0r[1] - k2i[2], k2r[0] + k1r[1] + k0r[2]};
    const T zi[3] = {k0i[0] + k2r[1] + k1r[2], k1i[0] + k0i[1] + k2r[2], k2i[0] + k1i[1] + k0i[2]};
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const T p = (a0[b] << 4) + a2[b], q = (a0[b] << 4) - a2[b];
        out[b] = p + zr[b];
        out[3 + b] = q + zi[b];
        out[6 + b] = p - zr[b];
        out[9 + b] = q - zi[b];
    }
}
// Lanes that only pass through linear layers (1..11 during the 22 partial rounds) never leave piece form:
//   lane = p0 + p1 2^22 + p2 2^44 (mod p),  p2 in [0, 2^20),  |p0|, |p1| < 2^22 + 2^19  (two's complement),
// so a partial round splits and recombines one lane instead of twelve.  All piece arithmetic is signed wrapping
// 32-bit with floor shifts; outputs of the convolution stay below 264 * (2^22 + 2^19) < 2^31 in magnitude.
TMX_HD void poseidon_split(gl x, uint32_t* p0, uint32_t* p1, uint32_t* p2) {
    *p0 = (uint32_t)x & 0x3FFFFFu;
    *p1 = (uint32_t)(x >> 22) & 0x3FFFFFu;
    *p2 = (uint32_t)(x >> 44);
}
// o0 + o1 2^22 + o2 2^44 + rc (0 <= value < 2^73) -> [0, 2^64), congruent mod p.  Only o0 can be negative: p1 and p2 of
// every lane are non-negative (see poseidon_normalize), so their convolutions are too.
TMX_HD gl poseidon_recombine(uint32_t o0, uint32_t o1, uint32_t o2, gl rc) {
    const int64_t A = (int64_t)(int32_t)o0 + (int64_t)((gl)o1 << 22);
    const gl H = (gl)(A >> 32) + ((gl)o2 << 12);  // >= 0: the words above bit 32
    gl c;
    const gl t = gl_add_carry((H << 32) | (uint32_t)A, rc, &c);
    const gl ov = (H >> 32) + c;  // weight 2^64 = 2^32 - 1, ov < 2^12
    const gl v = gl_add_carry(t, (ov << 32) - ov, &c);
    return v + ((0 - c) & GL_EPS);
}
// the same value back in piece form (carry propagation; the part above bit 64 re-enters as ov 2^32 - ov)
TMX_HD void poseidon_normalize(uint32_t o0, uint32_t o1, uint32_t o2, uint32_t* p0, uint32_t* p1, uint32_t* p2) {
    const int32_t a0 = (int32_t)o0;
    const int32_t a1 = (int32_t)o1 + (a0 >> 22);
    const int32_t a2 = (int32_t)o2 + (a1 >> 22);
    const int32_t ov = a2 >> 20;
    *p0 = (uint32_t)((a0 & 0x3FFFFF) - ov);
    *p1 = (uint32_t)((a1 & 0x3FFFFF) + ov * 1024);
    *p2 = (uint32_t)(a2 & 0xFFFFF);
}

// One loop over the 30 rounds with a single copy of the S-box and convolution code (a three-loop version was
// instruction-cache bound: ncu showed no_instruction as the dominant stall with ~59 KB of straight-line code).
// Round r: S-boxes, then the linear layer, then the constants of round r + 1 (zeros after the last round).
// The loop carries ONE register image of the state, w0/w1/w2[12]: in word form w0 = low word, w1 = high word
// (w2 unused); in piece form w0/w1/w2 = p0/p1/p2.  (Carrying words and pieces side by side doubles the register
// count and halves the occupancy.)
TMX_HD void poseidon_permute_fast(gl s[12]) {
    uint32_t w0[12], w1[12], w2[12], o0[12], o1[12], o2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const gl x = gl_add(s[i], poseidon_rc_fast(i));  // round 0 constants (inputs are canonical)
        w0[i] = (uint32_t)x;
        w1[i] = (uint32_t)(x >> 32);
        w2[i] = 0;
    }
#pragma unroll 1
    for (int r = 0; r < POSEIDON_ROUNDS; r++) {
        const bool full = r < POSEIDON_HALF_FULL || r >= POSEIDON_HALF_FULL + POSEIDON_PARTIAL;
        // lanes 1..11 are needed as 64-bit words again when the next round is full (or the permutation ends)
        const bool words_out = r + 1 < POSEIDON_HALF_FULL || r + 1 >= POSEIDON_HALF_FULL + POSEIDON_PARTIAL;
        poseidon_split(poseidon_sbox_nc(((gl)w1[0] << 32) | w0[0]), &w0[0], &w1[0], &w2[0]);
        if (full) {
#pragma unroll
            for (int i = 1; i < 12; i++) poseidon_split(poseidon_sbox_nc(((gl)w1[i] << 32) | w0[i]), &w0[i], &w1[i], &w2[i]);
        }
        mds_conv12_pieces(w0, o0);
        mds_conv12_pieces(w1, o1);
        mds_conv12_pieces(w2, o2);
        o0[0] += w0[0] << 3;  // + diag(8, 0, ..., 0)
        o1[0] += w1[0] << 3;
        o2[0] += w2[0] << 3;
        const int rc_base = 12 * (r + 1);
        const gl x0 = poseidon_recombine(o0[0], o1[0], o2[0], poseidon_rc_fast(rc_base));
        w0[0] = (uint32_t)x0;
        w1[0] = (uint32_t)(x0 >> 32);
        if (words_out) {
#pragma unroll
            for (int i = 1; i < 12; i++) {
                const gl x = poseidon_recombine(o0[i], o1[i], o2[i], poseidon_rc_fast(rc_base + i));
                w0[i] = (uint32_t)x;
                w1[i] = (uint32_t)(x >> 32);
            }
        } else {
#pragma unroll
            for (int i = 1; i < 12; i++) poseidon_normalize(o0[i], o1[i], o2[i], &w0[i], &w1[i], &w2[i]);  // (their constants are deferred)
        }
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(((gl)w1[i] << 32) | w0[i]);
}

// Host formulation for the transcript (the Fiat-Shamir challenger absorbs every opening on the CPU while the GPU
// waits): the same multiplier-free convolution on the two 32-bit halves of every lane in 64-bit wrapping arithmetic
// (outputs < 264 * 2^32), one 128-bit recombination per lane.
inline void poseidon_permute_host(gl s[12]) {
    uint64_t lo[12], hi[12], ol[12], oh[12];
    for (int r = 0; r < POSEIDON_ROUNDS; r++) {
        const gl* rc = h_poseidon_rc + 12 * r;
        if (r < POSEIDON_HALF_FULL || r >= POSEIDON_HALF_FULL + POSEIDON_PARTIAL) {
            for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add(s[i], rc[i]));
        } else {
            s[0] = poseidon_sbox(gl_add(s[0], rc[0]));
            for (int i = 1; i < 12; i++) s[i] = gl_add(s[i], rc[i]);
        }
        for (int i = 0; i < 12; i++) {
            lo[i] = (uint32_t)s[i];
            hi[i] = s[i] >> 32;
        }
        mds_conv12_pieces<uint64_t>(lo, ol);
        mds_conv12_pieces<uint64_t>(hi, oh);
        ol[0] += lo[0] << 3;
        oh[0] += hi[0] << 3;
        for (int i = 0; i < 12; i++) {
            const unsigned __int128 v = (unsigned __int128)ol[i] + ((unsigned __int128)oh[i] << 32);
            s[i] = gl_reduce128((gl)v, (gl)(v >> 64));
        }
    }
}

TMX_HD void poseidon_permute(gl s[12]) {
    // host too (transcript, verifier): measured 3.2 us per permutation against 5.2 us for poseidon_permute_host and
    // 9.8 us for the plain formulation on the authoring container's CPU
    poseidon_permute_fast(s);
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
