// Per-thread building blocks of the witness kernels (K6 SHA-256, K7 SHA-512, K8 Ed25519): compression with
// per-round history, and the functions that turn one round / one ladder step into the cells of its trace row
// (layout: include/tmx_trace.h).  Host+device so tests can run the same per-thread logic on a CPU
// (tools/hostsim.cpp); the kernels in witness.cu are thin wrappers that map rows to threads.
//
// Replaces the trace generation inside plonky2x's `curta_sha256_variable` and
// `curta_eddsa_verify_sigs_conditional` [REF circuits/builder/verify.rs:202,248-259; validator.rs:228;
// shared.rs:194] and the Merkle gadgets get_root_from_merkle_proof / get_root_from_hashed_leaves
// [REF verify.rs:147,165,285,376; validator.rs:248].
#pragma once
#include "fe25519.cuh"
#include "../../include/tmx_types.h"

namespace tmx {

#define TMX_K256_INIT                                                                                               \
    {0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,    \
     0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,    \
     0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,    \
     0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,    \
     0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,    \
     0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,    \
     0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,    \
     0xc67178f2}
#define TMX_K512_INIT                                                                                               \
    {0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL, \
     0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, \
     0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, \
     0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL, \
     0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL, \
     0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, \
     0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, \
     0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL, \
     0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL, \
     0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, \
     0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, \
     0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL, \
     0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL, \
     0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, \
     0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, \
     0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL}

static const uint32_t h_K256[64] = TMX_K256_INIT;
static const uint64_t h_K512[80] = TMX_K512_INIT;
#if defined(__CUDACC__)
static __constant__ uint32_t d_K256[64] = TMX_K256_INIT;
static __constant__ uint64_t d_K512[80] = TMX_K512_INIT;
#endif
TMX_HD uint32_t k256(int i) {
#if defined(__CUDA_ARCH__)
    return d_K256[i];
#else
    return h_K256[i];
#endif
}
TMX_HD uint64_t k512(int i) {
#if defined(__CUDA_ARCH__)
    return d_K512[i];
#else
    return h_K512[i];
#endif
}
TMX_HD uint32_t iv256(int i) {
    const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    return iv[i];
}
TMX_HD uint64_t iv512(int i) {
    const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                            0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    return iv[i];
}

TMX_HD uint32_t rotr32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
TMX_HD uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

// ------------------------------------------------------------------------------------------- SHA-256
// History of one compression: ah[t + 3] = a before round t (t = -3..64), same for eh; W[0..63]; cv[8].
struct Sha256Hist {
    uint32_t ah[68], eh[68], W[64], cv[8];
};

// blk: 64 message bytes.  Fills hist, writes the new chaining value to out (may alias cv).
TMX_HD void sha256_compress_hist(const uint32_t cv[8], const uint8_t* blk, Sha256Hist* hs, uint32_t out[8]) {
    uint32_t* W = hs->W;
    for (int i = 0; i < 16; i++)
        W[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t x = W[i - 15], y = W[i - 2];
        W[i] = W[i - 16] + (rotr32(x, 7) ^ rotr32(x, 18) ^ (x >> 3)) + W[i - 7] + (rotr32(y, 17) ^ rotr32(y, 19) ^ (y >> 10));
    }
    for (int i = 0; i < 8; i++) hs->cv[i] = cv[i];
    uint32_t a = cv[0], b = cv[1], c = cv[2], d = cv[3], e = cv[4], f = cv[5], g = cv[6], h = cv[7];
    hs->ah[0] = d; hs->ah[1] = c; hs->ah[2] = b; hs->ah[3] = a;
    hs->eh[0] = h; hs->eh[1] = g; hs->eh[2] = f; hs->eh[3] = e;
    for (int t = 0; t < 64; t++) {
        uint32_t t1 = h + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + k256(t) + W[t];
        uint32_t t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        hs->ah[t + 4] = a;
        hs->eh[t + 4] = e;
    }
    uint32_t fin[8] = {a, b, c, d, e, f, g, h};
    for (int i = 0; i < 8; i++) out[i] = hs->cv[i] + fin[i];
}

// All S256_COLS cells of round t (row `row`) from the history.
TMX_HD void sha256_row_cells(gl* trace, size_t n_rows, size_t row, int t, const Sha256Hist* hs) {
    gl* p = trace + row;
    const uint32_t a = hs->ah[t + 3], b = hs->ah[t + 2], c = hs->ah[t + 1], d = hs->ah[t];
    const uint32_t e = hs->eh[t + 3], f = hs->eh[t + 2], g = hs->eh[t + 1], h = hs->eh[t];
    const uint32_t an = hs->ah[t + 4], en = hs->eh[t + 4];
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        p[(size_t)(S256_A + i) * n_rows] = (a >> i) & 1;
        p[(size_t)(S256_B + i) * n_rows] = (b >> i) & 1;
        p[(size_t)(S256_C + i) * n_rows] = (c >> i) & 1;
        p[(size_t)(S256_E + i) * n_rows] = (e >> i) & 1;
        p[(size_t)(S256_F + i) * n_rows] = (f >> i) & 1;
        p[(size_t)(S256_G + i) * n_rows] = (g >> i) & 1;
        p[(size_t)(S256_AN + i) * n_rows] = (an >> i) & 1;
        p[(size_t)(S256_EN + i) * n_rows] = (en >> i) & 1;
    }
    p[(size_t)S256_D * n_rows] = d;
    p[(size_t)S256_H * n_rows] = h;
    const uint64_t t1 = (uint64_t)h + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + k256(t) + hs->W[t];
    const uint64_t t2 = (uint64_t)(rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    const uint64_t ca = (t1 + t2) >> 32, ce = ((uint64_t)d + t1) >> 32;
    for (int i = 0; i < 3; i++) {
        p[(size_t)(S256_CA + i) * n_rows] = (ca >> i) & 1;
        p[(size_t)(S256_CE + i) * n_rows] = (ce >> i) & 1;
    }
    for (int j = 0; j < 16; j++) {
        const int idx = t - 15 + j;
        p[(size_t)(S256_W + j) * n_rows] = idx >= 0 ? hs->W[idx] : 0;
    }
    const uint32_t w14 = t >= 1 ? hs->W[t - 1] : 0, w1 = t >= 14 ? hs->W[t - 14] : 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        p[(size_t)(S256_WB14 + i) * n_rows] = (w14 >> i) & 1;
        p[(size_t)(S256_WB1 + i) * n_rows] = (w1 >> i) & 1;
    }
    for (int j = 0; j < 8; j++) p[(size_t)(S256_CV + j) * n_rows] = hs->cv[j];
    {
        const uint32_t x = t >= 1 ? hs->W[t - 1] : 0, y = t >= 14 ? hs->W[t - 14] : 0;
        const uint64_t s = (uint64_t)(rotr32(x, 17) ^ rotr32(x, 19) ^ (x >> 10)) + (t >= 6 ? hs->W[t - 6] : 0) +
                           (rotr32(y, 7) ^ rotr32(y, 18) ^ (y >> 3)) + (t >= 15 ? hs->W[t - 15] : 0);
        p[(size_t)S256_CW * n_rows] = (s >> 32) & 1;
        p[(size_t)(S256_CW + 1) * n_rows] = (s >> 33) & 1;
        p[(size_t)S256_WS * n_rows] = (uint32_t)s;
    }
    const uint32_t fin[8] = {an, a, b, c, en, e, f, g};
    for (int j = 0; j < 8; j++) {
        uint64_t s = t == 63 ? (uint64_t)hs->cv[j] + fin[j] : 0;
        p[(size_t)(S256_DG + j) * n_rows] = (uint32_t)s;
        p[(size_t)(S256_DC + j) * n_rows] = s >> 32;
    }
}

// FIPS 180-4 padding of msg[0..len) into nb = ceil((len + 9) / 64) blocks written to buf (nb * 64 bytes)
TMX_HD int sha256_pad_blocks(const uint8_t* msg, int len, uint8_t* buf) {
    const int nb = (len + 9 + 63) / 64;
    for (int i = 0; i < nb * 64; i++) buf[i] = i < len ? msg[i] : 0;
    buf[len] = 0x80;
    const uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) buf[nb * 64 - 1 - i] = (uint8_t)(bits >> (8 * i));
    return nb;
}
TMX_HD void sha256_state_to_bytes(const uint32_t st[8], uint8_t out[32]) {
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
    }
}

// REF circuits/builder/shared.rs:67-156 -- nine bytes, continuation bit on every septet below the last non-zero one
TMX_HD void marshal_int64_varint(uint64_t v, uint8_t out[9]) {
    int last = 0;
    for (int i = 0; i < 9; i++)
        if ((v >> (7 * i)) & 0x7F) last = i;
    for (int i = 0; i < 9; i++) out[i] = (uint8_t)(((v >> (7 * i)) & 0x7F) | (i < last ? 0x80 : 0));
}
// REF circuits/builder/validator.rs:185-229 -- leaf message 0x00 || 0a 22 0a 20 <pk> 10 <varint9>, 47 bytes
TMX_HD void validator_leaf_message(const uint8_t pk[32], uint64_t power, uint8_t out[47]) {
    out[0] = 0; out[1] = 10; out[2] = 34; out[3] = 10; out[4] = 32;
    for (int i = 0; i < 32; i++) out[5 + i] = pk[i];
    out[37] = 16;
    marshal_int64_varint(power, out + 38);
}

// ------------------------------------------------------------------------------------------- SHA-512
// history of one chunk: the 80 rounds and the 48 continuation rounds with round constant 0 (include/tmx_trace.h)
struct Sha512Hist {
    uint64_t ah[S512_ROWS_PER_CHUNK + 4], eh[S512_ROWS_PER_CHUNK + 4], W[S512_ROWS_PER_CHUNK], cv[8];
    uint64_t digest[8];  // chaining value + state after round 79 (stays on rows 79..127 of the chunk)
    uint64_t two;        // 1 when the validator slot's message has two blocks (S512_TWO)
};

TMX_HD void sha512_compress_hist(const uint64_t cv[8], const uint8_t* blk, Sha512Hist* hs, uint64_t out[8]) {
    uint64_t* W = hs->W;
    for (int i = 0; i < 16; i++) {
        uint64_t x = 0;
        for (int j = 0; j < 8; j++) x = (x << 8) | blk[8 * i + j];
        W[i] = x;
    }
    for (int i = 16; i < S512_ROWS_PER_CHUNK; i++) {
        uint64_t x = W[i - 15], y = W[i - 2];
        W[i] = W[i - 16] + (rotr64(x, 1) ^ rotr64(x, 8) ^ (x >> 7)) + W[i - 7] + (rotr64(y, 19) ^ rotr64(y, 61) ^ (y >> 6));
    }
    for (int i = 0; i < 8; i++) hs->cv[i] = cv[i];
    uint64_t a = cv[0], b = cv[1], c = cv[2], d = cv[3], e = cv[4], f = cv[5], g = cv[6], h = cv[7];
    hs->ah[0] = d; hs->ah[1] = c; hs->ah[2] = b; hs->ah[3] = a;
    hs->eh[0] = h; hs->eh[1] = g; hs->eh[2] = f; hs->eh[3] = e;
    for (int t = 0; t < S512_ROWS_PER_CHUNK; t++) {
        if (t == S512_ROUNDS) {  // the digest is the state after round 79; the rows beyond only keep the table periodic
            uint64_t fin[8] = {a, b, c, d, e, f, g, h};
            for (int i = 0; i < 8; i++) out[i] = hs->digest[i] = hs->cv[i] + fin[i];
        }
        uint64_t t1 = h + (rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41)) + ((e & f) ^ (~e & g)) + (t < S512_ROUNDS ? k512(t) : 0) + W[t];
        uint64_t t2 = (rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        hs->ah[t + 4] = a;
        hs->eh[t + 4] = e;
    }
}

TMX_HD void sha512_row_cells(gl* trace, size_t n_rows, size_t row, int t, const Sha512Hist* hs) {
    gl* p = trace + row;
    const uint64_t a = hs->ah[t + 3], b = hs->ah[t + 2], c = hs->ah[t + 1], d = hs->ah[t];
    const uint64_t e = hs->eh[t + 3], f = hs->eh[t + 2], g = hs->eh[t + 1], h = hs->eh[t];
    const uint64_t an = hs->ah[t + 4], en = hs->eh[t + 4];
#pragma unroll 4
    for (int i = 0; i < 64; i++) {
        p[(size_t)(S512_A + i) * n_rows] = (a >> i) & 1;
        p[(size_t)(S512_B + i) * n_rows] = (b >> i) & 1;
        p[(size_t)(S512_C + i) * n_rows] = (c >> i) & 1;
        p[(size_t)(S512_E + i) * n_rows] = (e >> i) & 1;
        p[(size_t)(S512_F + i) * n_rows] = (f >> i) & 1;
        p[(size_t)(S512_G + i) * n_rows] = (g >> i) & 1;
        p[(size_t)(S512_AN + i) * n_rows] = (an >> i) & 1;
        p[(size_t)(S512_EN + i) * n_rows] = (en >> i) & 1;
    }
#define TMX_LO(x) ((uint64_t)(uint32_t)(x))
#define TMX_HI(x) ((uint64_t)((x) >> 32))
    p[(size_t)S512_D * n_rows] = TMX_LO(d);
    p[(size_t)(S512_D + 1) * n_rows] = TMX_HI(d);
    p[(size_t)S512_H * n_rows] = TMX_LO(h);
    p[(size_t)(S512_H + 1) * n_rows] = TMX_HI(h);
    const uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41), ch = (e & f) ^ (~e & g);
    const uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39), mj = (a & b) ^ (a & c) ^ (b & c);
    const uint64_t K = t < S512_ROUNDS ? k512(t) : 0, w = hs->W[t];
    const uint64_t t1lo = TMX_LO(h) + TMX_LO(S1) + TMX_LO(ch) + TMX_LO(K) + TMX_LO(w);
    const uint64_t t1hi = TMX_HI(h) + TMX_HI(S1) + TMX_HI(ch) + TMX_HI(K) + TMX_HI(w);
    const uint64_t calo = (t1lo + TMX_LO(S0) + TMX_LO(mj)) >> 32;
    const uint64_t cahi = (t1hi + TMX_HI(S0) + TMX_HI(mj) + calo) >> 32;
    const uint64_t celo = (TMX_LO(d) + t1lo) >> 32;
    const uint64_t cehi = (TMX_HI(d) + t1hi + celo) >> 32;
    for (int i = 0; i < 3; i++) {
        p[(size_t)(S512_CA + i) * n_rows] = (calo >> i) & 1;
        p[(size_t)(S512_CA + 3 + i) * n_rows] = (cahi >> i) & 1;
        p[(size_t)(S512_CE + i) * n_rows] = (celo >> i) & 1;
        p[(size_t)(S512_CE + 3 + i) * n_rows] = (cehi >> i) & 1;
    }
    for (int j = 0; j < 16; j++) {
        const int idx = t - 15 + j;
        const uint64_t v = idx >= 0 ? hs->W[idx] : 0;
        p[(size_t)(S512_W + 2 * j) * n_rows] = TMX_LO(v);
        p[(size_t)(S512_W + 2 * j + 1) * n_rows] = TMX_HI(v);
    }
    const uint64_t w14 = t >= 1 ? hs->W[t - 1] : 0, w1 = t >= 14 ? hs->W[t - 14] : 0;
#pragma unroll 4
    for (int i = 0; i < 64; i++) {
        p[(size_t)(S512_WB14 + i) * n_rows] = (w14 >> i) & 1;
        p[(size_t)(S512_WB1 + i) * n_rows] = (w1 >> i) & 1;
    }
    for (int j = 0; j < 8; j++) {
        p[(size_t)(S512_CV + 2 * j) * n_rows] = TMX_LO(hs->cv[j]);
        p[(size_t)(S512_CV + 2 * j + 1) * n_rows] = TMX_HI(hs->cv[j]);
    }
    // schedule sum of the row's window: sigma1(w[14]) + w[9] + sigma0(w[1]) + w[0], zeros before the block starts
    const uint64_t w9 = t >= 6 ? hs->W[t - 6] : 0, w0 = t >= 15 ? hs->W[t - 15] : 0;
    const uint64_t sg1 = rotr64(w14, 19) ^ rotr64(w14, 61) ^ (w14 >> 6), sg0 = rotr64(w1, 1) ^ rotr64(w1, 8) ^ (w1 >> 7);
    const uint64_t wslo = TMX_LO(sg1) + TMX_LO(w9) + TMX_LO(sg0) + TMX_LO(w0);
    const uint64_t cwlo = wslo >> 32;
    const uint64_t wshi = TMX_HI(sg1) + TMX_HI(w9) + TMX_HI(sg0) + TMX_HI(w0) + cwlo;
    const uint64_t cwhi = wshi >> 32;
    p[(size_t)S512_WS * n_rows] = TMX_LO(wslo);
    p[(size_t)(S512_WS + 1) * n_rows] = TMX_LO(wshi);
    p[(size_t)S512_CW * n_rows] = cwlo & 1;
    p[(size_t)(S512_CW + 1) * n_rows] = (cwlo >> 1) & 1;
    p[(size_t)(S512_CW + 2) * n_rows] = cwhi & 1;
    p[(size_t)(S512_CW + 3) * n_rows] = (cwhi >> 1) & 1;
    const uint64_t fin[8] = {an, a, b, c, en, e, f, g};
    for (int j = 0; j < 8; j++) {
        uint64_t dlo = 0, dhi = 0, clo = 0, chi = 0;
        if (t == 79) {
            const uint64_t lo = TMX_LO(hs->cv[j]) + TMX_LO(fin[j]);
            clo = lo >> 32;
            const uint64_t hi = TMX_HI(hs->cv[j]) + TMX_HI(fin[j]) + clo;
            chi = hi >> 32;
        }
        if (t >= 79) {  // the digest stays on rows 79..127 so that the slot's second chunk can chain from it
            dlo = TMX_LO(hs->digest[j]);
            dhi = TMX_HI(hs->digest[j]);
        }
        p[(size_t)(S512_DG + 2 * j) * n_rows] = dlo;
        p[(size_t)(S512_DG + 2 * j + 1) * n_rows] = dhi;
        p[(size_t)(S512_DC + 2 * j) * n_rows] = clo;
        p[(size_t)(S512_DC + 2 * j + 1) * n_rows] = chi;
    }
    p[(size_t)S512_TWO * n_rows] = hs->two;
#undef TMX_LO
#undef TMX_HI
}

TMX_HD int sha512_pad_blocks(const uint8_t* msg, int len, uint8_t* buf) {
    const int nb = (len + 17 + 127) / 128;
    for (int i = 0; i < nb * 128; i++) buf[i] = i < len ? msg[i] : 0;
    buf[len] = 0x80;
    const uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) buf[nb * 128 - 1 - i] = (uint8_t)(bits >> (8 * i));
    return nb;
}

// ------------------------------------------------------------------------------------------- Ed25519
TMX_HD void fe256_to_limbs(const fe256& x, int32_t l[16]) {
#pragma unroll
    for (int i = 0; i < 16; i++) l[i] = (int32_t)((x.w[i >> 2] >> (16 * (i & 3))) & 0xFFFF);
}
TMX_HD int32_t p_limb(int i) { return i == 0 ? 0xFFED : (i == 15 ? 0x7FFF : 0xFFFF); }

// ---- joint (Straus) evaluation of [s]B + [h](-A): acc' = 2 acc + T[bs + 2 bh], most significant bits first ----
// affine addend in the cached form the mixed addition consumes
struct ge_cached51 {
    fe51 ypx, ymx, t2d;  // y + x, y - x, 2 d x y
};
// canonical accumulator (X, Y, Z) of one row, 96 bytes
struct ge_acc_packed {
    fe256 X, Y, Z;
};
struct ge_acc51 {
    fe51 X, Y, Z;
};
// the four addends of a slot as the ED_ADD cells hold them (48 limbs each: y + x, y - x, 2dxy); see ed_slot_table
struct EdAddendTable {
    int32_t limb[4][48];
};

// one row in field arithmetic: dbl-2008-hwcd (a = -1) followed by add-2008-hwcd-3 with Z2 = 1, no T output
TMX_HD ge_acc51 ed_straus_step(const ge_acc51& p, const ge_cached51& q) {
    const fe51 A = fe_sq(p.X), B = fe_sq(p.Y), Cz = fe_sq(p.Z);
    const fe51 S = fe_sq(fe_add(p.X, p.Y));
    const fe51 E = fe_sub(fe_sub(S, A), B);
    const fe51 G = fe_sub(B, A);
    const fe51 F = fe_sub(G, fe_add(Cz, Cz));
    const fe51 H = fe_sub(fe_zero(), fe_add(A, B));
    const fe51 X3 = fe_mul(E, F), Y3 = fe_mul(G, H), T3 = fe_mul(E, H), Z3 = fe_mul(F, G);
    const fe51 a = fe_mul(fe_sub(Y3, X3), q.ymx), b = fe_mul(fe_add(Y3, X3), q.ypx), c = fe_mul(T3, q.t2d);
    const fe51 d = fe_add(Z3, Z3);
    const fe51 E2 = fe_sub(b, a), F2 = fe_sub(d, c), G2 = fe_add(d, c), H2 = fe_add(b, a);
    ge_acc51 r;
    r.X = fe_mul(E2, F2);
    r.Y = fe_mul(G2, H2);
    r.Z = fe_mul(F2, G2);
    return r;
}

// Addend table of a slot from the decompressed (affine) public key: T = {O, B, -A, B - A}.  The field values go to
// `tab` (for the sequential evaluation), the cell values to `cells`: canonical limbs for O and B; for -A and B - A the
// sums / differences are taken LIMB-WISE (plus 2p where a difference could go negative), because that is the linear
// expression of committed cells the logic table provides on the bus.  (xD, yD) = affine B - A.
TMX_HD void ed_slot_table(const ge51& A, ge_cached51 tab[4], EdAddendTable* cells, fe256* xD_out, fe256* yD_out) {
    const ge51 B = ge_base51();
    tab[0].ypx = fe_one(); tab[0].ymx = fe_one(); tab[0].t2d = fe_zero();
    tab[1].ypx = fe_add(B.Y, B.X); tab[1].ymx = fe_sub(B.Y, B.X); tab[1].t2d = fe_mul(fe_const_2d(), B.T);
    const fe51 tA = fe_mul(fe_const_2d(), A.T);
    tab[2].ypx = fe_sub(A.Y, A.X); tab[2].ymx = fe_add(A.Y, A.X); tab[2].t2d = fe_sub(fe_zero(), tA);
    ge51 nA;
    nA.X = fe_sub(fe_zero(), A.X); nA.Y = A.Y; nA.Z = fe_one(); nA.T = fe_sub(fe_zero(), A.T);
    const ge51 D = ge_add51(B, nA);
    const fe51 zi = fe_invert(D.Z);
    const fe51 xD = fe_mul(D.X, zi), yD = fe_mul(D.Y, zi);
    const fe51 tD = fe_mul(fe_const_2d(), fe_mul(xD, yD));
    tab[3].ypx = fe_add(yD, xD); tab[3].ymx = fe_sub(yD, xD); tab[3].t2d = tD;
    int32_t l0[16], l1[16], l2[16];
    for (int i = 0; i < 48; i++) cells->limb[0][i] = (i == 0 || i == 16) ? 1 : 0;
    fe256_to_limbs(fe_freeze(tab[1].ypx), l0); fe256_to_limbs(fe_freeze(tab[1].ymx), l1); fe256_to_limbs(fe_freeze(tab[1].t2d), l2);
    for (int i = 0; i < 16; i++) { cells->limb[1][i] = l0[i]; cells->limb[1][16 + i] = l1[i]; cells->limb[1][32 + i] = l2[i]; }
    fe256_to_limbs(fe_freeze(A.X), l0); fe256_to_limbs(fe_freeze(A.Y), l1); fe256_to_limbs(fe_freeze(tA), l2);
    for (int i = 0; i < 16; i++) {
        cells->limb[2][i] = l1[i] - l0[i] + 2 * p_limb(i);
        cells->limb[2][16 + i] = l1[i] + l0[i];
        cells->limb[2][32 + i] = 2 * p_limb(i) - l2[i];
    }
    const fe256 xf = fe_freeze(xD), yf = fe_freeze(yD);
    fe256_to_limbs(xf, l0); fe256_to_limbs(yf, l1); fe256_to_limbs(fe_freeze(tD), l2);
    for (int i = 0; i < 16; i++) {
        cells->limb[3][i] = l1[i] + l0[i];
        cells->limb[3][16 + i] = l1[i] - l0[i] + 2 * p_limb(i);
        cells->limb[3][32 + i] = l2[i];
    }
    if (xD_out) *xD_out = xf;
    if (yD_out) *yD_out = yf;
}
TMX_HD void ed_padding_table(ge_cached51 tab[4], EdAddendTable* cells) {
    for (int k = 0; k < 4; k++) {
        tab[k].ypx = fe_one(); tab[k].ymx = fe_one(); tab[k].t2d = fe_zero();
        for (int i = 0; i < 48; i++) cells->limb[k][i] = (i == 0 || i == 16) ? 1 : 0;
    }
}

// The 256 rows of a slot, sequentially: stores the canonical accumulator BEFORE each row and returns the final one.
// Every step continues from the canonical representative, so the stored rows define the next ones exactly.
TMX_HD ge_acc51 ed_straus_ladder(const uint64_t s[4], const uint64_t h[4], const ge_cached51 tab[4], ge_acc_packed* acc_out) {
    ge_acc51 acc;
    acc.X = fe_zero(); acc.Y = fe_one(); acc.Z = fe_one();
#pragma unroll 1
    for (int r = 0; r < ED_ROWS_PER_VALIDATOR; r++) {
        acc_out[r].X = fe_freeze(acc.X);
        acc_out[r].Y = fe_freeze(acc.Y);
        acc_out[r].Z = fe_freeze(acc.Z);
        const int j = 255 - r;
        const int sel = (int)((s[j >> 6] >> (j & 63)) & 1) + 2 * (int)((h[j >> 6] >> (j & 63)) & 1);
        acc = ed_straus_step(acc, tab[sel]);
    }
    return acc;
}

// All ED_COLS cells of row r of a slot from the canonical accumulator before the row, the two scalars and the slot's
// addend cells.  The operand builders are the limb-wise linear combinations of air_ed25519 (same slots as the CPU oracle).
TMX_HD void ed_row_cells(gl* trace, size_t n_rows, size_t row, int r, const uint64_t s[4], const uint64_t h[4], const ge_acc_packed& acc,
                         const EdAddendTable& tab) {
    gl* p = trace + row;
    const int j = 255 - r, pos = r & 15;
    const int bs = (int)((s[j >> 6] >> (j & 63)) & 1), bh = (int)((h[j >> 6] >> (j & 63)) & 1);
    // bits of the current 16-bit limb (index 15 - r / 16) already consumed by the rows above
    const int lsh = 16 * (15 - (r >> 4));
    const uint32_t limb_s = (uint32_t)((s[lsh >> 6] >> (lsh & 63)) & 0xFFFF), limb_h = (uint32_t)((h[lsh >> 6] >> (lsh & 63)) & 0xFFFF);
    p[(size_t)ED_BS * n_rows] = (gl)bs;
    p[(size_t)ED_BH * n_rows] = (gl)bh;
    p[(size_t)ED_SACC_S * n_rows] = (gl)(pos ? limb_s >> (16 - pos) : 0);
    p[(size_t)ED_SACC_H * n_rows] = (gl)(pos ? limb_h >> (16 - pos) : 0);
    int32_t X1[16], Y1[16], Z1[16];
    fe256_to_limbs(acc.X, X1); fe256_to_limbs(acc.Y, Y1); fe256_to_limbs(acc.Z, Z1);
    const int32_t* add = tab.limb[bs + 2 * bh];
    for (int i = 0; i < 16; i++) {
        p[(size_t)(ED_ACC + i) * n_rows] = (gl)X1[i];
        p[(size_t)(ED_ACC + 16 + i) * n_rows] = (gl)Y1[i];
        p[(size_t)(ED_ACC + 32 + i) * n_rows] = (gl)Z1[i];
    }
    for (int i = 0; i < 48; i++) p[(size_t)(ED_ADD + i) * n_rows] = (gl)add[i];
    int32_t u[16], A[16], B[16], C[16], S[16], E[16], F[16], G[16], H[16], X3[16], Y3[16], T3[16], Z3[16], dump[16];
    auto cells = [&](int m) { return p + (size_t)(ED_MUL + m * ED_MUL_STRIDE) * n_rows; };
    mul_gadget_cells(X1, X1, cells(ED_G_A), n_rows, A);
    mul_gadget_cells(Y1, Y1, cells(ED_G_B), n_rows, B);
    mul_gadget_cells(Z1, Z1, cells(ED_G_CZ), n_rows, C);
    for (int i = 0; i < 16; i++) u[i] = X1[i] + Y1[i];
    mul_gadget_cells(u, u, cells(ED_G_S), n_rows, S);
    for (int i = 0; i < 16; i++) {
        E[i] = S[i] - A[i] - B[i] + 2 * p_limb(i);
        G[i] = B[i] - A[i] + p_limb(i);
        F[i] = B[i] - A[i] - 2 * C[i] + 3 * p_limb(i);
        H[i] = 2 * p_limb(i) - A[i] - B[i];
    }
    mul_gadget_cells(E, F, cells(ED_G_X3), n_rows, X3);
    mul_gadget_cells(G, H, cells(ED_G_Y3), n_rows, Y3);
    mul_gadget_cells(E, H, cells(ED_G_T3), n_rows, T3);
    mul_gadget_cells(F, G, cells(ED_G_Z3), n_rows, Z3);
    for (int i = 0; i < 16; i++) u[i] = Y3[i] - X3[i] + p_limb(i);
    mul_gadget_cells(u, add + 16, cells(ED_G_AA), n_rows, A);
    for (int i = 0; i < 16; i++) u[i] = Y3[i] + X3[i];
    mul_gadget_cells(u, add, cells(ED_G_BB), n_rows, B);
    mul_gadget_cells(T3, add + 32, cells(ED_G_CC), n_rows, C);
    for (int i = 0; i < 16; i++) {
        E[i] = B[i] - A[i] + p_limb(i);
        F[i] = 2 * Z3[i] - C[i] + p_limb(i);
        G[i] = 2 * Z3[i] + C[i];
        H[i] = B[i] + A[i];
    }
    mul_gadget_cells(E, F, cells(ED_G_X4), n_rows, dump);
    mul_gadget_cells(G, H, cells(ED_G_Y4), n_rows, dump);
    mul_gadget_cells(F, G, cells(ED_G_Z4), n_rows, dump);
}

}  // namespace tmx
