/*
 * oracle/poseidon.c -- Poseidon-12 over Goldilocks, sponge hashing, Merkle caps and the
 * duplex challenger.  TEST INFRASTRUCTURE ONLY (see oracle/gl.h header).
 *
 * Restates plonky2 0.2.0 @ 4f8e6315 (hash/poseidon.rs, hash/hashing.rs, hash/merkle_tree.rs,
 * iop/challenger.rs).  The source is not under /root/reference (Cargo.lock:2957-2959); the
 * reference reaches it through `builder.build()` / `circuit.prove()` / `circuit.verify()`
 * [REF circuits/skip.rs:173,214,244,247].  Pinned by the three permutation known-answer vectors
 * and the first six round constants in SURVEY.md Appendix C.
 */
#include "oracle.h"
#include <string.h>
#include <stdlib.h>

/* ---- round constants: ChaCha8 keystream, rand-0.8 seed_from_u64(0), gen_range(0..p) ---- */
static gl_t RC[POSEIDON_N_ROUNDS * POSEIDON_WIDTH];
static int rc_ready = 0;

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t rotr32(uint32_t x, int r) { return (x >> r) | (x << ((32 - r) & 31)); }

#define QR(a, b, c, d)                     \
    a += b; d ^= a; d = rotl32(d, 16);     \
    c += d; b ^= c; b = rotl32(b, 12);     \
    a += b; d ^= a; d = rotl32(d, 8);      \
    c += d; b ^= c; b = rotl32(b, 7);

static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (uint32_t)counter;
    s[13] = (uint32_t)(counter >> 32);
    s[14] = 0;
    s[15] = 0;
    uint32_t x[16];
    memcpy(x, s, sizeof x);
    for (int r = 0; r < 4; r++) { /* 8 rounds = 4 double rounds */
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

static void poseidon_init_constants(void) {
    /* PCG32 expansion of the u64 seed 0 into the 32-byte ChaCha key (rand_core::SeedableRng::seed_from_u64) */
    uint64_t state = 0;
    uint32_t key[8];
    for (int i = 0; i < 8; i++) {
        state = state * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
        key[i] = rotr32(xs, (int)(state >> 59));
    }
    uint32_t blk[16];
    uint64_t ctr = 0;
    int pos = 16, n = 0;
    while (n < POSEIDON_N_ROUNDS * POSEIDON_WIDTH) {
        uint32_t w[2];
        for (int j = 0; j < 2; j++) {
            if (pos == 16) {
                chacha8_block(key, ctr++, blk);
                pos = 0;
            }
            w[j] = blk[pos++];
        }
        uint64_t v = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
        u128 m = (u128)v * GL_P; /* widening multiply sampler; zone = p - 1 */
        if ((uint64_t)m <= GL_P - 1) RC[n++] = (gl_t)(m >> 64);
    }
    rc_ready = 1;
}

const gl_t *poseidon_round_constants(void) {
    if (!rc_ready) poseidon_init_constants();
    return RC;
}


static inline gl_t sbox7(gl_t x) {
    gl_t x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

/* out[r] = sum_i s[(i + r) % 12] * C[i] + D[r] * s[r].  The constants are below 2^6, so the two 32-bit halves of
 * every lane are summed separately in 64 bits (no 128-bit products) and recombined once per output. */
static void mds_layer(gl_t s[12]) {
    static const uint32_t C32[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[24], hi[24];
    for (int i = 0; i < 12; i++) {
        lo[i] = lo[i + 12] = (uint32_t)s[i];
        hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32);
    }
    uint64_t al[12], ah[12];
    for (int r = 0; r < 12; r++) {
        uint64_t a = 0, b = 0;
        for (int i = 0; i < 12; i++) {
            a += (uint64_t)lo[i + r] * C32[i];
            b += (uint64_t)hi[i + r] * C32[i];
        }
        al[r] = a;
        ah[r] = b;
    }
    al[0] += (uint64_t)lo[0] * 8;
    ah[0] += (uint64_t)hi[0] * 8;
    for (int r = 0; r < 12; r++) s[r] = gl_reduce128((u128)al[r] + ((u128)ah[r] << 32));
}

void poseidon_permute(gl_t s[12]) {
    if (!rc_ready) poseidon_init_constants();
    int rc = 0;
    for (int r = 0; r < POSEIDON_N_ROUNDS; r++) {
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], RC[rc++]);
        int full = (r < 4) || (r >= 4 + 22);
        if (full)
            for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        else
            s[0] = sbox7(s[0]);
        mds_layer(s);
    }
}

/* hash_no_pad: overwrite-mode sponge, rate 8 */
void poseidon_hash_no_pad(const gl_t *in, size_t n, gl_t out[4]) {
    gl_t s[12] = {0};
    for (size_t off = 0; off < n; off += 8) {
        size_t m = n - off < 8 ? n - off : 8;
        for (size_t i = 0; i < m; i++) s[i] = in[off + i];
        poseidon_permute(s);
    }
    memcpy(out, s, 4 * sizeof(gl_t));
}

/* hash_or_noop: <= 4 elements are copied (zero padded) */
void poseidon_hash_or_noop(const gl_t *in, size_t n, gl_t out[4]) {
    if (n <= 4) {
        memset(out, 0, 4 * sizeof(gl_t));
        memcpy(out, in, n * sizeof(gl_t));
    } else
        poseidon_hash_no_pad(in, n, out);
}

void poseidon_two_to_one(const gl_t l[4], const gl_t r[4], gl_t out[4]) {
    gl_t s[12] = {0};
    memcpy(s, l, 32);
    memcpy(s + 4, r, 32);
    poseidon_permute(s);
    memcpy(out, s, 32);
}

/* ---- Merkle tree with cap (plonky2 hash/merkle_tree.rs semantics) ----
 * leaves: n_leaves rows of leaf_len elements (row-major).  digests layout here is OUR choice:
 * level 0 = leaf digests (n), level 1 = n/2, ... down to the cap level (1<<cap_height digests).
 * Cap digest i is the root of the subtree over leaves [i*n/2^cap, (i+1)*n/2^cap). */
void merkle_build(merkle_tree_t *t, const gl_t *leaves, size_t n_leaves, size_t leaf_len, unsigned cap_height) {
    unsigned lg = tmx_log2(n_leaves);
    if (cap_height > lg) cap_height = lg;
    t->n_leaves = n_leaves;
    t->leaf_len = leaf_len;
    t->cap_height = cap_height;
    t->n_levels = lg - cap_height + 1;
    size_t total = 0;
    for (unsigned l = 0; l < t->n_levels; l++) total += n_leaves >> l;
    t->digests = (gl_t *)malloc(total * 4 * sizeof(gl_t));
    gl_t *lvl = t->digests;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_leaves; i++) poseidon_hash_or_noop(leaves + i * leaf_len, leaf_len, lvl + 4 * i);
    for (unsigned l = 1; l < t->n_levels; l++) {
        size_t m = n_leaves >> l;
        gl_t *nxt = lvl + 4 * (n_leaves >> (l - 1));
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < m; i++) poseidon_two_to_one(lvl + 8 * i, lvl + 8 * i + 4, nxt + 4 * i);
        lvl = nxt;
    }
    t->cap = lvl;
}

void merkle_free(merkle_tree_t *t) {
    free(t->digests);
    t->digests = NULL;
}

/* sibling path from leaf up to (excluding) the cap level */
size_t merkle_prove(const merkle_tree_t *t, size_t leaf_index, gl_t *siblings) {
    const gl_t *lvl = t->digests;
    size_t idx = leaf_index, n = t->n_leaves, k = 0;
    for (unsigned l = 0; l + 1 < t->n_levels; l++) {
        memcpy(siblings + 4 * k++, lvl + 4 * (idx ^ 1), 32);
        lvl += 4 * n;
        n >>= 1;
        idx >>= 1;
    }
    return k;
}

int merkle_verify(const gl_t *leaf, size_t leaf_len, size_t leaf_index, const gl_t *siblings, size_t n_sib,
                  const gl_t *cap, unsigned cap_height) {
    gl_t cur[4], nxt[4];
    poseidon_hash_or_noop(leaf, leaf_len, cur);
    size_t idx = leaf_index;
    for (size_t k = 0; k < n_sib; k++) {
        if (idx & 1)
            poseidon_two_to_one(siblings + 4 * k, cur, nxt);
        else
            poseidon_two_to_one(cur, siblings + 4 * k, nxt);
        memcpy(cur, nxt, 32);
        idx >>= 1;
    }
    (void)cap_height;
    return memcmp(cur, cap + 4 * idx, 32) == 0;
}

/* ---- duplex challenger (plonky2 iop/challenger.rs) ---- */
void challenger_init(challenger_t *c) { memset(c, 0, sizeof *c); }

static void challenger_duplex(challenger_t *c) {
    for (int i = 0; i < c->n_in; i++) c->state[i] = c->in[i];
    c->n_in = 0;
    poseidon_permute(c->state);
    memcpy(c->out, c->state, 8 * sizeof(gl_t));
    c->n_out = 8;
}

void challenger_observe(challenger_t *c, gl_t x) {
    c->n_out = 0;
    c->in[c->n_in++] = x;
    if (c->n_in == 8) challenger_duplex(c);
}

void challenger_observe_many(challenger_t *c, const gl_t *x, size_t n) {
    for (size_t i = 0; i < n; i++) challenger_observe(c, x[i]);
}

gl_t challenger_get(challenger_t *c) {
    if (c->n_in != 0 || c->n_out == 0) challenger_duplex(c);
    return c->out[--c->n_out];
}

gl2_t challenger_get_ext(challenger_t *c) {
    gl_t a = challenger_get(c);
    gl_t b = challenger_get(c);
    return gl2_make(a, b);
}

/* proof-of-work: the SMALLEST witness w such that, after absorbing the pending inputs and w,
 * leading_zeros(state[7]) >= bits.  Upstream uses rayon find_any (non-deterministic). */
gl_t challenger_pow_grind(const challenger_t *c, unsigned bits) {
    gl_t base[12];
    memcpy(base, c->state, sizeof base);
    for (int i = 0; i < c->n_in; i++) base[i] = c->in[i];
    int pos = c->n_in;
    for (uint64_t cand = 0;; cand++) {
        gl_t s[12];
        memcpy(s, base, sizeof s);
        s[pos] = cand;
        poseidon_permute(s);
        if (bits == 0 || (s[7] >> (64 - bits)) == 0) return cand;
    }
}
