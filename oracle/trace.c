/*
 * oracle/trace.c -- fills the SHA-256, SHA-512 and Ed25519 witness tables (include/tmx_trace.h) on the CPU.
 * TEST INFRASTRUCTURE ONLY (see oracle/gl.h header).
 *
 * The tables witness the hashing and signature work that the reference requests at
 * [REF circuits/builder/verify.rs:202,205,248-259,285,376; validator.rs:228,248; shared.rs:194,197].
 * The SHA-256 chunk order is fixed by the circuit shape (DESIGN.md "SHA-256 schedule"):
 *   for each validator set (skip: trusted, then target; step: target):
 *       N leaf hashes (1 chunk each), then the inner nodes level by level (2 chunks each, Np-1 nodes);
 *   then the header proofs (leaf, then 4 inner nodes each):
 *       skip: trusted validators-hash, target validators-hash, chain id, height
 *       step: validators-hash, chain id, height, last-block-id (2-chunk leaf), prev next-validators-hash.
 */
#include "oracle_w.h"
#include "../include/tmx_trace.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    size_t n_rows, n_cols;
    uint64_t *data;
} trace_t;

static size_t pow2_at_least(size_t x) {
    size_t p = 1;
    while (p < x) p *= 2;
    return p;
}

size_t tm_sha256_chunks(uint32_t kind, uint32_t n_max) {
    size_t np = pow2_at_least(n_max);
    size_t vh = n_max + 2 * (np - 1);
    return kind == TMX_KIND_SKIP ? 2 * vh + 36 : vh + 46;
}

/* dims: rows256, cols256, rows512, cols512, rowsEd, colsEd */
void tm_trace_dims(uint32_t kind, uint32_t n_max, size_t dims[6]) {
    dims[0] = pow2_at_least(tm_sha256_chunks(kind, n_max) * S256_ROUNDS);
    dims[1] = S256_COLS;
    dims[2] = pow2_at_least((size_t)n_max * S512_ROWS_PER_VALIDATOR);
    dims[3] = S512_COLS;
    dims[4] = pow2_at_least((size_t)n_max * ED_ROWS_PER_VALIDATOR);
    dims[5] = ED_COLS;
}

#define CELL(t, col, row) (t)->data[(size_t)(col) * (t)->n_rows + (row)]

static inline uint32_t ror32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
static inline uint64_t ror64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

/* ------------------------------------------------------------------ SHA-256 rows */
static void sha256_rows(trace_t *t, size_t row0, const uint32_t cv[8], const uint8_t blk[64], uint32_t out_state[8]) {
    sha256_round_t r[64];
    uint32_t st[8];
    memcpy(st, cv, sizeof st);
    sha256_compress(st, blk, r);
    memcpy(out_state, st, sizeof st);
    uint32_t W[64];
    for (int i = 0; i < 64; i++) W[i] = r[i].w;
    for (int i = 0; i < 64; i++) {
        size_t row = row0 + i;
        const uint32_t *v = r[i].v;
        const int bit_groups[6][2] = {{S256_A, 0}, {S256_B, 1}, {S256_C, 2}, {S256_E, 4}, {S256_F, 5}, {S256_G, 6}};
        for (int g = 0; g < 6; g++)
            for (int b = 0; b < 32; b++) CELL(t, bit_groups[g][0] + b, row) = (v[bit_groups[g][1]] >> b) & 1;
        CELL(t, S256_D, row) = v[3];
        CELL(t, S256_H, row) = v[7];
        uint64_t sa = r[i].t1 + r[i].t2, se = (uint64_t)v[3] + r[i].t1;
        uint32_t an = (uint32_t)sa, en = (uint32_t)se;
        for (int b = 0; b < 32; b++) {
            CELL(t, S256_AN + b, row) = (an >> b) & 1;
            CELL(t, S256_EN + b, row) = (en >> b) & 1;
        }
        for (int b = 0; b < 3; b++) {
            CELL(t, S256_CA + b, row) = ((sa >> 32) >> b) & 1;
            CELL(t, S256_CE + b, row) = ((se >> 32) >> b) & 1;
        }
        for (int j = 0; j < 16; j++) {
            int idx = i - 15 + j;
            CELL(t, S256_W + j, row) = idx >= 0 ? W[idx] : 0;
        }
        uint32_t w14 = i >= 1 ? W[i - 1] : 0, w1 = i >= 14 ? W[i - 14] : 0;
        for (int b = 0; b < 32; b++) {
            CELL(t, S256_WB14 + b, row) = (w14 >> b) & 1;
            CELL(t, S256_WB1 + b, row) = (w1 >> b) & 1;
        }
        for (int j = 0; j < 8; j++) CELL(t, S256_CV + j, row) = cv[j];
        {
            uint32_t a = i >= 1 ? W[i - 1] : 0, c = i >= 14 ? W[i - 14] : 0;
            uint64_t s = (uint64_t)(ror32(a, 17) ^ ror32(a, 19) ^ (a >> 10)) + (i >= 6 ? W[i - 6] : 0) +
                         (ror32(c, 7) ^ ror32(c, 18) ^ (c >> 3)) + (i >= 15 ? W[i - 15] : 0);
            CELL(t, S256_CW, row) = (s >> 32) & 1;
            CELL(t, S256_CW + 1, row) = (s >> 33) & 1;
            CELL(t, S256_WS, row) = (uint32_t)s;
        }
        for (int j = 0; j < 8; j++) {
            uint64_t dg = 0, dc = 0;
            if (i == 63) {
                const uint32_t fin[8] = {an, v[0], v[1], v[2], en, v[4], v[5], v[6]};
                uint64_t s = (uint64_t)cv[j] + fin[j];
                dg = (uint32_t)s;
                dc = s >> 32;
            }
            CELL(t, S256_DG + j, row) = dg;
            CELL(t, S256_DC + j, row) = dc;
        }
    }
}

typedef struct {
    trace_t *t;
    size_t chunk; /* next free chunk slot */
} sha256_cursor_t;

/* hash msg[0..len) as one message occupying whole chunks; returns digest */
static void sha256_emit(sha256_cursor_t *c, const uint8_t *msg, size_t len, uint8_t digest[32]) {
    uint8_t buf[64 * 4];
    size_t nb = sha256_pad(msg, len, buf);
    uint32_t st[8], nxt[8];
    memcpy(st, SHA256_IV, sizeof st);
    for (size_t b = 0; b < nb; b++) {
        sha256_rows(c->t, c->chunk * S256_ROUNDS, st, buf + 64 * b, nxt);
        memcpy(st, nxt, sizeof st);
        c->chunk++;
    }
    for (int i = 0; i < 8; i++) {
        digest[4 * i] = st[i] >> 24; digest[4 * i + 1] = st[i] >> 16; digest[4 * i + 2] = st[i] >> 8; digest[4 * i + 3] = st[i];
    }
}

static void emit_inner(sha256_cursor_t *c, const uint8_t l[32], const uint8_t r[32], uint8_t out[32]) {
    uint8_t m[65];
    m[0] = 1;
    memcpy(m + 1, l, 32);
    memcpy(m + 33, r, 32);
    sha256_emit(c, m, 65, out);
}

static void emit_validator_set(sha256_cursor_t *c, const uint8_t (*pk)[32], const uint64_t *power, const uint32_t *blen,
                               size_t n_max, size_t nb_enabled) {
    size_t np = pow2_at_least(n_max);
    uint8_t *val = (uint8_t *)calloc(np, 32);
    uint8_t *en = (uint8_t *)calloc(np, 1);
    for (size_t i = 0; i < n_max; i++) {
        uint8_t buf[64] = {0};
        tm_marshal_validator(pk[i], power[i], buf + 1);
        size_t len = 1 + blen[i];
        if (len > 47) len = 47;
        sha256_emit(c, buf, len, val + 32 * i);
        en[i] = i < nb_enabled;
    }
    for (size_t m = np; m > 1; m /= 2)
        for (size_t i = 0; i < m / 2; i++) {
            uint8_t h[32];
            emit_inner(c, val + 64 * i, val + 64 * i + 32, h);
            if (en[2 * i] && en[2 * i + 1])
                memcpy(val + 32 * i, h, 32);
            else
                memmove(val + 32 * i, val + 64 * i, 32);
            en[i] = en[2 * i];
        }
    free(val);
    free(en);
}

static void emit_header_proof(sha256_cursor_t *c, const uint8_t *leaf_msg, size_t leaf_len, const uint8_t aunts[4][32],
                              unsigned index) {
    uint8_t h[32], n[32];
    sha256_emit(c, leaf_msg, leaf_len, h);
    for (int i = 0; i < 4; i++) {
        if ((index >> i) & 1)
            emit_inner(c, aunts[i], h, n);
        else
            emit_inner(c, h, aunts[i], n);
        memcpy(h, n, 32);
    }
}

static void build_sha256(trace_t *t, const tmx_offchain_head *h, const tmx_validator *vals, const tmx_hash_field *tf) {
    sha256_cursor_t c = {t, 0};
    const size_t n = h->n_max;
    uint8_t(*pk)[32] = (uint8_t(*)[32])malloc(n * 32);
    uint64_t *power = (uint64_t *)malloc(n * 8);
    uint32_t *blen = (uint32_t *)malloc(n * 4);
    if (h->kind == TMX_KIND_SKIP) {
        for (size_t i = 0; i < n; i++) {
            memcpy(pk[i], tf[i].pubkey, 32);
            power[i] = tf[i].voting_power;
            blen[i] = tf[i].validator_byte_length;
        }
        emit_validator_set(&c, (const uint8_t(*)[32])pk, power, blen, n, h->nb_trusted);
    }
    for (size_t i = 0; i < n; i++) {
        memcpy(pk[i], vals[i].pubkey, 32);
        power[i] = vals[i].voting_power;
        blen[i] = vals[i].validator_byte_length;
    }
    emit_validator_set(&c, (const uint8_t(*)[32])pk, power, blen, n, h->nb_validators);
    free(pk); free(power); free(blen);

    uint8_t m[80];
    /* leaf messages: 0x00 || protobuf bytes */
    uint8_t chain_msg[64] = {0}, height_msg[64] = {0};
    memcpy(chain_msg + 1, h->chain_id_proof.chain_id, 52);
    size_t chain_len = (size_t)h->chain_id_proof.enc_chain_id_byte_length + 1;
    if (chain_len > 55) chain_len = 55;
    height_msg[1] = 0x08;
    tm_marshal_int64_varint(h->height_proof.height, height_msg + 2);
    size_t height_len = (size_t)h->height_proof.enc_height_byte_length + 1;
    if (height_len > 55) height_len = 55;
    if (h->kind == TMX_KIND_SKIP) {
        m[0] = 0; memcpy(m + 1, h->aux_hash_proof.leaf, 34);
        emit_header_proof(&c, m, 35, h->aux_hash_proof.aunts, TMX_VALIDATORS_HASH_INDEX);
    }
    m[0] = 0; memcpy(m + 1, h->validators_hash_proof.leaf, 34);
    emit_header_proof(&c, m, 35, h->validators_hash_proof.aunts, TMX_VALIDATORS_HASH_INDEX);
    emit_header_proof(&c, chain_msg, chain_len, h->chain_id_proof.aunts, TMX_CHAIN_ID_INDEX);
    emit_header_proof(&c, height_msg, height_len, h->height_proof.aunts, TMX_BLOCK_HEIGHT_INDEX);
    if (h->kind == TMX_KIND_STEP) {
        m[0] = 0; memcpy(m + 1, h->last_block_id_proof.leaf, 72);
        emit_header_proof(&c, m, 73, h->last_block_id_proof.aunts, TMX_LAST_BLOCK_ID_INDEX);
        m[0] = 0; memcpy(m + 1, h->aux_hash_proof.leaf, 34);
        emit_header_proof(&c, m, 35, h->aux_hash_proof.aunts, TMX_NEXT_VALIDATORS_HASH_INDEX);
    }
    /* padding: compressions of the zero block from the IV */
    uint8_t zero[64] = {0};
    uint32_t tmp[8];
    while ((c.chunk + 1) * S256_ROUNDS <= t->n_rows) {
        sha256_rows(t, c.chunk * S256_ROUNDS, SHA256_IV, zero, tmp);
        c.chunk++;
    }
}

/* ------------------------------------------------------------------ SHA-512 rows */
static void put_halves(trace_t *t, int col, size_t row, uint64_t v) {
    CELL(t, col, row) = (uint32_t)v;
    CELL(t, col + 1, row) = v >> 32;
}

/* rows [row0, row0 + nrows) of one compression: 80 rounds, then 48 continuation rounds with round constant 0
 * (include/tmx_trace.h); nrows < 128 only for a truncated last padding chunk */
static void sha512_rows(trace_t *t, size_t row0, size_t nrows, const uint64_t cv[8], const uint8_t blk[128], uint64_t out_state[8],
                        int two) {
    sha512_round_t r[S512_ROWS_PER_CHUNK];
    uint64_t st[8];
    memcpy(st, cv, sizeof st);
    sha512_compress(st, blk, r);
    if (out_state) memcpy(out_state, st, sizeof st);
    uint64_t W[S512_ROWS_PER_CHUNK];
    for (int i = 0; i < 80; i++) W[i] = r[i].w;
    for (int i = 80; i < S512_ROWS_PER_CHUNK; i++) {
        uint64_t s0 = ror64(W[i - 15], 1) ^ ror64(W[i - 15], 8) ^ (W[i - 15] >> 7);
        uint64_t s1 = ror64(W[i - 2], 19) ^ ror64(W[i - 2], 61) ^ (W[i - 2] >> 6);
        W[i] = W[i - 16] + s0 + W[i - 7] + s1;
    }
    {   /* working variables entering rounds 80..127 */
        uint64_t v[8];
        memcpy(v, r[79].v, sizeof v);
        for (int i = 79; i < S512_ROWS_PER_CHUNK - 1; i++) {
            uint64_t a = v[0], bb = v[1], c = v[2], d = v[3], e = v[4], f = v[5], g = v[6], hh = v[7];
            uint64_t S1 = ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41), ch = (e & f) ^ (~e & g);
            uint64_t S0 = ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39), mj = (a & bb) ^ (a & c) ^ (bb & c);
            uint64_t t1 = hh + S1 + ch + (i < 80 ? SHA512_K[i] : 0) + W[i], t2 = S0 + mj;
            v[7] = g; v[6] = f; v[5] = e; v[4] = d + t1; v[3] = c; v[2] = bb; v[1] = a; v[0] = t1 + t2;
            memcpy(r[i + 1].v, v, sizeof v);
            r[i + 1].w = W[i + 1];
        }
    }
    for (size_t i = 0; i < nrows; i++) {
        size_t row = row0 + i;
        const uint64_t *v = r[i].v;
        const int bit_groups[6][2] = {{S512_A, 0}, {S512_B, 1}, {S512_C, 2}, {S512_E, 4}, {S512_F, 5}, {S512_G, 6}};
        for (int g = 0; g < 6; g++)
            for (int b = 0; b < 64; b++) CELL(t, bit_groups[g][0] + b, row) = (v[bit_groups[g][1]] >> b) & 1;
        put_halves(t, S512_D, row, v[3]);
        put_halves(t, S512_H, row, v[7]);
        uint64_t a = v[0], bb = v[1], c = v[2], d = v[3], e = v[4], f = v[5], g = v[6], hh = v[7];
        uint64_t S1 = ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41), ch = (e & f) ^ (~e & g);
        uint64_t S0 = ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39), mj = (a & bb) ^ (a & c) ^ (bb & c);
        uint64_t K = i < 80 ? SHA512_K[i] : 0, w = W[i];
#define LO(x) ((uint64_t)(uint32_t)(x))
#define HI(x) ((uint64_t)((x) >> 32))
        uint64_t t1lo = LO(hh) + LO(S1) + LO(ch) + LO(K) + LO(w);
        uint64_t t1hi = HI(hh) + HI(S1) + HI(ch) + HI(K) + HI(w);
        uint64_t salo = t1lo + LO(S0) + LO(mj), calo = salo >> 32;
        uint64_t sahi = t1hi + HI(S0) + HI(mj) + calo, cahi = sahi >> 32;
        uint64_t an = LO(salo) | (LO(sahi) << 32);
        uint64_t selo = LO(d) + t1lo, celo = selo >> 32;
        uint64_t sehi = HI(d) + t1hi + celo, cehi = sehi >> 32;
        uint64_t en = LO(selo) | (LO(sehi) << 32);
        for (int b = 0; b < 64; b++) {
            CELL(t, S512_AN + b, row) = (an >> b) & 1;
            CELL(t, S512_EN + b, row) = (en >> b) & 1;
        }
        for (int b = 0; b < 3; b++) {
            CELL(t, S512_CA + b, row) = (calo >> b) & 1;
            CELL(t, S512_CA + 3 + b, row) = (cahi >> b) & 1;
            CELL(t, S512_CE + b, row) = (celo >> b) & 1;
            CELL(t, S512_CE + 3 + b, row) = (cehi >> b) & 1;
        }
        for (int j = 0; j < 16; j++) {
            int idx = (int)i - 15 + j;
            put_halves(t, S512_W + 2 * j, row, idx >= 0 ? W[idx] : 0);
        }
        uint64_t w14 = i >= 1 ? W[i - 1] : 0, w1 = i >= 14 ? W[i - 14] : 0;
        for (int b = 0; b < 64; b++) {
            CELL(t, S512_WB14 + b, row) = (w14 >> b) & 1;
            CELL(t, S512_WB1 + b, row) = (w1 >> b) & 1;
        }
        for (int j = 0; j < 8; j++) put_halves(t, S512_CV + 2 * j, row, cv[j]);
        {   /* schedule sum of the row's window: sigma1(w[14]) + w[9] + sigma0(w[1]) + w[0], zeros before the block starts */
            uint64_t x = w14, y = w1, w9 = i >= 6 ? W[i - 6] : 0, w0 = i >= 15 ? W[i - 15] : 0;
            uint64_t s1 = ror64(x, 19) ^ ror64(x, 61) ^ (x >> 6), s0 = ror64(y, 1) ^ ror64(y, 8) ^ (y >> 7);
            uint64_t lo = LO(s1) + LO(w9) + LO(s0) + LO(w0);
            uint64_t cwlo = lo >> 32;
            uint64_t hi = HI(s1) + HI(w9) + HI(s0) + HI(w0) + cwlo;
            uint64_t cwhi = hi >> 32;
            CELL(t, S512_CW, row) = cwlo & 1;
            CELL(t, S512_CW + 1, row) = (cwlo >> 1) & 1;
            CELL(t, S512_CW + 2, row) = cwhi & 1;
            CELL(t, S512_CW + 3, row) = (cwhi >> 1) & 1;
            CELL(t, S512_WS, row) = LO(lo);
            CELL(t, S512_WS + 1, row) = LO(hi);
        }
        for (int j = 0; j < 8; j++) {
            uint64_t dlo = 0, dhi = 0, clo = 0, chi = 0;
            if (i >= 79) { /* the digest stays on the rows 79..127 so that the next chunk can chain from it; carries on row 79 only */
                const uint64_t *v79 = r[79].v;
                uint64_t a79 = v79[0], b79 = v79[1], c79 = v79[2], d79 = v79[3], e79 = v79[4], f79 = v79[5], g79 = v79[6], h79 = v79[7];
                uint64_t t1 = h79 + (ror64(e79, 14) ^ ror64(e79, 18) ^ ror64(e79, 41)) + ((e79 & f79) ^ (~e79 & g79)) + SHA512_K[79] + W[79];
                uint64_t t2 = (ror64(a79, 28) ^ ror64(a79, 34) ^ ror64(a79, 39)) + ((a79 & b79) ^ (a79 & c79) ^ (b79 & c79));
                const uint64_t fin[8] = {t1 + t2, a79, b79, c79, d79 + t1, e79, f79, g79};
                uint64_t lo = LO(cv[j]) + LO(fin[j]);
                uint64_t hi = HI(cv[j]) + HI(fin[j]) + (lo >> 32);
                if (i == 79) {
                    clo = lo >> 32;
                    chi = hi >> 32;
                }
                dlo = LO(lo);
                dhi = LO(hi);
            }
            CELL(t, S512_DG + 2 * j, row) = dlo;
            CELL(t, S512_DG + 2 * j + 1, row) = dhi;
            CELL(t, S512_DC + 2 * j, row) = clo;
            CELL(t, S512_DC + 2 * j + 1, row) = chi;
        }
        CELL(t, S512_TWO, row) = (uint64_t)(two != 0);
#undef LO
#undef HI
    }
}

/* the (pk, sig, msg, len) triple the Ed25519 gadget actually verifies for slot i */
static void effective_triple(const tmx_validator *v, uint8_t pk[32], uint8_t sig[64], uint8_t msg[124], size_t *len) {
    if (v->is_signed) {
        memcpy(pk, v->pubkey, 32);
        memcpy(sig, v->sig_r, 32);
        memcpy(sig + 32, v->sig_s, 32);
        memcpy(msg, v->message, 124);
        *len = v->message_byte_length > 124 ? 124 : v->message_byte_length;
    } else {
        memcpy(pk, TMX_DUMMY_PUBLIC_KEY, 32);
        memcpy(sig, TMX_DUMMY_SIGNATURE, 64);
        memset(msg, 0, 124);
        *len = 32;
    }
}

static void build_sha512(trace_t *t, const tmx_offchain_head *h, const tmx_validator *vals, uint8_t (*hdigest)[64]) {
    const uint8_t zero[128] = {0};
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t i = 0; i < h->n_max; i++) {
        uint8_t pk[32], sig[64], msg[124], m[64 + 124], buf[128 * 3];
        size_t len;
        effective_triple(&vals[i], pk, sig, msg, &len);
        memcpy(m, sig, 32);
        memcpy(m + 32, pk, 32);
        memcpy(m + 64, msg, len);
        size_t nb = sha512_pad(m, 64 + len, buf);
        uint64_t st[8], nxt[8];
        memcpy(st, SHA512_IV, sizeof st);
        size_t row = i * S512_ROWS_PER_VALIDATOR;
        for (size_t b = 0; b < 2; b++) {
            if (b < nb) {
                sha512_rows(t, row + S512_ROWS_PER_CHUNK * b, S512_ROWS_PER_CHUNK, st, buf + 128 * b, nxt, nb == 2);
                memcpy(st, nxt, sizeof st);
            } else
                sha512_rows(t, row + S512_ROWS_PER_CHUNK * b, S512_ROWS_PER_CHUNK, SHA512_IV, zero, NULL, 0); /* unused second slot */
        }
        for (int k = 0; k < 8; k++)
            for (int j = 0; j < 8; j++) hdigest[i][8 * k + j] = (uint8_t)(st[k] >> (56 - 8 * j));
    }
    for (size_t row = (size_t)h->n_max * S512_ROWS_PER_VALIDATOR; row < t->n_rows; row += S512_ROWS_PER_CHUNK) {
        size_t nr = t->n_rows - row < S512_ROWS_PER_CHUNK ? t->n_rows - row : S512_ROWS_PER_CHUNK;
        sha512_rows(t, row, nr, SHA512_IV, zero, NULL, 0);
    }
}

/* ------------------------------------------------------------------ Ed25519 rows */
static void put_fe(trace_t *t, int col, size_t row, const fe_t *a) {
    for (int i = 0; i < 16; i++) CELL(t, col + i, row) = (uint64_t)a->l[i];
}

/* rows of one validator slot: acc' = 2 acc + T[bs + 2 bh], most significant scalar bits first (include/tmx_trace.h) */
static void ed_slot_rows(trace_t *t, size_t row0, const uint8_t s[32], const uint8_t h[32], const ge_cached_t T[4]) {
    fe_t acc[3], nxt[3];
    fe_mul_witness_t wit[ED_N_MUL];
    memset(acc, 0, sizeof acc);
    acc[1].l[0] = 1;
    acc[2].l[0] = 1;
    uint64_t sacc_s = 0, sacc_h = 0;
    for (int r = 0; r < ED_ROWS_PER_VALIDATOR; r++) {
        const size_t row = row0 + r;
        const int j = 255 - r;
        const int bs = (s[j >> 3] >> (j & 7)) & 1, bh = (h[j >> 3] >> (j & 7)) & 1;
        const ge_cached_t *add = &T[bs + 2 * bh];
        if ((r & 15) == 0) sacc_s = sacc_h = 0;
        CELL(t, ED_BS, row) = bs;
        CELL(t, ED_BH, row) = bh;
        CELL(t, ED_SACC_S, row) = sacc_s;
        CELL(t, ED_SACC_H, row) = sacc_h;
        sacc_s = 2 * sacc_s + bs;
        sacc_h = 2 * sacc_h + bh;
        for (int co = 0; co < 3; co++) put_fe(t, ED_ACC + 16 * co, row, &acc[co]);
        for (int i = 0; i < 16; i++) {
            CELL(t, ED_ADD + i, row) = (uint64_t)add->ypx[i];
            CELL(t, ED_ADD + 16 + i, row) = (uint64_t)add->ymx[i];
            CELL(t, ED_ADD + 32 + i, row) = (uint64_t)add->t2d[i];
        }
        ge_straus_row(acc, add, nxt, wit);
        for (int m = 0; m < ED_N_MUL; m++) {
            const int base = ED_MUL + m * ED_MUL_STRIDE;
            for (int k = 0; k < 16; k++) CELL(t, base + k, row) = (uint64_t)wit[m].c[k];
            for (int k = 0; k < 17; k++) CELL(t, base + ED_MUL_Q + k, row) = (uint64_t)wit[m].q[k];
            /* the limb equations are checked in pairs, so only the carries out of the odd limbs are committed, split into a
             * 16-bit and an 11-bit part (both range checked on the bus) */
            for (int k = 0; k < ED_MUL_NW; k++) {
                const int64_t w = wit[m].w[2 * k + 1] + ED_W_OFFSET;
                if (w < 0 || w >= ((int64_t)1 << 27)) {
                    fprintf(stderr, "oracle: Ed25519 carry out of range\n");
                    abort();
                }
                CELL(t, base + ED_MUL_WLO + k, row) = (uint64_t)(w & 0xFFFF);
                CELL(t, base + ED_MUL_WHI + k, row) = (uint64_t)(w >> 16);
            }
        }
        memcpy(acc, nxt, sizeof acc);
    }
}

static int build_ed25519(trace_t *t, const tmx_offchain_head *h, const tmx_validator *vals, uint8_t (*hdigest)[64]) {
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
    for (size_t i = 0; i < h->n_max; i++) {
        uint8_t pk[32], sig[64], msg[124], hs[32];
        size_t len;
        effective_triple(&vals[i], pk, sig, msg, &len);
        ge_t A;
        if (ge_decompress(&A, pk) != 0) {
            bad |= 1;
            continue;
        }
        sc_reduce512(hs, hdigest[i]);
        ge_cached_t T[4];
        fe_t xD, yD;
        ge_straus_table(&A, T, &xD, &yD);
        ed_slot_rows(t, i * ED_ROWS_PER_VALIDATOR, sig + 32, hs, T);
    }
    const size_t first = (size_t)h->n_max * ED_ROWS_PER_VALIDATOR;
    if (first < t->n_rows) {
        /* padding slots: all scalar bits zero, addend O in every row (valid rows; nothing of them reaches the bus); compute
         * one slot and copy */
        uint8_t zero[32] = {0};
        ge_cached_t T[4];
        memset(T, 0, sizeof T);
        T[0].ypx[0] = 1;
        T[0].ymx[0] = 1;
        ed_slot_rows(t, first, zero, zero, T);
        for (size_t row = first + ED_ROWS_PER_VALIDATOR; row < t->n_rows; row += ED_ROWS_PER_VALIDATOR)
            for (size_t c = 0; c < t->n_cols; c++)
                memcpy(&CELL(t, c, row), &CELL(t, c, first), ED_ROWS_PER_VALIDATOR * sizeof(uint64_t));
    }
    return bad;
}

/* Allocates and fills the three tables.  Returns 0, or TMX_CHECK_SIGNATURE if a key does not decompress. */
int tm_build_traces(const uint8_t *blob, size_t blob_len, trace_t out[3]) {
    if (blob_len < sizeof(tmx_offchain_head)) return TMX_CHECK_INPUT;
    const tmx_offchain_head *h = (const tmx_offchain_head *)blob;
    if (h->magic != TMX_BLOB_MAGIC || blob_len != TMX_BLOB_SIZE(h->kind, h->n_max)) return TMX_CHECK_INPUT;
    const tmx_validator *vals = (const tmx_validator *)(blob + sizeof(tmx_offchain_head));
    const tmx_hash_field *tf = (const tmx_hash_field *)(vals + h->n_max);
    size_t dims[6];
    tm_trace_dims(h->kind, h->n_max, dims);
    for (int k = 0; k < 3; k++) {
        out[k].n_rows = dims[2 * k];
        out[k].n_cols = dims[2 * k + 1];
        out[k].data = (uint64_t *)calloc(out[k].n_rows * out[k].n_cols, sizeof(uint64_t));
    }
    {
        ge_t tmp;
        ge_identity(&tmp);
    }
    uint8_t(*hd)[64] = (uint8_t(*)[64])malloc((size_t)h->n_max * 64);
    build_sha256(&out[0], h, vals, tf);
    build_sha512(&out[1], h, vals, hd);
    int bad = build_ed25519(&out[2], h, vals, hd);
    free(hd);
    return bad ? TMX_CHECK_SIGNATURE : TMX_CHECK_OK;
}

void tm_free_traces(trace_t out[3]) {
    for (int k = 0; k < 3; k++) {
        free(out[k].data);
        out[k].data = NULL;
    }
}
