/*
 * oracle/stark.c -- deterministic CPU prover and verifier for the three-table STARK that witnesses one skip /
 * step proof: trace commitment (PolynomialBatch), constraint quotient, openings at zeta and g*zeta, FRI
 * (arity 16, final polynomial, 16-bit proof of work with the MINIMUM witness, 84 queries).
 * TEST INFRASTRUCTURE ONLY (see oracle/gl.h header); also the CPU baseline that bench.py times.
 *
 * Restates the plonky2 0.2.0 / starky pipeline the reference reaches through `circuit.prove()` /
 * `circuit.verify()` [REF circuits/skip.rs:214,244,247; circuits/step.rs:196,223,226] -- fri/oracle.rs
 * (prove_openings), fri/prover.rs (fri_committed_trees, fri_proof_of_work, query rounds), fri/verifier.rs,
 * starky prover.rs (quotient polynomials, opening set) -- with Curta's STARK configuration (rate_bits 1,
 * cap height 4, 84 queries).  The dependency sources are absent and the reference pins no proof bytes:
 * PARITY UNPINNED at the proof level; the primitives underneath are KAT-pinned.
 * The FRI commit phase deliberately works in COEFFICIENT space like plonky2 (fold coefficients, re-evaluate with
 * a coset FFT), while the CUDA product folds evaluations; agreement of the two is part of the parity test.
 */
#include "oracle.h"
#include "oracle_w.h"
#include "../include/tmx_trace.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <omp.h>

#define RATE_BITS 1
#define CAP_HEIGHT 4
#define NUM_CHALLENGES 2
#define POW_BITS 16
#define NUM_QUERIES 84
#define ARITY_BITS 4
#define FINAL_POLY_BITS 5
#define QDF 2 /* quotient degree factor */
#define N_QUOT (NUM_CHALLENGES * QDF)
#define PROOF_MAGIC 0x50584D54ULL /* "TMXP" */

typedef struct {
    size_t n_rows, n_cols;
    uint64_t *data;
} trace_t;
int tm_build_traces(const uint8_t *blob, size_t blob_len, trace_t out[3]);
void tm_free_traces(trace_t out[3]);
void tm_trace_dims(uint32_t kind, uint32_t n_max, size_t dims[6]);

/* ------------------------------------------------------------------ AIR instantiations */
#define FT gl_t
#define F_ADD gl_add
#define F_SUB gl_sub
#define F_MUL gl_mul
#define F_C(x) ((gl_t)(x))
#define SUF(name) name##_b
#include "air.inc"
#undef FT
#undef F_ADD
#undef F_SUB
#undef F_MUL
#undef F_C
#undef SUF
#define FT gl2_t
#define F_ADD gl2_add
#define F_SUB gl2_sub
#define F_MUL gl2_mul
#define F_C(x) gl2_from((gl_t)(x))
#define SUF(name) name##_e
#include "air.inc"
#undef FT
#undef F_ADD
#undef F_SUB
#undef F_MUL
#undef F_C
#undef SUF

enum { T_SHA256 = 0, T_SHA512 = 1, T_ED = 2, N_TABLES = 3 };
static const int TABLE_COLS[3] = {S256_COLS, S512_COLS, ED_COLS};
static const int TABLE_NPER[3] = {6, 11, 3};
/* The SHA-256 table's last two "periodic" columns are not periodic: which chunk starts a message (chaining value = IV)
 * and which chunk continues one (chaining value = previous digest) is fixed by the circuit shape (kind, n_max), see
 * sha256_chunk_continues().  They are public columns of full length: the prover evaluates them on the LDE coset, the
 * verifier evaluates their interpolant at zeta itself.  So the "period" of that table is its length. */
static __thread uint32_t g_kind, g_n_max; /* shape of the circuit being proved / verified (set by the entry points) */
static size_t table_period(int table, size_t n) {
    return table == 0 ? n : (table == 1 ? S512_ROWS_PER_VALIDATOR : ED_ROWS_PER_VALIDATOR);
}
/* Does 64-row chunk c of the SHA-256 table continue the message of chunk c - 1?  Layout (oracle/trace.c build_sha256):
 * per validator set n_max one-chunk leaf hashes, then np - 1 two-chunk inner nodes; then the header proofs, each a leaf
 * (two chunks for the 72-byte last-block-id leaf of the step circuit, else one) and four two-chunk inner nodes; then
 * one-chunk padding messages. */
static int sha256_chunk_continues(uint32_t kind, uint32_t n_max, size_t c) {
    size_t np = 1;
    while (np < n_max) np *= 2;
    const size_t set_chunks = n_max + 2 * (np - 1), sets = kind == TMX_KIND_SKIP ? 2 : 1;
    if (c < sets * set_chunks) {
        const size_t local = c % set_chunks;
        return local >= n_max && ((local - n_max) & 1);
    }
    size_t h = c - sets * set_chunks;
    const int n_proofs = kind == TMX_KIND_SKIP ? 4 : 5;
    for (int k = 0; k < n_proofs; k++) {
        const size_t leaf = (kind == TMX_KIND_STEP && k == 3) ? 2 : 1, len = leaf + 8;
        if (h < len) return h < leaf ? h == 1 : ((h - leaf) & 1);
        h -= len;
    }
    return 0;
}

/* periodic pattern value of column pc at row r (r < period) */
static gl_t periodic_pattern(int table, int pc, size_t row) {
    const int r = (int)(row & 511);
    if (table == T_SHA256) {
        const int rr = (int)(row & 63);
        const size_t chunk = row >> 6;
        switch (pc) {
            case 0: return SHA256_K[rr];
            case 1: return rr == 63;
            case 2: return rr != 63;
            case 3: return rr >= 15 && rr <= 62;
            case 4: return rr == 0 && !sha256_chunk_continues(g_kind, g_n_max, chunk);      /* FIRST: chaining value = IV */
            default: return rr == 63 && sha256_chunk_continues(g_kind, g_n_max, chunk + 1); /* LINK: next chunk chains */
        }
    }
    if (table == T_SHA512) {
        const int rr = r % S512_ROWS_PER_CHUNK; /* row inside the chunk; r inside the validator's two-chunk slot */
        switch (pc) {
            case 0: return rr < 80 ? (uint32_t)SHA512_K[rr] : 0;
            case 1: return rr < 80 ? SHA512_K[rr] >> 32 : 0;
            case 2: return rr == 79;
            case 3: return rr != 79;
            case 4: return rr != S512_ROWS_PER_CHUNK - 1;
            case 5: return rr >= 15 && rr <= S512_ROWS_PER_CHUNK - 2;
            case 6: return r == 0;
            case 7: return rr < 79;
            case 8: return rr >= 79 && rr <= S512_ROWS_PER_CHUNK - 2;
            case 9: return r == S512_ROWS_PER_CHUNK - 1;
            default: return r != S512_ROWS_PER_VALIDATOR - 1;
        }
    }
    switch (pc) { /* Ed25519 */
        case 0: return (r & 255) != 255;
        case 1: return r == 0;
        default: return r == 256;
    }
}

/* ------------------------------------------------------------------ small utilities */
typedef struct {
    gl_t *v;
    size_t n, cap;
} wbuf_t;
static void wb_push(wbuf_t *b, gl_t x) {
    if (b->n == b->cap) {
        b->cap = b->cap ? 2 * b->cap : 1 << 16;
        b->v = (gl_t *)realloc(b->v, b->cap * sizeof(gl_t));
    }
    b->v[b->n++] = x;
}
static void wb_push_many(wbuf_t *b, const gl_t *x, size_t n) {
    for (size_t i = 0; i < n; i++) wb_push(b, x[i]);
}
static void wb_push_ext(wbuf_t *b, gl2_t x) {
    wb_push(b, x.a0);
    wb_push(b, x.a1);
}

typedef struct {
    const gl_t *v;
    size_t n, pos;
    int err;
} rbuf_t;
static gl_t rb_get(rbuf_t *r) {
    if (r->pos >= r->n) {
        r->err = 1;
        return 0;
    }
    return r->v[r->pos++];
}
static const gl_t *rb_take(rbuf_t *r, size_t n) {
    if (r->pos + n > r->n) {
        r->err = 1;
        return NULL;
    }
    const gl_t *p = r->v + r->pos;
    r->pos += n;
    return p;
}
static gl2_t rb_get_ext(rbuf_t *r) {
    gl_t a = rb_get(r), b = rb_get(r);
    return gl2_make(a, b);
}

static void observe_cap(challenger_t *ch, const gl_t *cap, size_t n_digests) { challenger_observe_many(ch, cap, 4 * n_digests); }
static void observe_ext(challenger_t *ch, gl2_t x) {
    challenger_observe(ch, x.a0);
    challenger_observe(ch, x.a1);
}

static size_t fri_num_layers(unsigned degree_bits) {
    size_t l = 0;
    while (degree_bits > FINAL_POLY_BITS && degree_bits + RATE_BITS - ARITY_BITS >= CAP_HEIGHT) {
        l++;
        degree_bits -= ARITY_BITS;
    }
    return l;
}

/* circuit digest: binds kind, sizes, chain id, skip_max and the protocol parameters */
static void circuit_digest(uint32_t kind, uint32_t n_max, const uint8_t *chain_id, size_t chain_id_len, uint64_t skip_max,
                           gl_t out[4]) {
    gl_t in[96];
    size_t k = 0;
    in[k++] = PROOF_MAGIC; in[k++] = kind; in[k++] = n_max; in[k++] = skip_max;
    in[k++] = RATE_BITS; in[k++] = CAP_HEIGHT; in[k++] = NUM_CHALLENGES; in[k++] = POW_BITS; in[k++] = NUM_QUERIES;
    in[k++] = ARITY_BITS; in[k++] = FINAL_POLY_BITS;
    size_t dims[6];
    tm_trace_dims(kind, n_max, dims);
    for (int i = 0; i < 6; i++) in[k++] = dims[i];
    in[k++] = chain_id_len;
    for (size_t i = 0; i < chain_id_len && i < 64; i++) in[k++] = chain_id[i];
    poseidon_hash_no_pad(in, k, out);
}

static void ext_poly_fft(gl2_t *a, size_t n, gl_t shift) { /* coset evaluation, natural order, componentwise */
    gl_t *t = (gl_t *)malloc(n * sizeof(gl_t));
    for (int comp = 0; comp < 2; comp++) {
        for (size_t i = 0; i < n; i++) t[i] = comp ? a[i].a1 : a[i].a0;
        ntt_coset_forward(t, n, shift);
        for (size_t i = 0; i < n; i++)
            if (comp) a[i].a1 = t[i]; else a[i].a0 = t[i];
    }
    free(t);
}

static gl2_t ext_poly_eval(const gl2_t *c, size_t n, gl2_t x) {
    gl2_t acc = gl2_from(0);
    for (size_t i = n; i-- > 0;) acc = gl2_add(gl2_mul(acc, x), c[i]);
    return acc;
}
static gl2_t base_poly_eval(const gl_t *c, size_t n, gl2_t x) {
    gl2_t acc = gl2_from(0);
    for (size_t i = n; i-- > 0;) acc = gl2_add(gl2_mul(acc, x), gl2_from(c[i]));
    return acc;
}

/* interpolate the 16 points (xs[i], ys[i]) and evaluate at beta */
static gl2_t interpolate16(const gl_t xs[16], const gl2_t ys[16], gl2_t beta) {
    gl2_t acc = gl2_from(0);
    for (int i = 0; i < 16; i++) {
        gl2_t num = gl2_from(1);
        gl_t den = 1;
        for (int j = 0; j < 16; j++)
            if (j != i) {
                num = gl2_mul(num, gl2_sub(beta, gl2_from(xs[j])));
                den = gl_mul(den, gl_sub(xs[i], xs[j]));
            }
        acc = gl2_add(acc, gl2_mul(ys[i], gl2_scale(num, gl_inv(den))));
    }
    return acc;
}

/* plonky2 compute_evaluation: evals are the 16 leaf values (bit-reversed order inside the coset) */
static gl2_t fri_fold_coset(gl_t x, unsigned idx_in_coset, const gl2_t evals[16], gl2_t beta) {
    gl_t g = gl_root_of_unity(ARITY_BITS);
    gl2_t ys[16];
    gl_t xs[16];
    for (unsigned i = 0; i < 16; i++) ys[tmx_bitrev(i, ARITY_BITS)] = evals[i];
    unsigned rev = (unsigned)tmx_bitrev(idx_in_coset, ARITY_BITS);
    gl_t start = gl_mul(x, gl_pow(g, 16 - rev));
    gl_t cur = start;
    for (int i = 0; i < 16; i++) {
        xs[i] = cur;
        cur = gl_mul(cur, g);
    }
    return interpolate16(xs, ys, beta);
}

typedef struct {
    gl_t acc[NUM_CHALLENGES];
    gl_t alpha[NUM_CHALLENGES];
} acc_b_t;
static void emit_b(void *ctx, gl_t c) {
    acc_b_t *a = (acc_b_t *)ctx;
    for (int i = 0; i < NUM_CHALLENGES; i++) a->acc[i] = gl_add(gl_mul(a->acc[i], a->alpha[i]), c);
}
typedef struct {
    gl2_t acc[NUM_CHALLENGES];
    gl2_t alpha[NUM_CHALLENGES];
} acc_e_t;
static void emit_e(void *ctx, gl2_t c) {
    acc_e_t *a = (acc_e_t *)ctx;
    for (int i = 0; i < NUM_CHALLENGES; i++) a->acc[i] = gl2_add(gl2_mul(a->acc[i], a->alpha[i]), c);
}
static void air_eval_b(int table, const gl_t *l, const gl_t *n, const gl_t *per, acc_b_t *a) {
    if (table == T_SHA256) air_sha256_b(l, n, per, emit_b, a);
    else if (table == T_SHA512) air_sha512_b(l, n, per, emit_b, a);
    else air_ed25519_b(l, n, per, emit_b, a);
}
static void air_eval_e(int table, const gl2_t *l, const gl2_t *n, const gl2_t *per, acc_e_t *a) {
    if (table == T_SHA256) air_sha256_e(l, n, per, emit_e, a);
    else if (table == T_SHA512) air_sha512_e(l, n, per, emit_e, a);
    else air_ed25519_e(l, n, per, emit_e, a);
}

/* ------------------------------------------------------------------ prover: one table */
static void merkle_open(wbuf_t *w, const merkle_tree_t *t, size_t idx) {
    gl_t sib[4 * 40];
    size_t k = merkle_prove(t, idx, sib);
    wb_push_many(w, sib, 4 * k);
}

static double t_last;
static void tick(const char *what) {
    if (!getenv("TMX_ORACLE_TIMING")) return;
    double t = omp_get_wtime();
    fprintf(stderr, "  [oracle] %-28s %.2f s\n", what, t - t_last);
    t_last = t;
}

static void prove_table(int table, const trace_t *tr, challenger_t *ch, wbuf_t *w) {
    t_last = omp_get_wtime();
    const size_t n = tr->n_rows, C = tr->n_cols, m = n << RATE_BITS;
    const unsigned k = tmx_log2(n), km = k + RATE_BITS;
    /* 1. trace commitment */
    gl_t *lde = (gl_t *)malloc(C * m * sizeof(gl_t));
    gl_t *coeffs = (gl_t *)malloc(C * n * sizeof(gl_t));
    ntt_lde_batch(tr->data, C, n, RATE_BITS, lde, coeffs);
    tick("lde");
    merkle_tree_t tree_t;
    commit_columns(&tree_t, lde, C, m, CAP_HEIGHT);
    const size_t cap_n = (size_t)1 << tree_t.cap_height;
    wb_push_many(w, tree_t.cap, 4 * cap_n);
    observe_cap(ch, tree_t.cap, cap_n);
    tick("trace merkle");
    /* 2. constraint challenges */
    gl_t alpha[NUM_CHALLENGES];
    for (int i = 0; i < NUM_CHALLENGES; i++) alpha[i] = challenger_get(ch);
    /* 3. quotient on the LDE coset */
    const int nper = TABLE_NPER[table];
    const size_t P = table_period(table, n);
    gl_t *pertab = NULL; /* [nper][2P] values at natural LDE index mod 2P */
    if (nper) {
        pertab = (gl_t *)malloc((size_t)nper * 2 * P * sizeof(gl_t));
        gl_t shift = gl_pow(GL_GENERATOR, n / P);
        for (int pc = 0; pc < nper; pc++) {
            gl_t *t = pertab + (size_t)pc * 2 * P;
            for (size_t r = 0; r < P; r++) t[r] = periodic_pattern(table, pc, r);
            ntt_inverse(t, P);
            memset(t + P, 0, P * sizeof(gl_t));
            ntt_coset_forward(t, 2 * P, shift);
        }
    }
    gl_t *qv = (gl_t *)malloc((size_t)NUM_CHALLENGES * m * sizeof(gl_t)); /* natural order */
    const gl_t gn = gl_pow(GL_GENERATOR, n);
    const gl_t zh_inv[2] = {gl_inv(gl_sub(gn, 1)), gl_inv(gl_sub(gl_neg(gn), 1))};
#pragma omp parallel
    {
        gl_t *loc = (gl_t *)malloc(C * sizeof(gl_t)), *nxt = (gl_t *)malloc(C * sizeof(gl_t));
        gl_t per[16];
#pragma omp for schedule(static)
        for (size_t j = 0; j < m; j++) {
            size_t p = tmx_bitrev(j, km), p2 = tmx_bitrev((j + (1u << RATE_BITS)) & (m - 1), km);
            for (size_t c = 0; c < C; c++) {
                loc[c] = lde[c * m + p];
                nxt[c] = lde[c * m + p2];
            }
            for (int pc = 0; pc < nper; pc++) per[pc] = pertab[(size_t)pc * 2 * P + (j & (2 * P - 1))];
            acc_b_t a;
            for (int i = 0; i < NUM_CHALLENGES; i++) {
                a.acc[i] = 0;
                a.alpha[i] = alpha[i];
            }
            air_eval_b(table, loc, nxt, per, &a);
            for (int i = 0; i < NUM_CHALLENGES; i++) qv[(size_t)i * m + j] = gl_mul(a.acc[i], zh_inv[j & 1]);
        }
        free(loc);
        free(nxt);
    }
    tick("quotient eval");
    /* quotient chunks: coefficients of degree < 2n split in two */
    gl_t *qcoef = (gl_t *)malloc((size_t)N_QUOT * n * sizeof(gl_t));
    for (int i = 0; i < NUM_CHALLENGES; i++) {
        ntt_coset_inverse(qv + (size_t)i * m, m, GL_GENERATOR);
        memcpy(qcoef + (size_t)(QDF * i) * n, qv + (size_t)i * m, m * sizeof(gl_t)); /* m = QDF * n */
    }
    gl_t *qlde = (gl_t *)malloc((size_t)N_QUOT * m * sizeof(gl_t));
    for (int q = 0; q < N_QUOT; q++) {
        gl_t *buf = qv; /* reuse */
        memcpy(buf, qcoef + (size_t)q * n, n * sizeof(gl_t));
        memset(buf + n, 0, (m - n) * sizeof(gl_t));
        ntt_coset_forward(buf, m, GL_GENERATOR);
        for (size_t j = 0; j < m; j++) qlde[(size_t)q * m + j] = buf[tmx_bitrev(j, km)];
    }
    merkle_tree_t tree_q;
    commit_columns(&tree_q, qlde, N_QUOT, m, CAP_HEIGHT);
    wb_push_many(w, tree_q.cap, 4 * cap_n);
    observe_cap(ch, tree_q.cap, cap_n);
    tick("quotient commit");
    /* 4. openings */
    const gl2_t zeta = challenger_get_ext(ch);
    const gl2_t zeta_next = gl2_scale(zeta, gl_root_of_unity(k));
    gl2_t *op_local = (gl2_t *)malloc(C * sizeof(gl2_t)), *op_next = (gl2_t *)malloc(C * sizeof(gl2_t));
    gl2_t op_quot[N_QUOT];
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t c = 0; c < C; c++) {
        op_local[c] = base_poly_eval(coeffs + c * n, n, zeta);
        op_next[c] = base_poly_eval(coeffs + c * n, n, zeta_next);
    }
    for (int q = 0; q < N_QUOT; q++) op_quot[q] = base_poly_eval(qcoef + (size_t)q * n, n, zeta);
    for (size_t c = 0; c < C; c++) wb_push_ext(w, op_local[c]);
    for (size_t c = 0; c < C; c++) wb_push_ext(w, op_next[c]);
    for (int q = 0; q < N_QUOT; q++) wb_push_ext(w, op_quot[q]);
    for (size_t c = 0; c < C; c++) observe_ext(ch, op_local[c]);
    for (int q = 0; q < N_QUOT; q++) observe_ext(ch, op_quot[q]);
    for (size_t c = 0; c < C; c++) observe_ext(ch, op_next[c]);
    tick("openings");
    /* 5. FRI batch polynomial, coefficient space (plonky2 fri/oracle.rs prove_openings) */
    const gl2_t fa = challenger_get_ext(ch);
    gl2_t *final_poly = (gl2_t *)calloc(m, sizeof(gl2_t));
    {
        gl2_t *comp = (gl2_t *)malloc(n * sizeof(gl2_t));
        for (int batch = 0; batch < 2; batch++) {
            const size_t npoly = batch == 0 ? C + N_QUOT : C;
            const gl2_t z = batch == 0 ? zeta : zeta_next;
            /* composition = sum_j alpha^j f_j (Horner from the last polynomial) */
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; i++) {
                gl2_t acc = gl2_from(0);
                for (size_t j = npoly; j-- > 0;) {
                    gl_t v = j < C ? coeffs[j * n + i] : qcoef[(j - C) * n + i];
                    acc = gl2_add(gl2_mul(acc, fa), gl2_from(v));
                }
                comp[i] = acc;
            }
            /* divide by (X - z), drop the remainder: b_{i-1} = a_i + z * b_i */
            gl2_t carry = gl2_from(0);
            for (size_t i = n; i-- > 0;) {
                gl2_t ai = comp[i];
                comp[i] = carry; /* coefficient of X^i of the quotient (top one is zero) */
                carry = gl2_add(ai, gl2_mul(z, carry));
            }
            /* final = final * alpha^npoly + quotient */
            gl2_t sh = gl2_pow(fa, npoly);
            for (size_t i = 0; i < n; i++) final_poly[i] = gl2_add(gl2_mul(final_poly[i], sh), comp[i]);
        }
        free(comp);
    }
    tick("fri batch poly");
    /* 6. FRI commit phase (fri_committed_trees) */
    const size_t n_layers = fri_num_layers(k);
    merkle_tree_t *layer_trees = (merkle_tree_t *)calloc(n_layers ? n_layers : 1, sizeof(merkle_tree_t));
    gl_t **layer_leaves = (gl_t **)calloc(n_layers ? n_layers : 1, sizeof(gl_t *));
    size_t cur_len = m; /* coefficient vector length (upper part zero) */
    gl2_t *cf = final_poly;
    gl2_t *vals = (gl2_t *)malloc(m * sizeof(gl2_t));
    memcpy(vals, cf, m * sizeof(gl2_t));
    gl_t shift = GL_GENERATOR;
    ext_poly_fft(vals, m, shift);
    for (size_t l = 0; l < n_layers; l++) {
        const unsigned lg = tmx_log2(cur_len);
        gl_t *leaves = (gl_t *)malloc(cur_len * 2 * sizeof(gl_t)); /* rows of 16 ext = 32 elements, bit-reversed order */
        for (size_t p = 0; p < cur_len; p++) {
            gl2_t v = vals[tmx_bitrev(p, lg)];
            leaves[2 * p] = v.a0;
            leaves[2 * p + 1] = v.a1;
        }
        merkle_build(&layer_trees[l], leaves, cur_len >> ARITY_BITS, 32, CAP_HEIGHT);
        layer_leaves[l] = leaves;
        const size_t lcap = (size_t)1 << layer_trees[l].cap_height;
        wb_push_many(w, layer_trees[l].cap, 4 * lcap);
        observe_cap(ch, layer_trees[l].cap, lcap);
        const gl2_t beta = challenger_get_ext(ch);
        size_t new_len = cur_len >> ARITY_BITS;
        for (size_t i = 0; i < new_len; i++) {
            gl2_t acc = gl2_from(0);
            for (int j = 15; j >= 0; j--) acc = gl2_add(gl2_mul(acc, beta), cf[16 * i + j]);
            cf[i] = acc;
        }
        cur_len = new_len;
        shift = gl_pow(shift, 16);
        memcpy(vals, cf, cur_len * sizeof(gl2_t));
        ext_poly_fft(vals, cur_len, shift);
    }
    const size_t final_len = cur_len >> RATE_BITS;
    wb_push(w, final_len);
    for (size_t i = 0; i < final_len; i++) {
        wb_push_ext(w, cf[i]);
        observe_ext(ch, cf[i]);
    }
    tick("fri commit phase");
    /* 7. proof of work: minimum witness */
    gl_t pow_witness = challenger_pow_grind(ch, POW_BITS);
    challenger_observe(ch, pow_witness);
    (void)challenger_get(ch);
    wb_push(w, pow_witness);
    tick("pow");
    /* 8. queries */
    for (int qi = 0; qi < NUM_QUERIES; qi++) {
        size_t x = (size_t)(challenger_get(ch) % m);
        for (size_t c = 0; c < C; c++) wb_push(w, lde[c * m + x]);
        merkle_open(w, &tree_t, x);
        for (int q = 0; q < N_QUOT; q++) wb_push(w, qlde[(size_t)q * m + x]);
        merkle_open(w, &tree_q, x);
        for (size_t l = 0; l < n_layers; l++) {
            size_t coset = x >> ARITY_BITS;
            wb_push_many(w, layer_leaves[l] + 32 * coset, 32);
            merkle_open(w, &layer_trees[l], coset);
            x = coset;
        }
    }
    for (size_t l = 0; l < n_layers; l++) {
        merkle_free(&layer_trees[l]);
        free(layer_leaves[l]);
    }
    free(layer_trees); free(layer_leaves); free(vals); free(final_poly);
    free(op_local); free(op_next); free(qlde); free(qcoef); free(qv); free(pertab);
    merkle_free(&tree_t); merkle_free(&tree_q);
    free(lde); free(coeffs);
}

static void transcript_init(challenger_t *ch, uint32_t kind, uint32_t n_max, const uint8_t *chain_id, size_t chain_id_len,
                            uint64_t skip_max, const uint8_t *input, size_t input_len, const uint8_t out32[32]) {
    gl_t dg[4], pub[128], ph[4];
    circuit_digest(kind, n_max, chain_id, chain_id_len, skip_max, dg);
    challenger_init(ch);
    challenger_observe_many(ch, dg, 4);
    size_t k = 0;
    for (size_t i = 0; i < input_len; i++) pub[k++] = input[i];
    for (size_t i = 0; i < 32; i++) pub[k++] = out32[i];
    poseidon_hash_no_pad(pub, k, ph);
    challenger_observe_many(ch, ph, 4);
}

/* debug / test hook: the NEXT tm_prove() on this thread adds one to cell (col, row) of `table` after witness generation,
 * i.e. it plays a prover that commits to an invalid witness; the verifiers must reject what it outputs. */
static _Thread_local struct { int active, table; size_t col, row; } g_corrupt;
void tm_debug_corrupt_next_proof(int table, size_t col, size_t row) {
    g_corrupt.active = 1;
    g_corrupt.table = table;
    g_corrupt.col = col;
    g_corrupt.row = row;
}

/* proof = header (8 u64), then the three table proofs.  Returns a TMX_CHECK id (0 = ok). */
int tm_prove(const uint8_t *input, size_t input_len, const uint8_t *blob, size_t blob_len, const uint8_t *chain_id,
             size_t chain_id_len, uint64_t skip_max, uint64_t **proof_out, size_t *proof_len, uint8_t out32[32]) {
    int rc = tm_verify_circuit(input, input_len, blob, blob_len, chain_id, chain_id_len, skip_max, out32);
    if (rc) return rc;
    const tmx_offchain_head *h = (const tmx_offchain_head *)blob;
    trace_t tr[3];
    rc = tm_build_traces(blob, blob_len, tr);
    if (rc) return rc;
    if (g_corrupt.active) {
        trace_t *t = &tr[g_corrupt.table];
        uint64_t *cell = &t->data[(g_corrupt.col % t->n_cols) * t->n_rows + g_corrupt.row % t->n_rows];
        *cell = gl_add(*cell, 1);
        g_corrupt.active = 0;
    }
    challenger_t ch;
    transcript_init(&ch, h->kind, h->n_max, chain_id, chain_id_len, skip_max, input, input_len, out32);
    g_kind = h->kind;
    g_n_max = h->n_max;
    wbuf_t w = {0};
    wb_push(&w, PROOF_MAGIC);
    wb_push(&w, h->kind);
    wb_push(&w, h->n_max);
    wb_push(&w, N_TABLES);
    for (int i = 0; i < 4; i++) {
        gl_t x = 0;
        for (int j = 0; j < 8; j++) x |= (gl_t)out32[8 * i + j] << (8 * j);
        wb_push(&w, x);
    }
    for (int t = 0; t < N_TABLES; t++) prove_table(t, &tr[t], &ch, &w);
    tm_free_traces(tr);
    *proof_out = w.v;
    *proof_len = w.n;
    return 0;
}

void tm_proof_free(uint64_t *p) { free(p); }

/* ------------------------------------------------------------------ verifier */
static int verify_table(int table, size_t n, rbuf_t *r, challenger_t *ch) {
    const size_t C = TABLE_COLS[table], m = n << RATE_BITS;
    const unsigned k = tmx_log2(n), km = k + RATE_BITS;
    const unsigned cap_h = km < CAP_HEIGHT ? km : CAP_HEIGHT;
    const size_t cap_n = (size_t)1 << cap_h;
    const gl_t *cap_t = rb_take(r, 4 * cap_n);
    if (r->err) return 1;
    observe_cap(ch, cap_t, cap_n);
    gl_t alpha[NUM_CHALLENGES];
    for (int i = 0; i < NUM_CHALLENGES; i++) alpha[i] = challenger_get(ch);
    const gl_t *cap_q = rb_take(r, 4 * cap_n);
    if (r->err) return 1;
    observe_cap(ch, cap_q, cap_n);
    const gl2_t zeta = challenger_get_ext(ch);
    const gl2_t zeta_next = gl2_scale(zeta, gl_root_of_unity(k));
    gl2_t *op_local = (gl2_t *)malloc(C * sizeof(gl2_t)), *op_next = (gl2_t *)malloc(C * sizeof(gl2_t));
    gl2_t op_quot[N_QUOT];
    for (size_t c = 0; c < C; c++) op_local[c] = rb_get_ext(r);
    for (size_t c = 0; c < C; c++) op_next[c] = rb_get_ext(r);
    for (int q = 0; q < N_QUOT; q++) op_quot[q] = rb_get_ext(r);
    int bad = r->err;
    for (size_t c = 0; c < C && !bad; c++) observe_ext(ch, op_local[c]);
    for (int q = 0; q < N_QUOT && !bad; q++) observe_ext(ch, op_quot[q]);
    for (size_t c = 0; c < C && !bad; c++) observe_ext(ch, op_next[c]);
    /* constraint identity at zeta */
    if (!bad) {
        const int nper = TABLE_NPER[table];
        const size_t P = table_period(table, n);
        gl2_t per[16];
        gl2_t y = gl2_pow(zeta, n / P);
        gl_t *pat = (gl_t *)malloc(P * sizeof(gl_t));
        for (int pc = 0; pc < nper; pc++) {
            for (size_t rr = 0; rr < P; rr++) pat[rr] = periodic_pattern(table, pc, rr);
            ntt_inverse(pat, P);
            per[pc] = base_poly_eval(pat, P, y);
        }
        free(pat);
        acc_e_t a;
        for (int i = 0; i < NUM_CHALLENGES; i++) {
            a.acc[i] = gl2_from(0);
            a.alpha[i] = gl2_from(alpha[i]);
        }
        air_eval_e(table, op_local, op_next, per, &a);
        gl2_t zh = gl2_sub(gl2_pow(zeta, n), gl2_from(1));
        gl2_t zn = gl2_pow(zeta, n);
        for (int i = 0; i < NUM_CHALLENGES; i++) {
            gl2_t q = gl2_add(op_quot[QDF * i], gl2_mul(zn, op_quot[QDF * i + 1]));
            if (!gl2_eq(gl2_mul(q, zh), a.acc[i])) bad = 2;
        }
    }
    /* FRI */
    const gl2_t fa = challenger_get_ext(ch);
    gl2_t red[2];
    if (!bad) {
        for (int batch = 0; batch < 2; batch++) {
            const size_t npoly = batch == 0 ? C + N_QUOT : C;
            gl2_t acc = gl2_from(0);
            for (size_t j = npoly; j-- > 0;) {
                gl2_t v = batch == 0 ? (j < C ? op_local[j] : op_quot[j - C]) : op_next[j];
                acc = gl2_add(gl2_mul(acc, fa), v);
            }
            red[batch] = acc;
        }
    }
    const size_t n_layers = fri_num_layers(k);
    const gl_t *layer_caps[16];
    gl2_t betas[16];
    size_t layer_rows = m;
    for (size_t l = 0; l < n_layers; l++) {
        layer_rows >>= ARITY_BITS;
        unsigned lg = tmx_log2(layer_rows);
        size_t lcap = (size_t)1 << (lg < CAP_HEIGHT ? lg : CAP_HEIGHT);
        layer_caps[l] = rb_take(r, 4 * lcap);
        if (r->err) { bad = 1; break; }
        observe_cap(ch, layer_caps[l], lcap);
        betas[l] = challenger_get_ext(ch);
    }
    size_t final_len = (size_t)rb_get(r);
    if (final_len != ((m >> (ARITY_BITS * n_layers)) >> RATE_BITS)) bad = bad ? bad : 3;
    gl2_t final_coeffs[64];
    for (size_t i = 0; i < final_len && i < 64 && !bad; i++) {
        final_coeffs[i] = rb_get_ext(r);
        observe_ext(ch, final_coeffs[i]);
    }
    gl_t pow_witness = rb_get(r);
    if (!bad) {
        challenger_observe(ch, pow_witness);
        gl_t resp = challenger_get(ch);
        if (POW_BITS && (resp >> (64 - POW_BITS)) != 0) bad = 4;
    }
    const gl2_t alpha_c = gl2_pow(fa, C);
    for (int qi = 0; qi < NUM_QUERIES && !bad; qi++) {
        size_t x = (size_t)(challenger_get(ch) % m);
        const gl_t *row_t = rb_take(r, C);
        const gl_t *path_t = rb_take(r, 4 * (km - cap_h));
        const gl_t *row_q = rb_take(r, N_QUOT);
        const gl_t *path_q = rb_take(r, 4 * (km - cap_h));
        if (r->err) { bad = 1; break; }
        if (!merkle_verify(row_t, C, x, path_t, km - cap_h, cap_t, cap_h)) { bad = 5; break; }
        if (!merkle_verify(row_q, N_QUOT, x, path_q, km - cap_h, cap_q, cap_h)) { bad = 5; break; }
        gl_t sx = gl_mul(GL_GENERATOR, gl_pow(gl_root_of_unity(km), tmx_bitrev(x, km)));
        /* fri_combine_initial */
        gl2_t sum = gl2_from(0);
        for (int batch = 0; batch < 2; batch++) {
            const size_t npoly = batch == 0 ? C + N_QUOT : C;
            gl2_t acc = gl2_from(0);
            for (size_t j = npoly; j-- > 0;) {
                gl_t v = j < C ? row_t[j] : row_q[j - C];
                acc = gl2_add(gl2_mul(acc, fa), gl2_from(v));
            }
            gl2_t num = gl2_sub(acc, red[batch]);
            gl2_t den = gl2_sub(gl2_from(sx), batch == 0 ? zeta : zeta_next);
            gl2_t sh = batch == 0 ? gl2_pow(fa, C + N_QUOT) : alpha_c;
            sum = gl2_add(gl2_mul(sum, sh), gl2_mul(num, gl2_inv(den)));
        }
        gl2_t old = sum;
        size_t rows = m;
        for (size_t l = 0; l < n_layers; l++) {
            rows >>= ARITY_BITS;
            unsigned lg = tmx_log2(rows);
            unsigned lcap_h = lg < CAP_HEIGHT ? lg : CAP_HEIGHT;
            const gl_t *leaf = rb_take(r, 32);
            const gl_t *path = rb_take(r, 4 * (lg - lcap_h));
            if (r->err) { bad = 1; break; }
            gl2_t ev[16];
            for (int i = 0; i < 16; i++) ev[i] = gl2_make(leaf[2 * i], leaf[2 * i + 1]);
            unsigned within = x & 15;
            size_t coset = x >> ARITY_BITS;
            if (!gl2_eq(ev[within], old)) { bad = 6; break; }
            old = fri_fold_coset(sx, within, ev, betas[l]);
            if (!merkle_verify(leaf, 32, coset, path, lg - lcap_h, layer_caps[l], lcap_h)) { bad = 5; break; }
            sx = gl_pow(sx, 16);
            x = coset;
        }
        if (bad) break;
        if (!gl2_eq(ext_poly_eval(final_coeffs, final_len, gl2_from(sx)), old)) bad = 7;
    }
    free(op_local);
    free(op_next);
    return bad;
}

/* Returns 0 if the proof verifies for (input, output) under the circuit (kind, n_max, chain id, skip_max). */
int tm_verify_proof(const uint64_t *proof, size_t proof_len, const uint8_t *input, size_t input_len, const uint8_t *chain_id,
                    size_t chain_id_len, uint64_t skip_max, uint32_t kind, uint32_t n_max, const uint8_t out32[32]) {
    rbuf_t r = {proof, proof_len, 0, 0};
    if (rb_get(&r) != PROOF_MAGIC || rb_get(&r) != kind || rb_get(&r) != n_max || rb_get(&r) != N_TABLES) return 100;
    for (int i = 0; i < 4; i++) {
        gl_t x = 0;
        for (int j = 0; j < 8; j++) x |= (gl_t)out32[8 * i + j] << (8 * j);
        if (rb_get(&r) != x) return 101;
    }
    if (input_len != (kind == TMX_KIND_SKIP ? 48u : 40u)) return 102;
    challenger_t ch;
    transcript_init(&ch, kind, n_max, chain_id, chain_id_len, skip_max, input, input_len, out32);
    size_t dims[6];
    tm_trace_dims(kind, n_max, dims);
    g_kind = kind;
    g_n_max = n_max;
    for (int t = 0; t < N_TABLES; t++) {
        int rc = verify_table(t, dims[2 * t], &r, &ch);
        if (rc) return 10 * (t + 1) + rc;
    }
    if (r.err || r.pos != r.n) return 103;
    return 0;
}

void tm_debug_set_shape(uint32_t kind, uint32_t n_max) {
    g_kind = kind;
    g_n_max = n_max;
}

/* debug / test hook (shape from tm_debug_set_shape): LDE of a trace and the quotient values on the LDE coset in natural order, [2][m] */
void tm_debug_quotient(int table, const uint64_t *trace, size_t n, size_t C, const uint64_t alpha[2], uint64_t *lde_out,
                       uint64_t *qv_out) {
    const size_t m = n << RATE_BITS;
    const unsigned km = tmx_log2(m);
    gl_t *coeffs = (gl_t *)malloc(C * n * sizeof(gl_t));
    ntt_lde_batch(trace, C, n, RATE_BITS, lde_out, coeffs);
    free(coeffs);
    const int nper = TABLE_NPER[table];
    const size_t P = table_period(table, n);
    gl_t *pertab = (gl_t *)calloc((size_t)(nper ? nper : 1) * 2 * P, sizeof(gl_t));
    gl_t shift = gl_pow(GL_GENERATOR, n / P);
    for (int pc = 0; pc < nper; pc++) {
        gl_t *t = pertab + (size_t)pc * 2 * P;
        for (size_t r = 0; r < P; r++) t[r] = periodic_pattern(table, pc, r);
        ntt_inverse(t, P);
        memset(t + P, 0, P * sizeof(gl_t));
        ntt_coset_forward(t, 2 * P, shift);
    }
    const gl_t gn = gl_pow(GL_GENERATOR, n);
    const gl_t zh_inv[2] = {gl_inv(gl_sub(gn, 1)), gl_inv(gl_sub(gl_neg(gn), 1))};
    gl_t *loc = (gl_t *)malloc(C * sizeof(gl_t)), *nxt = (gl_t *)malloc(C * sizeof(gl_t));
    for (size_t j = 0; j < m; j++) {
        size_t p = tmx_bitrev(j, km), p2 = tmx_bitrev((j + 2) & (m - 1), km);
        for (size_t c = 0; c < C; c++) {
            loc[c] = lde_out[c * m + p];
            nxt[c] = lde_out[c * m + p2];
        }
        gl_t per[16];
        for (int pc = 0; pc < nper; pc++) per[pc] = pertab[(size_t)pc * 2 * P + (j & (2 * P - 1))];
        acc_b_t a;
        for (int i = 0; i < 2; i++) { a.acc[i] = 0; a.alpha[i] = alpha[i]; }
        air_eval_b(table, loc, nxt, per, &a);
        for (int i = 0; i < 2; i++) qv_out[(size_t)i * m + j] = gl_mul(a.acc[i], zh_inv[j & 1]);
    }
    free(loc); free(nxt); free(pertab);
}

/* debug / test hook (shape from tm_debug_set_shape): the folded constraint values (two challenges) of the transitions
 * rows[i] -> rows[i] + 1 (cyclic) evaluated on the TRACE domain itself; all zero for a satisfying trace.  trace is
 * column-major [C][n]; out is [n_rows][2]. */
void tm_debug_constraints_at_rows(int table, const uint64_t *trace, size_t n, size_t C, const uint64_t alpha[2], const uint64_t *rows,
                                  size_t n_rows, uint64_t *out) {
    const int nper = TABLE_NPER[table];
    const size_t P = table_period(table, n);
    gl_t *loc = (gl_t *)malloc(C * sizeof(gl_t)), *nxt = (gl_t *)malloc(C * sizeof(gl_t));
    for (size_t i = 0; i < n_rows; i++) {
        const size_t r = rows[i] % n, r2 = (r + 1) % n;
        for (size_t c = 0; c < C; c++) {
            loc[c] = trace[c * n + r];
            nxt[c] = trace[c * n + r2];
        }
        gl_t per[16];
        for (int pc = 0; pc < nper; pc++) per[pc] = periodic_pattern(table, pc, r % P);
        acc_b_t a;
        for (int k = 0; k < 2; k++) { a.acc[k] = 0; a.alpha[k] = alpha[k]; }
        air_eval_b(table, loc, nxt, per, &a);
        out[2 * i] = a.acc[0];
        out[2 * i + 1] = a.acc[1];
    }
    free(loc); free(nxt);
}
