/*
 * oracle/stark.c -- deterministic CPU prover and verifier for the multi-table STARK with a shared bus that witnesses
 * one skip / step proof.  TEST INFRASTRUCTURE ONLY (see oracle/gl.h header); also the CPU baseline that bench.py times.
 *
 * Protocol (DESIGN.md "Proof format and protocol"): round 1 commits the witness tables (SHA-256, SHA-512, Ed25519,
 * logic, range); the bus challenges (beta, gamma) are drawn; round 2 commits, per table, the helper columns of its
 * bus interactions and a running sum; then per table: constraint challenges, quotient commitment, openings at zeta and
 * g*zeta of the constant / first-round / second-round columns, FRI (arity 16, final polynomial, 16-bit proof of work
 * with the MINIMUM witness, 84 queries).  The verifier also checks that the per-table bus totals and its own
 * public-input terms sum to zero.
 *
 * Restates the plonky2 0.2.0 / starky / Curta pipeline the reference reaches through `circuit.prove()` /
 * `circuit.verify()` [REF circuits/skip.rs:214,244,247; circuits/step.rs:196,223,226] -- fri/oracle.rs
 * (prove_openings), fri/prover.rs (fri_committed_trees, fri_proof_of_work, query rounds), fri/verifier.rs, starky
 * prover.rs (quotient polynomials, opening set), Curta's lookup / bus accumulators -- with Curta's STARK
 * configuration (rate_bits 1, cap height 4, 84 queries).  The dependency sources are absent and the reference pins no
 * proof bytes: PARITY UNPINNED at the proof level; the primitives underneath are
 * KAT-pinned.  The constraint systems are not typed in here: they are DATA, read from the build artefact
 * (oracle/circuit.c) and evaluated by this file's own interpreter, bus logic and quotient code.  The FRI commit phase
 * deliberately works in COEFFICIENT space like plonky2 (fold coefficients, re-evaluate with a coset FFT), while the
 * CUDA product folds evaluations; agreement of the two is part of the parity test.
 */
#include "oracle.h"
#include "oracle_w.h"
#include "circuit.h"
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
#define PROOF_MAGIC 0x32504D54ULL /* "TMP2" */

typedef struct {
    size_t n_rows, n_cols;
    uint64_t *data;
} trace_t;
int tm_build_traces(const uint8_t *blob, size_t blob_len, trace_t out[3]);
void tm_free_traces(trace_t out[3]);

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

/* ------------------------------------------------------------------ bus */
/* fingerprint gamma + tag + beta v_0 + beta^2 v_1 + ... of one bus item part at prog[p] = {tag, m, len, v..} */
static gl2_t bus_fingerprint_b(const uint64_t *part, const gl_t *v, gl2_t beta, gl2_t gamma) {
    gl2_t acc = gl2_from(0);
    const uint64_t len = part[2];
    for (uint64_t i = len; i-- > 0;) {
        acc.a0 = gl_add(acc.a0, v[part[3 + i]]);
        acc = gl2_mul(acc, beta);
    }
    acc.a0 = gl_add(acc.a0, v[part[0]]);
    return gl2_add(acc, gamma);
}

/* the same over openings at zeta: every component is an extension element, the algebra is F[X]/(X^2 - 7) on top */
typedef struct {
    gl2_t a0, a1;
} e2e_t;
static e2e_t e2e_add(e2e_t a, e2e_t b) { return (e2e_t){gl2_add(a.a0, b.a0), gl2_add(a.a1, b.a1)}; }
static e2e_t e2e_sub(e2e_t a, e2e_t b) { return (e2e_t){gl2_sub(a.a0, b.a0), gl2_sub(a.a1, b.a1)}; }
static e2e_t e2e_mul(e2e_t a, e2e_t b) {
    return (e2e_t){gl2_add(gl2_mul(a.a0, b.a0), gl2_scale(gl2_mul(a.a1, b.a1), 7)), gl2_add(gl2_mul(a.a0, b.a1), gl2_mul(a.a1, b.a0))};
}
static e2e_t e2e_scale(e2e_t a, gl2_t s) { return (e2e_t){gl2_mul(a.a0, s), gl2_mul(a.a1, s)}; }
static e2e_t bus_fingerprint_e(const uint64_t *part, const gl2_t *v, gl2_t beta, gl2_t gamma) {
    const e2e_t b = {gl2_from(beta.a0), gl2_from(beta.a1)};
    e2e_t acc = {gl2_from(0), gl2_from(0)};
    const uint64_t len = part[2];
    for (uint64_t i = len; i-- > 0;) {
        acc.a0 = gl2_add(acc.a0, v[part[3 + i]]);
        acc = e2e_mul(acc, b);
    }
    acc.a0 = gl2_add(acc.a0, v[part[0]]);
    return e2e_add(acc, (e2e_t){gl2_from(gamma.a0), gl2_from(gamma.a1)});
}

/* periodic values of the row (trace domain) */
static void periodic_at_row(const table_def_t *d, size_t row, gl_t *per) {
    for (uint32_t pc = 0; pc < d->n_per; pc++) per[pc] = d->periodic[(size_t)pc * d->period + row % d->period];
}

/* first-round trace of the range table: how often each value is looked up by the other tables */
#define HIST_SIZE ((1u << 16) + (1u << 11) + (1u << 8) + 2)
static int count_lookups(const circuit_def_t *c, trace_t tr[TMX_N_TABLES], uint64_t *hist) {
    int bad = 0;
    for (int ti = 0; ti < TMX_N_TABLES; ti++) {
        const table_def_t *d = &c->t[ti];
        if (!d->present || ti == TMX_T_RANGE) continue;
        const size_t n = tr[ti].n_rows;
#pragma omp parallel
        {
            gl_t *v = (gl_t *)malloc((d->n_nodes ? d->n_nodes : 1) * sizeof(gl_t));
            gl_t *loc = (gl_t *)malloc(d->n_main * sizeof(gl_t)), *nxt = (gl_t *)malloc(d->n_main * sizeof(gl_t));
            gl_t *k = (gl_t *)malloc((d->n_const ? d->n_const : 1) * sizeof(gl_t)), per[16];
            uint64_t *local_hist = (uint64_t *)calloc(HIST_SIZE, sizeof(uint64_t));
#pragma omp for schedule(static)
            for (size_t r = 0; r < n; r++) {
                for (size_t col = 0; col < d->n_main; col++) {
                    loc[col] = tr[ti].data[col * n + r];
                    nxt[col] = tr[ti].data[col * n + (r + 1) % n];
                }
                for (size_t col = 0; col < d->n_const; col++) k[col] = d->constants[col * n + r];
                periodic_at_row(d, r, per);
                circuit_eval_b(d, d->bus_mask, loc, nxt, k, per, v);
                for (size_t p = 0; p < d->prog_len;) {
                    const uint64_t kind = d->prog[p++];
                    if (kind == 0) { p++; continue; }
                    for (uint64_t part = 0; part < kind; part++) {
                        const uint64_t *it = d->prog + p;
                        p += 3 + it[2];
                        if (v[it[1]] != GL_P - 1) continue;
                        size_t base, lim;
                        const gl_t tag = v[it[0]];
                        if (tag == BUS_R16) { base = 0; lim = 1u << 16; }
                        else if (tag == BUS_R11) { base = 1u << 16; lim = 1u << 11; }
                        else if (tag == BUS_R8) { base = (1u << 16) + (1u << 11); lim = 1u << 8; }
                        else if (tag == BUS_R1) { base = (1u << 16) + (1u << 11) + (1u << 8); lim = 2; }
                        else continue;
                        const gl_t val = v[it[3]];
                        if (val >= lim) {
#pragma omp atomic write
                            bad = 1;
                        } else
                            local_hist[base + val]++;
                    }
                }
            }
#pragma omp critical
            for (size_t i = 0; i < HIST_SIZE; i++) hist[i] += local_hist[i];
            free(v); free(loc); free(nxt); free(k); free(local_hist);
        }
    }
    return bad;
}

/* second-round trace of one table: helper columns of its bus items and the running sum; returns the table's total */
static gl2_t build_aux(const table_def_t *d, const trace_t *tr, gl2_t beta, gl2_t gamma, gl_t *aux /* [2 (H + 1)][n] */) {
    const size_t n = tr->n_rows, H = d->n_helpers;
    gl2_t *rowsum = (gl2_t *)malloc(n * sizeof(gl2_t));
#pragma omp parallel
    {
        gl_t *v = (gl_t *)malloc((d->n_nodes ? d->n_nodes : 1) * sizeof(gl_t));
        gl_t *loc = (gl_t *)malloc(d->n_main * sizeof(gl_t)), *nxt = (gl_t *)malloc(d->n_main * sizeof(gl_t));
        gl_t *k = (gl_t *)malloc((d->n_const ? d->n_const : 1) * sizeof(gl_t)), per[16];
#pragma omp for schedule(static)
        for (size_t r = 0; r < n; r++) {
            for (size_t col = 0; col < d->n_main; col++) {
                loc[col] = tr->data[col * n + r];
                nxt[col] = tr->data[col * n + (r + 1) % n];
            }
            for (size_t col = 0; col < d->n_const; col++) k[col] = d->constants[col * n + r];
            periodic_at_row(d, r, per);
            circuit_eval_b(d, d->bus_mask, loc, nxt, k, per, v);
            gl2_t sum = gl2_from(0);
            size_t h = 0;
            for (size_t p = 0; p < d->prog_len;) {
                const uint64_t kind = d->prog[p++];
                if (kind == 0) { p++; continue; }
                gl2_t val;
                if (kind == 1) {
                    const uint64_t *a = d->prog + p;
                    p += 3 + a[2];
                    const gl_t m = v[a[1]];
                    val = m ? gl2_scale(gl2_inv(bus_fingerprint_b(a, v, beta, gamma)), m) : gl2_from(0);
                } else {
                    const uint64_t *a = d->prog + p;
                    p += 3 + a[2];
                    const uint64_t *b = d->prog + p;
                    p += 3 + b[2];
                    const gl_t ma = v[a[1]], mb = v[b[1]];
                    if (!ma && !mb)
                        val = gl2_from(0);
                    else {
                        const gl2_t fa = bus_fingerprint_b(a, v, beta, gamma), fb = bus_fingerprint_b(b, v, beta, gamma);
                        val = gl2_mul(gl2_add(gl2_scale(fb, ma), gl2_scale(fa, mb)), gl2_inv(gl2_mul(fa, fb)));
                    }
                }
                aux[(2 * h) * n + r] = val.a0;
                aux[(2 * h + 1) * n + r] = val.a1;
                sum = gl2_add(sum, val);
                h++;
            }
            rowsum[r] = sum;
        }
        free(v); free(loc); free(nxt); free(k);
    }
    gl2_t total = gl2_from(0);
    for (size_t r = 0; r < n; r++) total = gl2_add(total, rowsum[r]);
    /* Z(0) = 0, Z(r + 1) = Z(r) + rowsum(r) - total / n: closes up cyclically */
    const gl2_t step = gl2_scale(total, gl_inv((gl_t)n));
    gl2_t z = gl2_from(0);
    for (size_t r = 0; r < n; r++) {
        aux[(2 * H) * n + r] = z.a0;
        aux[(2 * H + 1) * n + r] = z.a1;
        z = gl2_sub(gl2_add(z, rowsum[r]), step);
    }
    free(rowsum);
    return total;
}

/* folds the constraints of one evaluation point (base field): table constraints, helper constraints, running sum */
static void fold_constraints_b(const table_def_t *d, const gl_t *v, const gl_t *aux_l, const gl_t *aux_n, gl2_t beta, gl2_t gamma,
                               gl2_t s_over_n, const gl_t alpha[NUM_CHALLENGES], gl_t acc[NUM_CHALLENGES]) {
#define EMIT(x)                                                                              \
    do {                                                                                     \
        const gl_t _c = (x);                                                                 \
        for (int _i = 0; _i < NUM_CHALLENGES; _i++) acc[_i] = gl_add(gl_mul(acc[_i], alpha[_i]), _c); \
    } while (0)
    for (int i = 0; i < NUM_CHALLENGES; i++) acc[i] = 0;
    gl2_t sum = gl2_from(0);
    size_t h = 0;
    for (size_t p = 0; p < d->prog_len;) {
        const uint64_t kind = d->prog[p++];
        if (kind == 0) {
            EMIT(v[d->prog[p++]]);
            continue;
        }
        const gl2_t H = gl2_make(aux_l[2 * h], aux_l[2 * h + 1]);
        h++;
        sum = gl2_add(sum, H);
        gl2_t c;
        if (kind == 1) {
            const uint64_t *a = d->prog + p;
            p += 3 + a[2];
            c = gl2_mul(H, bus_fingerprint_b(a, v, beta, gamma));
            c.a0 = gl_sub(c.a0, v[a[1]]);
        } else {
            const uint64_t *a = d->prog + p;
            p += 3 + a[2];
            const uint64_t *b = d->prog + p;
            p += 3 + b[2];
            const gl2_t fa = bus_fingerprint_b(a, v, beta, gamma), fb = bus_fingerprint_b(b, v, beta, gamma);
            c = gl2_sub(gl2_mul(H, gl2_mul(fa, fb)), gl2_add(gl2_scale(fb, v[a[1]]), gl2_scale(fa, v[b[1]])));
        }
        EMIT(c.a0);
        EMIT(c.a1);
    }
    const gl2_t z = gl2_make(aux_l[2 * h], aux_l[2 * h + 1]), zn = gl2_make(aux_n[2 * h], aux_n[2 * h + 1]);
    const gl2_t c = gl2_add(gl2_sub(gl2_sub(zn, z), sum), s_over_n);
    EMIT(c.a0);
    EMIT(c.a1);
#undef EMIT
}

static void fold_constraints_e(const table_def_t *d, const gl2_t *v, const gl2_t *aux_l, const gl2_t *aux_n, gl2_t beta, gl2_t gamma,
                               gl2_t s_over_n, const gl_t alpha[NUM_CHALLENGES], gl2_t acc[NUM_CHALLENGES]) {
#define EMIT(x)                                                                                              \
    do {                                                                                                     \
        const gl2_t _c = (x);                                                                                \
        for (int _i = 0; _i < NUM_CHALLENGES; _i++) acc[_i] = gl2_add(gl2_scale(acc[_i], alpha[_i]), _c);    \
    } while (0)
    for (int i = 0; i < NUM_CHALLENGES; i++) acc[i] = gl2_from(0);
    e2e_t sum = {gl2_from(0), gl2_from(0)};
    size_t h = 0;
    for (size_t p = 0; p < d->prog_len;) {
        const uint64_t kind = d->prog[p++];
        if (kind == 0) {
            EMIT(v[d->prog[p++]]);
            continue;
        }
        const e2e_t H = {aux_l[2 * h], aux_l[2 * h + 1]};
        h++;
        sum = e2e_add(sum, H);
        e2e_t c;
        if (kind == 1) {
            const uint64_t *a = d->prog + p;
            p += 3 + a[2];
            c = e2e_mul(H, bus_fingerprint_e(a, v, beta, gamma));
            c.a0 = gl2_sub(c.a0, v[a[1]]);
        } else {
            const uint64_t *a = d->prog + p;
            p += 3 + a[2];
            const uint64_t *b = d->prog + p;
            p += 3 + b[2];
            const e2e_t fa = bus_fingerprint_e(a, v, beta, gamma), fb = bus_fingerprint_e(b, v, beta, gamma);
            c = e2e_sub(e2e_mul(H, e2e_mul(fa, fb)), e2e_add(e2e_scale(fb, v[a[1]]), e2e_scale(fa, v[b[1]])));
        }
        EMIT(c.a0);
        EMIT(c.a1);
    }
    const e2e_t z = {aux_l[2 * h], aux_l[2 * h + 1]}, zn = {aux_n[2 * h], aux_n[2 * h + 1]};
    e2e_t c = e2e_sub(e2e_sub(zn, z), sum);
    c.a0 = gl2_add(c.a0, gl2_from(s_over_n.a0));
    c.a1 = gl2_add(c.a1, gl2_from(s_over_n.a1));
    EMIT(c.a0);
    EMIT(c.a1);
#undef EMIT
}

/* ------------------------------------------------------------------ prover */
typedef struct {
    const table_def_t *d;
    size_t n, m;
    unsigned k, km;
    size_t Kc, C, A;
    gl_t *lde_k, *coef_k, *lde_m, *coef_m, *lde_a, *coef_a;
    merkle_tree_t tree_k, tree_m, tree_a;
    gl2_t total;
} ptable_t;

static void commit_batch(const gl_t *values, size_t cols, size_t n, gl_t **lde, gl_t **coef, merkle_tree_t *tree) {
    const size_t m = n << RATE_BITS;
    *lde = (gl_t *)malloc((cols ? cols : 1) * m * sizeof(gl_t));
    *coef = (gl_t *)malloc((cols ? cols : 1) * n * sizeof(gl_t));
    memset(tree, 0, sizeof *tree);
    if (!cols) return;
    ntt_lde_batch(values, cols, n, RATE_BITS, *lde, *coef);
    commit_columns(tree, *lde, cols, m, CAP_HEIGHT);
}

/* values on the LDE coset (natural index mod 2P) of the periodic columns, [n_per][2P] */
static gl_t *periodic_lde(const table_def_t *d, size_t n) {
    const size_t P = d->period;
    gl_t *pertab = (gl_t *)calloc((size_t)(d->n_per ? d->n_per : 1) * 2 * P, sizeof(gl_t));
    const gl_t shift = gl_pow(GL_GENERATOR, n / P);
    for (uint32_t pc = 0; pc < d->n_per; pc++) {
        gl_t *t = pertab + (size_t)pc * 2 * P;
        memcpy(t, d->periodic + (size_t)pc * P, P * sizeof(gl_t));
        ntt_inverse(t, P);
        memset(t + P, 0, P * sizeof(gl_t));
        ntt_coset_forward(t, 2 * P, shift);
    }
    return pertab;
}

/* quotient values on the LDE coset, natural order, [NUM_CHALLENGES][m] */
static void quotient_values(const ptable_t *pt, gl2_t beta, gl2_t gamma, const gl_t alpha[NUM_CHALLENGES], gl_t *qv) {
    const table_def_t *d = pt->d;
    const size_t n = pt->n, m = pt->m, P = d->period;
    gl_t *pertab = periodic_lde(d, n);
    const gl_t gn = gl_pow(GL_GENERATOR, n);
    const gl_t zh_inv[2] = {gl_inv(gl_sub(gn, 1)), gl_inv(gl_sub(gl_neg(gn), 1))};
    const gl2_t s_over_n = gl2_scale(pt->total, gl_inv((gl_t)n));
#pragma omp parallel
    {
        gl_t *v = (gl_t *)malloc((d->n_nodes ? d->n_nodes : 1) * sizeof(gl_t));
        gl_t *loc = (gl_t *)malloc(pt->C * sizeof(gl_t)), *nxt = (gl_t *)malloc(pt->C * sizeof(gl_t));
        gl_t *k = (gl_t *)malloc((pt->Kc ? pt->Kc : 1) * sizeof(gl_t)), per[16];
        gl_t *al = (gl_t *)malloc(pt->A * sizeof(gl_t)), *an = (gl_t *)malloc(pt->A * sizeof(gl_t));
#pragma omp for schedule(static)
        for (size_t j = 0; j < m; j++) {
            const size_t p = tmx_bitrev(j, pt->km), p2 = tmx_bitrev((j + (1u << RATE_BITS)) & (m - 1), pt->km);
            for (size_t c = 0; c < pt->C; c++) {
                loc[c] = pt->lde_m[c * m + p];
                nxt[c] = pt->lde_m[c * m + p2];
            }
            for (size_t c = 0; c < pt->Kc; c++) k[c] = pt->lde_k[c * m + p];
            for (size_t c = 0; c < pt->A; c++) {
                al[c] = pt->lde_a[c * m + p];
                an[c] = pt->lde_a[c * m + p2];
            }
            for (uint32_t pc = 0; pc < d->n_per; pc++) per[pc] = pertab[(size_t)pc * 2 * P + (j & (2 * P - 1))];
            circuit_eval_b(d, NULL, loc, nxt, k, per, v);
            gl_t acc[NUM_CHALLENGES];
            fold_constraints_b(d, v, al, an, beta, gamma, s_over_n, alpha, acc);
            for (int i = 0; i < NUM_CHALLENGES; i++) qv[(size_t)i * m + j] = gl_mul(acc[i], zh_inv[j & 1]);
        }
        free(v); free(loc); free(nxt); free(k); free(al); free(an);
    }
    free(pertab);
}

/* everything of one table after both commitment rounds */
static void prove_table_tail(ptable_t *pt, gl2_t beta, gl2_t gamma, challenger_t *ch, wbuf_t *w) {
    t_last = omp_get_wtime();
    const size_t n = pt->n, m = pt->m, CT = pt->Kc + pt->C + pt->A;
    const unsigned k = pt->k, km = pt->km;
    gl_t alpha[NUM_CHALLENGES];
    for (int i = 0; i < NUM_CHALLENGES; i++) alpha[i] = challenger_get(ch);
    gl_t *qv = (gl_t *)malloc((size_t)NUM_CHALLENGES * m * sizeof(gl_t)); /* natural order */
    quotient_values(pt, beta, gamma, alpha, qv);
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
    const size_t cap_n = (size_t)1 << tree_q.cap_height;
    wb_push_many(w, tree_q.cap, 4 * cap_n);
    observe_cap(ch, tree_q.cap, cap_n);
    tick("quotient commit");
    /* polynomials in opening order: constant, first-round, second-round columns, then the quotient chunks */
    const gl_t **poly = (const gl_t **)malloc((CT + N_QUOT) * sizeof(gl_t *));
    {
        size_t o = 0;
        for (size_t c = 0; c < pt->Kc; c++) poly[o++] = pt->coef_k + c * n;
        for (size_t c = 0; c < pt->C; c++) poly[o++] = pt->coef_m + c * n;
        for (size_t c = 0; c < pt->A; c++) poly[o++] = pt->coef_a + c * n;
        for (int q = 0; q < N_QUOT; q++) poly[o++] = qcoef + (size_t)q * n;
    }
    /* openings */
    const gl2_t zeta = challenger_get_ext(ch);
    const gl2_t zeta_next = gl2_scale(zeta, gl_root_of_unity(k));
    gl2_t *op_local = (gl2_t *)malloc((CT + N_QUOT) * sizeof(gl2_t)), *op_next = (gl2_t *)malloc(CT * sizeof(gl2_t));
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t c = 0; c < CT + N_QUOT; c++) {
        op_local[c] = base_poly_eval(poly[c], n, zeta);
        if (c < CT) op_next[c] = base_poly_eval(poly[c], n, zeta_next);
    }
    for (size_t c = 0; c < CT; c++) wb_push_ext(w, op_local[c]);
    for (size_t c = 0; c < CT; c++) wb_push_ext(w, op_next[c]);
    for (int q = 0; q < N_QUOT; q++) wb_push_ext(w, op_local[CT + q]);
    for (size_t c = 0; c < CT; c++) observe_ext(ch, op_local[c]);
    for (int q = 0; q < N_QUOT; q++) observe_ext(ch, op_local[CT + q]);
    for (size_t c = 0; c < CT; c++) observe_ext(ch, op_next[c]);
    tick("openings");
    /* FRI batch polynomial, coefficient space (plonky2 fri/oracle.rs prove_openings) */
    const gl2_t fa = challenger_get_ext(ch);
    gl2_t *final_poly = (gl2_t *)calloc(m, sizeof(gl2_t));
    {
        gl2_t *comp = (gl2_t *)malloc(n * sizeof(gl2_t));
        for (int batch = 0; batch < 2; batch++) {
            const size_t npoly = batch == 0 ? CT + N_QUOT : CT;
            const gl2_t z = batch == 0 ? zeta : zeta_next;
            /* composition = sum_j alpha^j f_j (Horner from the last polynomial) */
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; i++) {
                gl2_t acc = gl2_from(0);
                for (size_t j = npoly; j-- > 0;) acc = gl2_add(gl2_mul(acc, fa), gl2_from(poly[j][i]));
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
    /* FRI commit phase (fri_committed_trees) */
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
        const gl2_t fbeta = challenger_get_ext(ch);
        size_t new_len = cur_len >> ARITY_BITS;
        for (size_t i = 0; i < new_len; i++) {
            gl2_t acc = gl2_from(0);
            for (int j = 15; j >= 0; j--) acc = gl2_add(gl2_mul(acc, fbeta), cf[16 * i + j]);
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
    /* proof of work: minimum witness */
    gl_t pow_witness = challenger_pow_grind(ch, POW_BITS);
    challenger_observe(ch, pow_witness);
    (void)challenger_get(ch);
    wb_push(w, pow_witness);
    tick("pow");
    /* queries */
    for (int qi = 0; qi < NUM_QUERIES; qi++) {
        size_t x = (size_t)(challenger_get(ch) % m);
        if (pt->Kc) {
            for (size_t c = 0; c < pt->Kc; c++) wb_push(w, pt->lde_k[c * m + x]);
            merkle_open(w, &pt->tree_k, x);
        }
        for (size_t c = 0; c < pt->C; c++) wb_push(w, pt->lde_m[c * m + x]);
        merkle_open(w, &pt->tree_m, x);
        for (size_t c = 0; c < pt->A; c++) wb_push(w, pt->lde_a[c * m + x]);
        merkle_open(w, &pt->tree_a, x);
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
    free(op_local); free(op_next); free(qlde); free(qcoef); free(qv); free(poly);
    merkle_free(&tree_q);
}

static void transcript_init(challenger_t *ch, const circuit_def_t *c, const uint8_t *input, size_t input_len, const uint8_t out32[32]) {
    gl_t pub[128], ph[4];
    challenger_init(ch);
    challenger_observe_many(ch, c->digest, 4);
    size_t k = 0;
    for (size_t i = 0; i < input_len && k < 96; i++) pub[k++] = input[i];
    for (size_t i = 0; i < 32; i++) pub[k++] = out32[i];
    poseidon_hash_no_pad(pub, k, ph);
    challenger_observe_many(ch, ph, 4);
}

/* The verifier's own bus terms: what ties the tables to the public input and output [REF circuits/skip.rs:119-133,
 * circuits/step.rs:106-117].  The verifier CONSUMES the root of every header proof (NODE message: id, 8 big-endian words, 1) --
 * the trusted header for the trusted validators-hash proof of a skip, the previous header for the next-validators-hash proof of
 * a step, the proven header (out32) for all others -- and PROVIDES the height as nine varint septets (height leaf), the
 * previous header (last-block-id leaf of a step) and (height, proven header) for the sign-bytes checks. */
static gl2_t fingerprint_words(gl2_t beta, gl2_t gamma, uint64_t tag, const uint64_t *v, size_t n) {
    gl2_t acc = gl2_from(0);
    for (size_t i = n; i-- > 0;) {
        acc.a0 = gl_add(acc.a0, v[i] % GL_P);
        acc = gl2_mul(acc, beta);
    }
    acc.a0 = gl_add(acc.a0, tag);
    return gl2_add(acc, gamma);
}
static uint64_t load_be64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}
static void header_words(const uint8_t *h, uint64_t *w) {
    for (int i = 0; i < 8; i++) w[i] = ((uint64_t)h[4 * i] << 24) | ((uint64_t)h[4 * i + 1] << 16) | ((uint64_t)h[4 * i + 2] << 8) | h[4 * i + 3];
}
static gl2_t public_terms(const circuit_def_t *c, const uint8_t *input, const uint8_t out32[32], gl2_t beta, gl2_t gamma) {
    gl2_t sum = gl2_from(0);
    if (!c->t[TMX_T_LOGIC].present) return sum;
    const int skip = c->kind == TMX_KIND_SKIP;
    const uint64_t height = skip ? load_be64(input + 40) : load_be64(input) + 1;
    const uint8_t *other = input + 8;
    const uint32_t n_proofs = skip ? 4 : 5;
    uint64_t v[16];
    for (uint32_t kp = 0; kp < n_proofs; kp++) {
        const uint8_t *hdr = (skip && kp == 0) || (!skip && kp == 4) ? other : out32;
        v[0] = ((uint64_t)0x7F << 24) | ((uint64_t)kp << 8) | 4;
        header_words(hdr, v + 1);
        v[9] = 1;
        sum = gl2_sub(sum, gl2_inv(fingerprint_words(beta, gamma, BUS_NODE, v, 10)));
    }
    if (!skip) {
        v[0] = PUB_PREV;
        header_words(other, v + 1);
        sum = gl2_add(sum, gl2_inv(fingerprint_words(beta, gamma, BUS_PUB, v, 9)));
    }
    v[0] = PUB_HEIGHT;
    for (int k = 0; k < 9; k++) v[1 + k] = (height >> (7 * k)) & 0x7F;
    sum = gl2_add(sum, gl2_inv(fingerprint_words(beta, gamma, BUS_PUB, v, 10)));
    v[0] = PUB_GLOB;
    v[1] = height & 0xFFFFFFFFULL;
    v[2] = height >> 32;
    header_words(out32, v + 3);
    sum = gl2_add(sum, gl2_inv(fingerprint_words(beta, gamma, BUS_PUB, v, 11)));
    return sum;
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
/* debug / test hook: the NEXT tm_prove() on this thread adds deltas[i] to cell (cols[i], rows[i]) of tables[i] after ALL witness
 * generation (range-table multiplicities included), up to 16 cells: a prover that commits to a chosen invalid witness */
static _Thread_local struct { int n, table[16]; size_t col[16], row[16]; uint64_t delta[16]; } g_patch;
void tm_debug_patch_next_proof(int n, const int *tables, const size_t *cols, const size_t *rows, const uint64_t *deltas) {
    g_patch.n = n > 16 ? 16 : n;
    for (int i = 0; i < g_patch.n; i++) {
        g_patch.table[i] = tables[i];
        g_patch.col[i] = cols[i];
        g_patch.row[i] = rows[i];
        g_patch.delta[i] = deltas[i] % GL_P;
    }
}
/* debug / test hook: the NEXT tm_prove() on this thread skips the statement pre-check (verify_skip / verify_step on the
 * inputs), i.e. it plays a prover that tries to prove a false statement with otherwise honest tables */
static _Thread_local int g_skip_precheck;
void tm_debug_skip_precheck_next_proof(void) { g_skip_precheck = 1; }

void *tm_circuit_load(const uint64_t *words, size_t n_words, int *rc) {
    circuit_def_t *c = (circuit_def_t *)malloc(sizeof *c);
    *rc = circuit_parse(words, n_words, c);
    if (*rc) {
        free(c);
        return NULL;
    }
    return c;
}
void tm_circuit_unload(void *c) {
    if (!c) return;
    circuit_free((circuit_def_t *)c);
    free(c);
}
void tm_circuit_digest(const void *c, uint64_t out[4]) { memcpy(out, ((const circuit_def_t *)c)->digest, 32); }

/* all first-round traces of one proof: the three witness tables, the logic table (given by the caller) and the range
 * table (counted here).  Returns a TMX_CHECK id. */
static int build_all_traces(const circuit_def_t *c, const uint8_t *blob, size_t blob_len, const uint64_t *logic_trace,
                            trace_t tr[TMX_N_TABLES]) {
    memset(tr, 0, TMX_N_TABLES * sizeof(trace_t));
    int rc = tm_build_traces(blob, blob_len, tr);
    if (rc) return rc;
    for (int t = 0; t < 3; t++)
        if (!c->t[t].present || tr[t].n_rows != ((size_t)1 << c->t[t].log_n) || tr[t].n_cols != c->t[t].n_main) return TMX_CHECK_INPUT;
    if (c->t[TMX_T_LOGIC].present) {
        const table_def_t *d = &c->t[TMX_T_LOGIC];
        if (!logic_trace) return TMX_CHECK_INPUT;
        tr[TMX_T_LOGIC].n_rows = (size_t)1 << d->log_n;
        tr[TMX_T_LOGIC].n_cols = d->n_main;
        const size_t cells = tr[TMX_T_LOGIC].n_rows * d->n_main;
        tr[TMX_T_LOGIC].data = (uint64_t *)malloc(cells * sizeof(uint64_t));
        memcpy(tr[TMX_T_LOGIC].data, logic_trace, cells * sizeof(uint64_t));
    }
    return 0;
}
static int build_range_trace(const circuit_def_t *c, trace_t tr[TMX_N_TABLES]) {
    const table_def_t *d = &c->t[TMX_T_RANGE];
    if (!d->present) return 0;
    const size_t n = (size_t)1 << d->log_n;
    uint64_t *hist = (uint64_t *)calloc(HIST_SIZE, sizeof(uint64_t));
    const int bad = count_lookups(c, tr, hist);
    tr[TMX_T_RANGE].n_rows = n;
    tr[TMX_T_RANGE].n_cols = RG_COLS;
    tr[TMX_T_RANGE].data = (uint64_t *)calloc(n * RG_COLS, sizeof(uint64_t));
    for (size_t i = 0; i < (1u << 16) && i < n; i++) tr[TMX_T_RANGE].data[RG_M16 * n + i] = hist[i];
    for (size_t i = 0; i < (1u << 11); i++) tr[TMX_T_RANGE].data[RG_M11 * n + i] = hist[(1u << 16) + i];
    for (size_t i = 0; i < (1u << 8); i++) tr[TMX_T_RANGE].data[RG_M8 * n + i] = hist[(1u << 16) + (1u << 11) + i];
    for (size_t i = 0; i < 2; i++) tr[TMX_T_RANGE].data[RG_M1 * n + i] = hist[(1u << 16) + (1u << 11) + (1u << 8) + i];
    free(hist);
    return bad;
}
static void free_all_traces(trace_t tr[TMX_N_TABLES]) {
    for (int t = 0; t < TMX_N_TABLES; t++) {
        free(tr[t].data);
        tr[t].data = NULL;
    }
}

/* proof = header (8 u64), round-1 caps, round-2 caps and totals, then the per-table tails.  Returns a TMX_CHECK id (0 = ok). */
int tm_prove(const void *circuit, const uint8_t *input, size_t input_len, const uint8_t *blob, size_t blob_len,
             const uint64_t *logic_trace, uint64_t **proof_out, size_t *proof_len, uint8_t out32[32]) {
    const circuit_def_t *c = (const circuit_def_t *)circuit;
    const tmx_offchain_head *h = (const tmx_offchain_head *)blob;
    if (blob_len < sizeof *h || h->kind != c->kind || h->n_max != c->n_max) return TMX_CHECK_INPUT;
    int rc = tm_verify_circuit(input, input_len, blob, blob_len, c->chain_id, c->chain_len, c->skip_max, out32);
    if (g_skip_precheck) {
        g_skip_precheck = 0;
        if (rc) memcpy(out32, h->header, 32);
        rc = 0;
    }
    if (rc) return rc;
    trace_t tr[TMX_N_TABLES];
    rc = build_all_traces(c, blob, blob_len, logic_trace, tr);
    if (rc) {
        free_all_traces(tr);
        return rc;
    }
    /* a value outside its range table cannot be proved: the honest prover stops, a cheating one would go on with wrong
     * multiplicities and be rejected by the bus balance (tm_debug hooks exercise that path) */
    const int range_bad = build_range_trace(c, tr);
    (void)range_bad;
    if (g_corrupt.active) {
        trace_t *t = &tr[g_corrupt.table];
        if (t->data) {
            uint64_t *cell = &t->data[(g_corrupt.col % t->n_cols) * t->n_rows + g_corrupt.row % t->n_rows];
            *cell = gl_add(*cell, 1);
        }
        g_corrupt.active = 0;
    }
    for (int i = 0; i < g_patch.n; i++) {
        trace_t *t = &tr[g_patch.table[i]];
        if (t->data) {
            uint64_t *cell = &t->data[(g_patch.col[i] % t->n_cols) * t->n_rows + g_patch.row[i] % t->n_rows];
            *cell = gl_add(*cell, g_patch.delta[i]);
        }
    }
    g_patch.n = 0;
    challenger_t ch;
    transcript_init(&ch, c, input, input_len, out32);
    wbuf_t w = {0};
    wb_push(&w, PROOF_MAGIC);
    wb_push(&w, h->kind);
    wb_push(&w, h->n_max);
    wb_push(&w, TMX_N_TABLES);
    for (int i = 0; i < 4; i++) {
        gl_t x = 0;
        for (int j = 0; j < 8; j++) x |= (gl_t)out32[8 * i + j] << (8 * j);
        wb_push(&w, x);
    }
    ptable_t pt[TMX_N_TABLES];
    memset(pt, 0, sizeof pt);
    t_last = omp_get_wtime();
    /* round 1 */
    for (int t = 0; t < TMX_N_TABLES; t++) {
        const table_def_t *d = &c->t[t];
        if (!d->present) continue;
        ptable_t *p = &pt[t];
        p->d = d;
        p->n = (size_t)1 << d->log_n;
        p->m = p->n << RATE_BITS;
        p->k = d->log_n;
        p->km = p->k + RATE_BITS;
        p->Kc = d->n_const; p->C = d->n_main; p->A = 2 * ((size_t)d->n_helpers + 1);
        commit_batch(d->constants, p->Kc, p->n, &p->lde_k, &p->coef_k, &p->tree_k);
        commit_batch(tr[t].data, p->C, p->n, &p->lde_m, &p->coef_m, &p->tree_m);
        const size_t cap_n = (size_t)1 << p->tree_m.cap_height;
        wb_push_many(&w, p->tree_m.cap, 4 * cap_n);
        observe_cap(&ch, p->tree_m.cap, cap_n);
    }
    tick("round 1 commitments");
    const gl2_t beta = challenger_get_ext(&ch), gamma = challenger_get_ext(&ch);
    /* round 2 */
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!c->t[t].present) continue;
        ptable_t *p = &pt[t];
        gl_t *aux = (gl_t *)malloc(p->A * p->n * sizeof(gl_t));
        p->total = build_aux(p->d, &tr[t], beta, gamma, aux);
        commit_batch(aux, p->A, p->n, &p->lde_a, &p->coef_a, &p->tree_a);
        free(aux);
        const size_t cap_n = (size_t)1 << p->tree_a.cap_height;
        wb_push_many(&w, p->tree_a.cap, 4 * cap_n);
        wb_push_ext(&w, p->total);
        observe_cap(&ch, p->tree_a.cap, cap_n);
        observe_ext(&ch, p->total);
    }
    tick("round 2 commitments");
    free_all_traces(tr);
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!c->t[t].present) continue;
        ptable_t *p = &pt[t];
        challenger_t fork = ch; /* each table continues on a fork of the transcript: common state + table index */
        challenger_observe(&fork, (gl_t)t);
        prove_table_tail(p, beta, gamma, &fork, &w);
        free(p->lde_k); free(p->coef_k); free(p->lde_m); free(p->coef_m); free(p->lde_a); free(p->coef_a);
        merkle_free(&p->tree_k); merkle_free(&p->tree_m); merkle_free(&p->tree_a);
    }
    *proof_out = w.v;
    *proof_len = w.n;
    return 0;
}

void tm_proof_free(uint64_t *p) { free(p); }

/* ------------------------------------------------------------------ verifier */
static int verify_table_tail(const circuit_def_t *c, int table, const gl_t *cap_m, const gl_t *cap_a, gl2_t beta, gl2_t gamma,
                             gl2_t total, rbuf_t *r, challenger_t *ch) {
    const table_def_t *d = &c->t[table];
    const size_t n = (size_t)1 << d->log_n, m = n << RATE_BITS;
    const size_t Kc = d->n_const, C = d->n_main, A = 2 * ((size_t)d->n_helpers + 1), CT = Kc + C + A;
    const unsigned k = d->log_n, km = k + RATE_BITS;
    const unsigned cap_h = km < CAP_HEIGHT ? km : CAP_HEIGHT;
    const size_t cap_n = (size_t)1 << cap_h;
    gl_t alpha[NUM_CHALLENGES];
    for (int i = 0; i < NUM_CHALLENGES; i++) alpha[i] = challenger_get(ch);
    const gl_t *cap_q = rb_take(r, 4 * cap_n);
    if (r->err) return 1;
    observe_cap(ch, cap_q, cap_n);
    const gl2_t zeta = challenger_get_ext(ch);
    const gl2_t zeta_next = gl2_scale(zeta, gl_root_of_unity(k));
    gl2_t *op_local = (gl2_t *)malloc((CT + N_QUOT) * sizeof(gl2_t)), *op_next = (gl2_t *)malloc(CT * sizeof(gl2_t));
    for (size_t i = 0; i < CT; i++) op_local[i] = rb_get_ext(r);
    for (size_t i = 0; i < CT; i++) op_next[i] = rb_get_ext(r);
    for (int q = 0; q < N_QUOT; q++) op_local[CT + q] = rb_get_ext(r);
    int bad = r->err;
    for (size_t i = 0; i < CT && !bad; i++) observe_ext(ch, op_local[i]);
    for (int q = 0; q < N_QUOT && !bad; q++) observe_ext(ch, op_local[CT + q]);
    for (size_t i = 0; i < CT && !bad; i++) observe_ext(ch, op_next[i]);
    /* constraint identity at zeta */
    if (!bad) {
        const size_t P = d->period;
        gl2_t per[16];
        const gl2_t y = gl2_pow(zeta, n / P);
        gl_t *pat = (gl_t *)malloc(P * sizeof(gl_t));
        for (uint32_t pc = 0; pc < d->n_per; pc++) {
            memcpy(pat, d->periodic + (size_t)pc * P, P * sizeof(gl_t));
            ntt_inverse(pat, P);
            per[pc] = base_poly_eval(pat, P, y);
        }
        free(pat);
        gl2_t *v = (gl2_t *)malloc((d->n_nodes ? d->n_nodes : 1) * sizeof(gl2_t));
        circuit_eval_e(d, op_local + Kc, op_next + Kc, op_local, per, v);
        gl2_t acc[NUM_CHALLENGES];
        fold_constraints_e(d, v, op_local + Kc + C, op_next + Kc + C, beta, gamma, gl2_scale(total, gl_inv((gl_t)n)), alpha, acc);
        free(v);
        const gl2_t zn = gl2_pow(zeta, n), zh = gl2_sub(zn, gl2_from(1));
        for (int i = 0; i < NUM_CHALLENGES; i++) {
            gl2_t q = gl2_add(op_local[CT + QDF * i], gl2_mul(zn, op_local[CT + QDF * i + 1]));
            if (!gl2_eq(gl2_mul(q, zh), acc[i])) bad = 2;
        }
    }
    /* FRI */
    const gl2_t fa = challenger_get_ext(ch);
    gl2_t red[2] = {gl2_from(0), gl2_from(0)};
    if (!bad) {
        for (size_t j = CT + N_QUOT; j-- > 0;) red[0] = gl2_add(gl2_mul(red[0], fa), op_local[j]);
        for (size_t j = CT; j-- > 0;) red[1] = gl2_add(gl2_mul(red[1], fa), op_next[j]);
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
    if (final_len != ((m >> (ARITY_BITS * n_layers)) >> RATE_BITS) || final_len > 64) bad = bad ? bad : 3;
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
    const gl2_t alpha_c = gl2_pow(fa, CT);
    gl_t *row = (gl_t *)malloc((CT + N_QUOT) * sizeof(gl_t));
    for (int qi = 0; qi < NUM_QUERIES && !bad; qi++) {
        size_t x = (size_t)(challenger_get(ch) % m);
        const unsigned ns = km - cap_h;
        const gl_t *row_k = Kc ? rb_take(r, Kc) : NULL;
        const gl_t *path_k = Kc ? rb_take(r, 4 * ns) : NULL;
        const gl_t *row_m = rb_take(r, C);
        const gl_t *path_m = rb_take(r, 4 * ns);
        const gl_t *row_a = rb_take(r, A);
        const gl_t *path_a = rb_take(r, 4 * ns);
        const gl_t *row_q = rb_take(r, N_QUOT);
        const gl_t *path_q = rb_take(r, 4 * ns);
        if (r->err) { bad = 1; break; }
        if (Kc && !merkle_verify(row_k, Kc, x, path_k, ns, d->const_cap, cap_h)) { bad = 5; break; }
        if (!merkle_verify(row_m, C, x, path_m, ns, cap_m, cap_h)) { bad = 5; break; }
        if (!merkle_verify(row_a, A, x, path_a, ns, cap_a, cap_h)) { bad = 5; break; }
        if (!merkle_verify(row_q, N_QUOT, x, path_q, ns, cap_q, cap_h)) { bad = 5; break; }
        if (Kc) memcpy(row, row_k, Kc * sizeof(gl_t));
        memcpy(row + Kc, row_m, C * sizeof(gl_t));
        memcpy(row + Kc + C, row_a, A * sizeof(gl_t));
        memcpy(row + CT, row_q, N_QUOT * sizeof(gl_t));
        gl_t sx = gl_mul(GL_GENERATOR, gl_pow(gl_root_of_unity(km), tmx_bitrev(x, km)));
        /* fri_combine_initial */
        gl2_t sum = gl2_from(0);
        for (int batch = 0; batch < 2; batch++) {
            const size_t npoly = batch == 0 ? CT + N_QUOT : CT;
            gl2_t acc = gl2_from(0);
            for (size_t j = npoly; j-- > 0;) acc = gl2_add(gl2_mul(acc, fa), gl2_from(row[j]));
            gl2_t num = gl2_sub(acc, red[batch]);
            gl2_t den = gl2_sub(gl2_from(sx), batch == 0 ? zeta : zeta_next);
            gl2_t sh = batch == 0 ? gl2_pow(fa, CT + N_QUOT) : alpha_c;
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
    free(row);
    free(op_local);
    free(op_next);
    return bad;
}

/* Returns 0 if the proof verifies for (input, output) under the circuit. */
int tm_verify_proof(const void *circuit, const uint64_t *proof, size_t proof_len, const uint8_t *input, size_t input_len,
                    const uint8_t out32[32]) {
    const circuit_def_t *c = (const circuit_def_t *)circuit;
    rbuf_t r = {proof, proof_len, 0, 0};
    if (rb_get(&r) != PROOF_MAGIC || rb_get(&r) != c->kind || rb_get(&r) != c->n_max || rb_get(&r) != TMX_N_TABLES) return 100;
    for (int i = 0; i < 4; i++) {
        gl_t x = 0;
        for (int j = 0; j < 8; j++) x |= (gl_t)out32[8 * i + j] << (8 * j);
        if (rb_get(&r) != x) return 101;
    }
    if (input_len != (c->kind == TMX_KIND_SKIP ? 48u : 40u)) return 102;
    for (size_t i = 0; i < proof_len; i++)
        if (proof[i] >= GL_P) return 104; /* non-canonical field element */
    challenger_t ch;
    transcript_init(&ch, c, input, input_len, out32);
    const gl_t *cap_m[TMX_N_TABLES], *cap_a[TMX_N_TABLES];
    gl2_t total[TMX_N_TABLES];
    size_t cap_words[TMX_N_TABLES];
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!c->t[t].present) continue;
        const unsigned km = c->t[t].log_n + RATE_BITS;
        cap_words[t] = 4 * ((size_t)1 << (km < CAP_HEIGHT ? km : CAP_HEIGHT));
        cap_m[t] = rb_take(&r, cap_words[t]);
        if (r.err) return 100;
        challenger_observe_many(&ch, cap_m[t], cap_words[t]);
    }
    const gl2_t beta = challenger_get_ext(&ch), gamma = challenger_get_ext(&ch);
    gl2_t balance = public_terms(c, input, out32, beta, gamma);
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!c->t[t].present) continue;
        cap_a[t] = rb_take(&r, cap_words[t]);
        total[t] = rb_get_ext(&r);
        if (r.err) return 100;
        challenger_observe_many(&ch, cap_a[t], cap_words[t]);
        observe_ext(&ch, total[t]);
        balance = gl2_add(balance, total[t]);
    }
    if (balance.a0 || balance.a1) return 200;
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!c->t[t].present) continue;
        challenger_t fork = ch;
        challenger_observe(&fork, (gl_t)t);
        int rc = verify_table_tail(c, t, cap_m[t], cap_a[t], beta, gamma, total[t], &r, &fork);
        if (rc) return 10 * (t + 1) + rc;
    }
    if (r.err || r.pos != r.n) return 103;
    return 0;
}

/* ------------------------------------------------------------------ debug / test hooks */
/* all first-round traces of a blob (tables in TMX_T_* order); out[t] must hold rows * cols words (NULL = skip) */
int tm_debug_traces(const void *circuit, const uint8_t *blob, size_t blob_len, const uint64_t *logic_trace, uint64_t *out[TMX_N_TABLES]) {
    const circuit_def_t *c = (const circuit_def_t *)circuit;
    trace_t tr[TMX_N_TABLES];
    int rc = build_all_traces(c, blob, blob_len, logic_trace, tr);
    if (!rc) rc = build_range_trace(c, tr) ? TMX_CHECK_INPUT : 0;
    for (int t = 0; t < TMX_N_TABLES && !rc; t++)
        if (out[t] && tr[t].data) memcpy(out[t], tr[t].data, tr[t].n_rows * tr[t].n_cols * sizeof(uint64_t));
    free_all_traces(tr);
    return rc;
}

/* second-round trace of one table for given first-round trace and bus challenges: aux [2 (H + 1)][n], total[2] */
void tm_debug_aux(const void *circuit, int table, const uint64_t *trace, const uint64_t beta[2], const uint64_t gamma[2], uint64_t *aux,
                  uint64_t total[2]) {
    const circuit_def_t *c = (const circuit_def_t *)circuit;
    const table_def_t *d = &c->t[table];
    trace_t tr = {(size_t)1 << d->log_n, d->n_main, (uint64_t *)trace};
    const gl2_t t = build_aux(d, &tr, gl2_make(beta[0], beta[1]), gl2_make(gamma[0], gamma[1]), aux);
    total[0] = t.a0;
    total[1] = t.a1;
}

/* LDE of the first- and second-round traces and the quotient values on the LDE coset in natural order, [2][m] */
void tm_debug_quotient(const void *circuit, int table, const uint64_t *trace, const uint64_t *aux, const uint64_t total[2],
                       const uint64_t beta[2], const uint64_t gamma[2], const uint64_t alpha[2], uint64_t *lde_main_out,
                       uint64_t *lde_aux_out, uint64_t *qv_out) {
    const circuit_def_t *c = (const circuit_def_t *)circuit;
    const table_def_t *d = &c->t[table];
    ptable_t p;
    memset(&p, 0, sizeof p);
    p.d = d;
    p.n = (size_t)1 << d->log_n;
    p.m = p.n << RATE_BITS;
    p.k = d->log_n;
    p.km = p.k + RATE_BITS;
    p.Kc = d->n_const; p.C = d->n_main; p.A = 2 * ((size_t)d->n_helpers + 1);
    p.total = gl2_make(total[0], total[1]);
    p.lde_k = (gl_t *)malloc((p.Kc ? p.Kc : 1) * p.m * sizeof(gl_t));
    if (p.Kc) ntt_lde_batch(d->constants, p.Kc, p.n, RATE_BITS, p.lde_k, NULL);
    p.lde_m = lde_main_out;
    p.lde_a = lde_aux_out;
    ntt_lde_batch(trace, p.C, p.n, RATE_BITS, p.lde_m, NULL);
    ntt_lde_batch(aux, p.A, p.n, RATE_BITS, p.lde_a, NULL);
    quotient_values(&p, gl2_make(beta[0], beta[1]), gl2_make(gamma[0], gamma[1]), alpha, qv_out);
    free(p.lde_k);
}

/* The folded constraint values (two challenges) of the transitions rows[i] -> rows[i] + 1 (cyclic) evaluated on the TRACE
 * domain itself, second-round columns included; all zero for a satisfying pair of traces.  out is [n_rows][2]. */
void tm_debug_constraints_at_rows(const void *circuit, int table, const uint64_t *trace, const uint64_t *aux, const uint64_t total[2],
                                  const uint64_t beta[2], const uint64_t gamma[2], const uint64_t alpha[2], const uint64_t *rows,
                                  size_t n_rows, uint64_t *out) {
    const circuit_def_t *c = (const circuit_def_t *)circuit;
    const table_def_t *d = &c->t[table];
    const size_t n = (size_t)1 << d->log_n, C = d->n_main, Kc = d->n_const, A = 2 * ((size_t)d->n_helpers + 1);
    const gl2_t s_over_n = gl2_scale(gl2_make(total[0], total[1]), gl_inv((gl_t)n));
    gl_t *v = (gl_t *)malloc((d->n_nodes ? d->n_nodes : 1) * sizeof(gl_t));
    gl_t *loc = (gl_t *)malloc(C * sizeof(gl_t)), *nxt = (gl_t *)malloc(C * sizeof(gl_t)), *k = (gl_t *)malloc((Kc ? Kc : 1) * sizeof(gl_t));
    gl_t *al = (gl_t *)malloc(A * sizeof(gl_t)), *an = (gl_t *)malloc(A * sizeof(gl_t)), per[16];
    for (size_t i = 0; i < n_rows; i++) {
        const size_t r = rows[i] % n, r2 = (r + 1) % n;
        for (size_t col = 0; col < C; col++) {
            loc[col] = trace[col * n + r];
            nxt[col] = trace[col * n + r2];
        }
        for (size_t col = 0; col < Kc; col++) k[col] = d->constants[col * n + r];
        for (size_t col = 0; col < A; col++) {
            al[col] = aux[col * n + r];
            an[col] = aux[col * n + r2];
        }
        periodic_at_row(d, r, per);
        circuit_eval_b(d, NULL, loc, nxt, k, per, v);
        gl_t acc[NUM_CHALLENGES];
        fold_constraints_b(d, v, al, an, gl2_make(beta[0], beta[1]), gl2_make(gamma[0], gamma[1]), s_over_n, alpha, acc);
        out[2 * i] = acc[0];
        out[2 * i + 1] = acc[1];
    }
    free(v); free(loc); free(nxt); free(k); free(al); free(an);
}
