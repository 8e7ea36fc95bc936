/*
 * oracle/circuit.c -- reads the build artefact (the circuit as data: table shapes, constant columns, periodic
 * columns, the constraint DAG and bus program of every table; tendermintx_b200/csrc/circuit_def.cuh documents the
 * format) and interprets the DAG over the base field and over the quadratic extension.
 * TEST INFRASTRUCTURE ONLY (see oracle/gl.h header).
 *
 * The artefact is what `builder.build()` leaves behind in the reference [REF circuits/skip.rs:165-173,
 * succinct.json:8,15 ./build/main.circuit].  The oracle does not trust it blindly: it recomputes the digest from the
 * words it read and re-commits the constant columns with its own NTT / Poseidon code; both must match the artefact.
 */
#include "circuit.h"
#include <stdlib.h>
#include <string.h>

static void hash_words(const uint64_t *w, size_t n, gl_t out[4]) {
    gl_t *t = (gl_t *)malloc((n ? n : 1) * sizeof(gl_t));
    for (size_t i = 0; i < n; i++) t[i] = w[i] % GL_P;
    poseidon_hash_no_pad(t, n, out);
    free(t);
}

void circuit_free(circuit_def_t *c) {
    for (int t = 0; t < TMX_N_TABLES; t++) {
        free(c->t[t].periodic);
        free(c->t[t].constants);
        free(c->t[t].const_cap);
        free(c->t[t].nodes);
        free(c->t[t].prog);
        free(c->t[t].bus_mask);
    }
    memset(c, 0, sizeof *c);
}

/* marks the nodes the bus items (multiplicities and tuple entries) depend on */
static void mark_bus_nodes(table_def_t *t) {
    t->bus_mask = (uint8_t *)calloc(t->n_nodes ? t->n_nodes : 1, 1);
    for (size_t p = 0; p < t->prog_len;) {
        uint64_t kind = t->prog[p++];
        if (kind == 0) {
            p++;
            continue;
        }
        for (uint64_t part = 0; part < kind; part++) {
            t->bus_mask[t->prog[p]] = 1;
            t->bus_mask[t->prog[p + 1]] = 1;
            uint64_t len = t->prog[p + 2];
            for (uint64_t i = 0; i < len; i++) t->bus_mask[t->prog[p + 3 + i]] = 1;
            p += 3 + len;
        }
    }
    for (size_t i = t->n_nodes; i-- > 0;)
        if (t->bus_mask[i] && t->nodes[i].op >= SYM_ADD) {
            t->bus_mask[t->nodes[i].a] = 1;
            t->bus_mask[t->nodes[i].b] = 1;
        }
}

/* Returns 0, or a positive diagnostic: 1 malformed, 2 digest mismatch, 3 constant-column commitment mismatch. */
int circuit_parse(const uint64_t *w, size_t nw, circuit_def_t *c) {
    memset(c, 0, sizeof *c);
    size_t pos = 0;
    int bad = 0;
#define GET() (pos < nw ? w[pos++] : (bad = 1, (uint64_t)0))
    /* words that enter the digest: everything except the constant columns (bound through their cap) */
    uint64_t *dw = (uint64_t *)malloc((nw + 1) * sizeof(uint64_t));
    size_t nd = 0;
    for (int i = 0; i < 20; i++) dw[nd++] = GET();
    if (bad || dw[0] != CIRCUIT_MAGIC || dw[4] > 50 || dw[19] != TMX_N_TABLES) {
        free(dw);
        return 1;
    }
    c->kind = (uint32_t)dw[1];
    c->n_max = (uint32_t)dw[2];
    c->skip_max = dw[3];
    c->chain_len = (size_t)dw[4];
    for (size_t k = 0; k < c->chain_len; k++) c->chain_id[k] = (uint8_t)(dw[5 + k / 8] >> (8 * (k % 8)));
    for (int i = 0; i < 6; i++) c->params[i] = dw[13 + i];
    for (int ti = 0; ti < TMX_N_TABLES && !bad; ti++) {
        table_def_t *t = &c->t[ti];
        uint64_t present = GET();
        dw[nd++] = present;
        if (!present) continue;
        uint64_t f[10];
        for (int i = 0; i < 10; i++) dw[nd++] = f[i] = GET();
        if (bad || f[0] > 28 || f[1] > (1u << 20) || f[2] > 4096 || f[3] > 64 || f[4] > (1u << 20) || f[7] > 4096 ||
            f[8] > (1u << 26) || f[9] > (1u << 28)) {
            bad = 1;
            break;
        }
        t->present = 1;
        t->log_n = (uint32_t)f[0]; t->n_main = (uint32_t)f[1]; t->n_const = (uint32_t)f[2]; t->n_per = (uint32_t)f[3];
        t->period = (uint32_t)f[4]; t->n_helpers = (uint32_t)f[5]; t->n_constraints = (uint32_t)f[6];
        t->cap_len = (size_t)f[7]; t->n_nodes = (size_t)f[8]; t->prog_len = (size_t)f[9];
        const size_t n = (size_t)1 << t->log_n;
        const size_t n_per = (size_t)t->n_per * t->period, n_cst = (size_t)t->n_const * n;
        if (pos + n_per + n_cst + t->cap_len + 2 * t->n_nodes + t->prog_len > nw) {
            bad = 1;
            break;
        }
        t->periodic = (gl_t *)malloc((n_per ? n_per : 1) * sizeof(gl_t));
        memcpy(t->periodic, w + pos, n_per * sizeof(gl_t));
        memcpy(dw + nd, w + pos, n_per * sizeof(uint64_t));
        nd += n_per; pos += n_per;
        t->constants = (gl_t *)malloc((n_cst ? n_cst : 1) * sizeof(gl_t));
        memcpy(t->constants, w + pos, n_cst * sizeof(gl_t));
        pos += n_cst;
        t->const_cap = (gl_t *)malloc((t->cap_len ? t->cap_len : 1) * sizeof(gl_t));
        memcpy(t->const_cap, w + pos, t->cap_len * sizeof(gl_t));
        memcpy(dw + nd, w + pos, t->cap_len * sizeof(uint64_t));
        nd += t->cap_len; pos += t->cap_len;
        t->nodes = (sym_node_t *)malloc((t->n_nodes ? t->n_nodes : 1) * sizeof(sym_node_t));
        for (size_t i = 0; i < t->n_nodes; i++) {
            uint64_t a = w[pos++], b = w[pos++];
            dw[nd++] = a;
            dw[nd++] = b;
            sym_node_t *nn = &t->nodes[i];
            nn->op = (uint32_t)(a & 0xFF);
            nn->deg = (uint32_t)((a >> 8) & 0xFF);
            nn->a = (uint32_t)(a >> 16);
            nn->b = nn->op == SYM_CONST ? 0 : (uint32_t)b;
            nn->val = nn->op == SYM_CONST ? b : 0;
            if (nn->op > SYM_MUL || (nn->op >= SYM_ADD && (nn->a >= i || nn->b >= i))) bad = 1;
            if (nn->op == SYM_COL && (nn->a > SRC_PERIODIC || nn->b >= (nn->a <= SRC_NEXT ? t->n_main : nn->a == SRC_CONST ? t->n_const : t->n_per)))
                bad = 1;
        }
        t->prog = (uint64_t *)malloc((t->prog_len ? t->prog_len : 1) * sizeof(uint64_t));
        memcpy(t->prog, w + pos, t->prog_len * sizeof(uint64_t));
        memcpy(dw + nd, w + pos, t->prog_len * sizeof(uint64_t));
        nd += t->prog_len; pos += t->prog_len;
        /* program sanity: node references in range, helper and constraint counts as declared */
        size_t n_emit = 0, n_help = 0;
        for (size_t p = 0; p < t->prog_len && !bad;) {
            uint64_t kind = t->prog[p++];
            if (kind == 0) {
                if (p >= t->prog_len || t->prog[p] >= t->n_nodes) bad = 1;
                p++;
                n_emit++;
            } else if (kind <= 2) {
                for (uint64_t part = 0; part < kind && !bad; part++) {
                    if (p + 3 > t->prog_len) { bad = 1; break; }
                    uint64_t len = t->prog[p + 2];
                    if (t->prog[p] >= t->n_nodes || t->prog[p + 1] >= t->n_nodes || len > 256 || p + 3 + len > t->prog_len) { bad = 1; break; }
                    for (uint64_t i = 0; i < len; i++)
                        if (t->prog[p + 3 + i] >= t->n_nodes) bad = 1;
                    p += 3 + len;
                }
                n_help++;
            } else
                bad = 1;
        }
        if (n_emit != t->n_constraints || n_help != t->n_helpers) bad = 1;
        if (!bad) mark_bus_nodes(t);
    }
    gl_t dg[4];
    for (int i = 0; i < 4; i++) dg[i] = GET();
#undef GET
    if (bad || pos != nw) {
        free(dw);
        circuit_free(c);
        return 1;
    }
    hash_words(dw, nd, c->digest);
    free(dw);
    if (memcmp(dg, c->digest, sizeof dg)) {
        circuit_free(c);
        return 2;
    }
    /* re-commit the constant columns with the oracle's own LDE / Merkle code */
    for (int ti = 0; ti < TMX_N_TABLES; ti++) {
        table_def_t *t = &c->t[ti];
        if (!t->present || !t->n_const) continue;
        const size_t n = (size_t)1 << t->log_n, m = n << 1;
        gl_t *lde = (gl_t *)malloc((size_t)t->n_const * m * sizeof(gl_t));
        ntt_lde_batch(t->constants, t->n_const, n, 1, lde, NULL);
        merkle_tree_t tree;
        commit_columns(&tree, lde, t->n_const, m, (unsigned)c->params[1]);
        const size_t cw = 4 * ((size_t)1 << tree.cap_height);
        int same = cw == t->cap_len && memcmp(tree.cap, t->const_cap, cw * sizeof(gl_t)) == 0;
        merkle_free(&tree);
        free(lde);
        if (!same) {
            circuit_free(c);
            return 3;
        }
    }
    return 0;
}

void circuit_eval_b(const table_def_t *t, const uint8_t *mask, const gl_t *local, const gl_t *next, const gl_t *k, const gl_t *per,
                    gl_t *v) {
    for (size_t i = 0; i < t->n_nodes; i++) {
        if (mask && !mask[i]) continue;
        const sym_node_t *n = &t->nodes[i];
        switch (n->op) {
            case SYM_CONST: v[i] = n->val; break;
            case SYM_COL: v[i] = n->a == SRC_LOCAL ? local[n->b] : n->a == SRC_NEXT ? next[n->b] : n->a == SRC_CONST ? k[n->b] : per[n->b]; break;
            case SYM_ADD: v[i] = gl_add(v[n->a], v[n->b]); break;
            case SYM_SUB: v[i] = gl_sub(v[n->a], v[n->b]); break;
            default: v[i] = gl_mul(v[n->a], v[n->b]); break;
        }
    }
}

void circuit_eval_e(const table_def_t *t, const gl2_t *local, const gl2_t *next, const gl2_t *k, const gl2_t *per, gl2_t *v) {
    for (size_t i = 0; i < t->n_nodes; i++) {
        const sym_node_t *n = &t->nodes[i];
        switch (n->op) {
            case SYM_CONST: v[i] = gl2_from(n->val); break;
            case SYM_COL: v[i] = n->a == SRC_LOCAL ? local[n->b] : n->a == SRC_NEXT ? next[n->b] : n->a == SRC_CONST ? k[n->b] : per[n->b]; break;
            case SYM_ADD: v[i] = gl2_add(v[n->a], v[n->b]); break;
            case SYM_SUB: v[i] = gl2_sub(v[n->a], v[n->b]); break;
            default: v[i] = gl2_mul(v[n->a], v[n->b]); break;
        }
    }
}
