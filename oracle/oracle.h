/*
 * oracle/oracle.h -- public surface of the CPU oracle (liboracle.so).
 *
 * TEST INFRASTRUCTURE ONLY: a CPU restatement of the algorithms on the hot path named by
 * BASELINE.json.north_star, used by tests/, bench.py (cpu_baseline / --impl reference) and
 * __graft_entry__.smoke() to CHECK the CUDA product.  The product never links or loads it.
 *
 * Parity status (also in DESIGN.md):
 *   - primitives (Goldilocks, NTT, Poseidon-12, sponge, Merkle cap, duplex challenger): pinned to
 *     plonky2's known-answer vectors in tests/test_oracle_primitives.py.
 *   - witness semantics (circuits/builder, circuits/input): pinned to the reference's own test
 *     vectors and mocha-4 fixtures in tests/test_oracle_witness.py.
 *   - proof bytes: PARITY UNPINNED -- the reference's tests never pin a proof
 *     (REF circuits/skip.rs:244-249) and the Rust prover cannot be built here.
 */
#ifndef TMX_ORACLE_H
#define TMX_ORACLE_H

#include "gl.h"

#ifdef __cplusplus
extern "C" {
#endif

#define POSEIDON_WIDTH 12
#define POSEIDON_N_ROUNDS 30

/* ---- poseidon.c ---- */
const gl_t *poseidon_round_constants(void);
void poseidon_permute(gl_t s[12]);
void poseidon_hash_no_pad(const gl_t *in, size_t n, gl_t out[4]);
void poseidon_hash_or_noop(const gl_t *in, size_t n, gl_t out[4]);
void poseidon_two_to_one(const gl_t l[4], const gl_t r[4], gl_t out[4]);

typedef struct {
    size_t n_leaves, leaf_len;
    unsigned cap_height, n_levels;
    gl_t *digests; /* level 0 (leaf digests) .. cap level, concatenated */
    gl_t *cap;     /* points into digests */
} merkle_tree_t;
void merkle_build(merkle_tree_t *t, const gl_t *leaves, size_t n_leaves, size_t leaf_len, unsigned cap_height);
void merkle_free(merkle_tree_t *t);
size_t merkle_prove(const merkle_tree_t *t, size_t leaf_index, gl_t *siblings);
int merkle_verify(const gl_t *leaf, size_t leaf_len, size_t leaf_index, const gl_t *siblings, size_t n_sib,
                  const gl_t *cap, unsigned cap_height);

typedef struct {
    gl_t state[12];
    gl_t in[8];
    gl_t out[8];
    int n_in, n_out;
} challenger_t;
void challenger_init(challenger_t *c);
void challenger_observe(challenger_t *c, gl_t x);
void challenger_observe_many(challenger_t *c, const gl_t *x, size_t n);
gl_t challenger_get(challenger_t *c);
gl2_t challenger_get_ext(challenger_t *c);
gl_t challenger_pow_grind(const challenger_t *c, unsigned bits);

/* ---- ntt.c ---- */
void ntt_forward(gl_t *a, size_t n);                 /* natural in/out: X_k = sum x_j w^{jk} */
void ntt_inverse(gl_t *a, size_t n);                 /* natural in/out, scaled by 1/n */
void ntt_naive_dft(const gl_t *in, gl_t *out, size_t n); /* O(n^2) check */
void ntt_coset_forward(gl_t *a, size_t n, gl_t shift);   /* evaluate coeffs on shift*<w_n> */
void ntt_coset_inverse(gl_t *a, size_t n, gl_t shift);   /* values on shift*<w_n> -> coeffs */
/* plonky2 PolynomialBatch LDE of a column batch.
 * values: column-major [n_cols][n]; out: column-major [n_cols][n << rate_bits], BIT-REVERSED row order
 * (out[c][j] = P_c(7 * w_{n<<r}^{bitrev(j)})).  coeffs_out (optional): [n_cols][n]. */
void ntt_lde_batch(const gl_t *values, size_t n_cols, size_t n, unsigned rate_bits, gl_t *out, gl_t *coeffs_out);
/* Poseidon Merkle commitment of a column-major matrix [n_cols][n_rows]: leaf j = hash_or_noop(row j) */
void commit_columns(merkle_tree_t *t, const gl_t *cols, size_t n_cols, size_t n_rows, unsigned cap_height);

#ifdef __cplusplus
}
#endif
#endif
