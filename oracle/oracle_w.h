/*
 * oracle/oracle_w.h -- witness-side CPU oracle: SHA-2, Ed25519, Tendermint encodings, the verify_skip /
 * verify_step predicate and the trace tables.  TEST INFRASTRUCTURE ONLY (see oracle/gl.h header).
 *
 * Follows circuits/builder/{verify,validator,shared,voting}.rs, circuits/input/{conversion,tendermint_utils,
 * utils}.rs, circuits/{consts,config,variables}.rs of the reference; each function cites its lines.
 */
#ifndef TMX_ORACLE_W_H
#define TMX_ORACLE_W_H

#include <stdint.h>
#include <stddef.h>
#include "../include/tmx_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- sha2.c ---- */
typedef struct {
    uint32_t v[8]; /* a..h before the round */
    uint32_t w;
    uint64_t t1, t2; /* unreduced sums */
} sha256_round_t;
typedef struct {
    uint64_t v[8];
    uint64_t w;
} sha512_round_t;
extern const uint32_t SHA256_K[64], SHA256_IV[8];
extern const uint64_t SHA512_K[80], SHA512_IV[8];
void sha256_compress(uint32_t st[8], const uint8_t blk[64], sha256_round_t *rounds);
void sha512_compress(uint64_t st[8], const uint8_t blk[128], sha512_round_t *rounds);
size_t sha256_pad(const uint8_t *msg, size_t len, uint8_t *out);
size_t sha512_pad(const uint8_t *msg, size_t len, uint8_t *out);
void sha256(const uint8_t *msg, size_t len, uint8_t out[32]);
void sha512(const uint8_t *msg, size_t len, uint8_t out[64]);

/* ---- ed25519.c ---- */
#define FE_LIMBS 16 /* 16 x 16-bit limbs, little endian */
typedef struct {
    int64_t l[FE_LIMBS];
} fe_t; /* canonical: every limb in [0, 2^16), value < p */
typedef struct {
    fe_t X, Y, Z, T;
} ge_t;

/* witness of one modular multiplication U * V = c + q * p (see DESIGN.md "field-multiplication gadget") */
typedef struct {
    int64_t c[16]; /* canonical product limbs            */
    int64_t q[17]; /* quotient limbs, each in [0, 2^16)  */
    int64_t w[31]; /* signed carries w_0..w_30 (w_31 = 0) */
} fe_mul_witness_t;

void fe_from_bytes(fe_t *r, const uint8_t b[32]); /* low 255 bits, not reduced check */
void fe_to_bytes(uint8_t b[32], const fe_t *a);
void fe_mul_gadget(const int64_t U[16], const int64_t V[16], fe_mul_witness_t *out);
int ge_decompress(ge_t *r, const uint8_t enc[32]); /* 0 ok; -1 not on curve / non-canonical */
void ge_compress(uint8_t enc[32], const ge_t *p);
void ge_identity(ge_t *r);
void ge_basepoint(ge_t *r);
/* one double-and-add row: sum = res + temp, dbl = 2*temp; witnesses for the 17 multiplications */
void ge_ladder_row(const ge_t *res, const ge_t *temp, ge_t *sum, ge_t *dbl, fe_mul_witness_t wit[17]);
void ge_scalarmult(ge_t *r, const uint8_t scalar[32], const ge_t *p); /* LSB-first double-and-add, 256 rows */
/* joint evaluation of [s]B + [h](-A): the row structure of the Ed25519 table (include/tmx_trace.h) */
typedef struct {
    int64_t ypx[16], ymx[16], t2d[16];
} ge_cached_t;
void ge_straus_table(const ge_t *A, ge_cached_t T[4], fe_t *xD, fe_t *yD);
void ge_straus_row(const fe_t acc[3], const ge_cached_t *add, fe_t out[3], fe_mul_witness_t wit[14]);
void ge_straus(const uint8_t s[32], const uint8_t h[32], const ge_t *A, fe_t out[3]);
int ge_projective_equals_affine(const fe_t q[3], const ge_t *R);
int ge_equal_projective(const ge_t *a, const ge_t *b);
void sc_reduce512(uint8_t out[32], const uint8_t in[64]); /* 512-bit LE mod l */
int sc_is_canonical(const uint8_t s[32]);
/* cofactorless check [s]B == R + [h]A with canonical-encoding rejection; 1 = valid */
int ed25519_verify(const uint8_t pk[32], const uint8_t sig[64], const uint8_t *msg, size_t len);

/* ---- tm.c: Tendermint encodings and the circuit predicate ---- */
size_t tm_varint(uint64_t v, uint8_t out[10]);
void tm_marshal_int64_varint(uint64_t v, uint8_t out[9]);               /* REF shared.rs:67-156 */
size_t tm_marshal_validator(const uint8_t pk[32], uint64_t power, uint8_t out[46]); /* REF validator.rs:185-207 */
void tm_leaf_hash(const uint8_t *x, size_t n, uint8_t out[32]);         /* REF tendermint_utils.rs:356-362 */
void tm_inner_hash(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]);
size_t tm_split_point(size_t n);                                        /* REF tendermint_utils.rs:338-349 */
void tm_merkle_root(const uint8_t *items, const size_t *lens, size_t stride, size_t n, uint8_t out[32]);
void tm_root_from_hashed_leaves(const uint8_t *leaves, size_t n_max, size_t nb_enabled, uint8_t out[32]);
void tm_root_from_proof(const uint8_t leafhash[32], const uint8_t aunts[4][32], unsigned index, uint8_t out[32]);

/* check ids reported on failure (mirrors the assertion sites of the reference) */
enum {
    TMX_CHECK_OK = 0,
    TMX_CHECK_SKIP_DISTANCE = 1,       /* verify.rs:508-526 */
    TMX_CHECK_TRUSTED_HEADER_PROOF = 2, /* verify.rs:373-379 */
    TMX_CHECK_TRUSTED_VALHASH = 3,     /* verify.rs:381-390 */
    TMX_CHECK_TRUSTED_THRESHOLD = 4,   /* verify.rs:427-436 */
    TMX_CHECK_SIGNATURE = 5,           /* verify.rs:248-259 */
    TMX_CHECK_VALHASH = 6,             /* verify.rs:262-280 */
    TMX_CHECK_VALHASH_PROOF = 7,       /* verify.rs:283-286 */
    TMX_CHECK_THRESHOLD = 8,           /* verify.rs:289-303 */
    TMX_CHECK_SIGN_BYTES = 9,          /* validator.rs:73-153 */
    TMX_CHECK_CHAIN_ID = 10,           /* verify.rs:180-222 */
    TMX_CHECK_HEIGHT = 11,             /* shared.rs:169-207 */
    TMX_CHECK_LAST_BLOCK_ID = 12,      /* verify.rs:137-154 */
    TMX_CHECK_NEXT_VALHASH = 13,       /* verify.rs:156-178 */
    TMX_CHECK_VOTING_OVERFLOW = 14,    /* voting.rs:44-58,84-104 */
    TMX_CHECK_VARINT_SIGN = 15,        /* shared.rs:80 */
    TMX_CHECK_ROUND_SIGN = 16,         /* validator.rs:73-78 */
    TMX_CHECK_INPUT = 17
};

/* Runs verify_skip (kind 1) or verify_step (kind 0) on public input bytes + off-chain blob.
 * Returns a TMX_CHECK_* id; on success writes the 32-byte output header. */
int tm_verify_circuit(const uint8_t *input, size_t input_len, const uint8_t *blob, size_t blob_len,
                      const uint8_t *chain_id, size_t chain_id_len, uint64_t skip_max, uint8_t out32[32]);
int tm_voting_threshold(const uint64_t *power, const uint8_t *in_group, size_t n_max, size_t nb_enabled, uint64_t num,
                        uint64_t den, int *result);

extern const uint8_t TMX_DUMMY_PUBLIC_KEY[32];
extern const uint8_t TMX_DUMMY_SIGNATURE[64];
extern const uint8_t TMX_DUMMY_MESSAGE[32];

#ifdef __cplusplus
}
#endif
#endif
