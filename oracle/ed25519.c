/*
 * oracle/ed25519.c -- Ed25519 verification as the circuit states it, plus the witness of every modular
 * multiplication of the double-and-add ladder.  TEST INFRASTRUCTURE ONLY (see oracle/gl.h header).
 *
 * The reference verifies signatures through plonky2x's `curta_eddsa_verify_sigs_conditional`
 * [REF circuits/builder/verify.rs:248-259] (dependency source absent: plonky2x v1.0.3 @ 9df6a9db,
 * starkyx @ ad8eb4ba); the host-side sanity check is tendermint's Verifier [REF circuits/input/conversion.rs:48].
 * Restated here from RFC 8032 with the cofactorless equation [s]B == R + [h]A, h = SHA-512(R || A || M) mod l,
 * canonical encodings required (divergence 1).  Pinned against libsodium / `cryptography`
 * and all fixture signatures in tests/test_oracle_witness.py.
 *
 * Arithmetic is deliberately naive: 16 x 16-bit limbs in int64, schoolbook products, explicit carry chains --
 * the same integers the trace columns hold.
 */
#include "oracle_w.h"
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

static const int64_t P_LIMBS[16] = {0xFFED, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF,
                                    0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0x7FFF};

static void fe_from_hex(fe_t *r, const char *hex) { /* big-endian hex, 64 digits */
    uint8_t b[32];
    for (int i = 0; i < 32; i++) {
        unsigned v;
        sscanf(hex + 2 * i, "%2x", &v);
        b[31 - i] = (uint8_t)v;
    }
    for (int i = 0; i < 16; i++) r->l[i] = b[2 * i] | ((int64_t)b[2 * i + 1] << 8);
}

static fe_t FE_D, FE_2D, FE_SQRTM1, FE_ONE, FE_ZERO, FE_BX, FE_BY;
static int consts_ready = 0;
static void init_consts(void) {
    if (consts_ready) return;
    fe_from_hex(&FE_D, "52036cee2b6ffe738cc740797779e89800700a4d4141d8ab75eb4dca135978a3");
    fe_from_hex(&FE_2D, "2406d9dc56dffce7198e80f2eef3d13000e0149a8283b156ebd69b9426b2f159");
    fe_from_hex(&FE_SQRTM1, "2b8324804fc1df0b2b4d00993dfbd7a72f431806ad2fe478c4ee1b274a0ea0b0");
    fe_from_hex(&FE_BX, "216936d3cd6e53fec0a4e231fdd6dc5c692cc7609525a7b2c9562d608f25d51a");
    fe_from_hex(&FE_BY, "6666666666666666666666666666666666666666666666666666666666666658");
    memset(&FE_ZERO, 0, sizeof FE_ZERO);
    memset(&FE_ONE, 0, sizeof FE_ONE);
    FE_ONE.l[0] = 1;
    consts_ready = 1;
}

void fe_from_bytes(fe_t *r, const uint8_t b[32]) {
    for (int i = 0; i < 16; i++) r->l[i] = b[2 * i] | ((int64_t)b[2 * i + 1] << 8);
    r->l[15] &= 0x7FFF;
}
void fe_to_bytes(uint8_t b[32], const fe_t *a) {
    for (int i = 0; i < 16; i++) {
        b[2 * i] = (uint8_t)a->l[i];
        b[2 * i + 1] = (uint8_t)(a->l[i] >> 8);
    }
}

static int64_t floordiv16(int64_t x) { return x >> 16; } /* arithmetic shift = floor for negatives */

/* canonical residue of the non-negative integer sum_k t[k] * 2^(16k), n coefficients (signed) */
static void reduce_to_canonical(const int64_t *t, int n, int64_t c[16]) {
    int64_t d[40];
    int nd = n + 4;
    int64_t carry = 0;
    for (int k = 0; k < nd; k++) {
        int64_t s = (k < n ? t[k] : 0) + carry;
        carry = floordiv16(s);
        d[k] = s - carry * 65536;
    }
    /* fold digits >= 16 down with 2^256 = 38 (mod p) until they vanish */
    for (;;) {
        int any = 0;
        for (int k = 16; k < nd; k++) any |= d[k] != 0;
        if (!any) break;
        int64_t e[40] = {0};
        for (int k = 0; k < 16; k++) e[k] = d[k];
        for (int k = 16; k < nd; k++) e[k - 16] += 38 * d[k];
        carry = 0;
        for (int k = 0; k < nd; k++) {
            int64_t s = e[k] + carry;
            carry = floordiv16(s);
            d[k] = s - carry * 65536;
        }
    }
    /* value < 2^256 = 2p + 38: subtract p while >= p */
    for (int it = 0; it < 3; it++) {
        int ge = 1;
        for (int k = 15; k >= 0; k--) {
            if (d[k] > P_LIMBS[k]) break;
            if (d[k] < P_LIMBS[k]) { ge = 0; break; }
        }
        if (!ge) break;
        int64_t borrow = 0;
        for (int k = 0; k < 16; k++) {
            int64_t s = d[k] - P_LIMBS[k] + borrow;
            borrow = floordiv16(s);
            d[k] = s - borrow * 65536;
        }
    }
    for (int k = 0; k < 16; k++) c[k] = d[k];
}

/* U * V = c + q * p as limb polynomials evaluated at 2^16: see header of this file / DESIGN.md.
 *   t_k = sum_{i+j=k} U_i V_j                      (k = 0..30, signed)
 *   s_k = t_k - c_k - sum_{i+j=k} q_i p_j + w_{k-1} must be a multiple of 2^16;  w_k = s_k / 2^16;  w_31 = 0 */
void fe_mul_gadget(const int64_t U[16], const int64_t V[16], fe_mul_witness_t *out) {
    int64_t t[31] = {0};
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j++) t[i + j] += U[i] * V[j];
    reduce_to_canonical(t, 31, out->c);
    const int64_t pinv = 0x35E5; /* 0xFFED * 0x35E5 = 1 (mod 2^16) */
    int64_t carry = 0;
    memset(out->q, 0, sizeof out->q);
    for (int k = 0; k < 32; k++) {
        int64_t s = (k < 31 ? t[k] : 0) - (k < 16 ? out->c[k] : 0) + carry;
        for (int i = 0; i < k && i < 17; i++)
            if (k - i < 16) s -= out->q[i] * P_LIMBS[k - i];
        if (k < 17) {
            int64_t low = s & 0xFFFF;
            int64_t qk = (low * pinv) & 0xFFFF;
            out->q[k] = qk;
            s -= qk * P_LIMBS[0];
        }
        if (s & 0xFFFF) {
            fprintf(stderr, "oracle: fe_mul_gadget carry chain not exact at k=%d\n", k);
            abort();
        }
        carry = s >> 16;
        if (k < 31) out->w[k] = carry;
    }
    if (carry != 0) {
        fprintf(stderr, "oracle: fe_mul_gadget final carry non-zero\n");
        abort();
    }
}

static void fe_mul(fe_t *r, const fe_t *a, const fe_t *b) {
    fe_mul_witness_t w;
    fe_mul_gadget(a->l, b->l, &w);
    memcpy(r->l, w.c, sizeof w.c);
}
static void fe_sqr(fe_t *r, const fe_t *a) { fe_mul(r, a, a); }
static void fe_lin_canon(fe_t *r, const int64_t U[16]) { /* canonical residue of a non-negative limb vector */
    reduce_to_canonical(U, 16, r->l);
}
static void fe_add(fe_t *r, const fe_t *a, const fe_t *b) {
    int64_t u[16];
    for (int i = 0; i < 16; i++) u[i] = a->l[i] + b->l[i];
    fe_lin_canon(r, u);
}
static void fe_sub(fe_t *r, const fe_t *a, const fe_t *b) {
    int64_t u[16];
    for (int i = 0; i < 16; i++) u[i] = a->l[i] - b->l[i] + P_LIMBS[i];
    fe_lin_canon(r, u);
}
static int fe_eq(const fe_t *a, const fe_t *b) { return memcmp(a->l, b->l, sizeof a->l) == 0; }
static int fe_is_zero(const fe_t *a) { return fe_eq(a, &FE_ZERO); }

/* a^(2^252 - 3) = a^((p-5)/8) */
static void fe_pow22523(fe_t *r, const fe_t *a) {
    /* exponent bits: 2^252 - 3 = 0b111...101 (250 ones, then 0, then 1) */
    fe_t acc = FE_ONE;
    for (int i = 251; i >= 0; i--) {
        fe_sqr(&acc, &acc);
        int bit = (i == 1) ? 0 : 1;
        if (bit) fe_mul(&acc, &acc, a);
    }
    *r = acc;
}

void ge_identity(ge_t *r) {
    init_consts();
    r->X = FE_ZERO; r->Y = FE_ONE; r->Z = FE_ONE; r->T = FE_ZERO;
}
void ge_basepoint(ge_t *r) {
    init_consts();
    r->X = FE_BX; r->Y = FE_BY; r->Z = FE_ONE;
    fe_mul(&r->T, &FE_BX, &FE_BY);
}

int ge_decompress(ge_t *r, const uint8_t enc[32]) {
    init_consts();
    fe_t y, y2, u, v, v3, v7, x, chk, t;
    int sign = enc[31] >> 7;
    fe_from_bytes(&y, enc);
    /* canonical y required */
    {
        int ge = 1;
        for (int k = 15; k >= 0; k--) {
            if (y.l[k] > P_LIMBS[k]) break;
            if (y.l[k] < P_LIMBS[k]) { ge = 0; break; }
        }
        if (ge) return -1;
    }
    fe_sqr(&y2, &y);
    fe_sub(&u, &y2, &FE_ONE);
    fe_mul(&v, &y2, &FE_D);
    fe_add(&v, &v, &FE_ONE);
    fe_sqr(&v3, &v); fe_mul(&v3, &v3, &v);
    fe_sqr(&v7, &v3); fe_mul(&v7, &v7, &v);
    fe_mul(&t, &u, &v7);
    fe_pow22523(&t, &t);
    fe_mul(&x, &u, &v3);
    fe_mul(&x, &x, &t);
    fe_sqr(&chk, &x); fe_mul(&chk, &chk, &v);
    if (!fe_eq(&chk, &u)) {
        fe_t nu;
        fe_sub(&nu, &FE_ZERO, &u);
        if (!fe_eq(&chk, &nu)) return -1;
        fe_mul(&x, &x, &FE_SQRTM1);
    }
    if (fe_is_zero(&x) && sign) return -1;
    if ((x.l[0] & 1) != sign) fe_sub(&x, &FE_ZERO, &x);
    r->X = x; r->Y = y; r->Z = FE_ONE;
    fe_mul(&r->T, &x, &y);
    return 0;
}

void ge_compress(uint8_t enc[32], const ge_t *p) {
    init_consts();
    /* z^-1 = z^(p-2) */
    fe_t zi = FE_ONE, x, y;
    /* p - 2 = 2^255 - 21: bits 254..0 all ones except bits 4 and 2 -> ...11101011 */
    for (int i = 254; i >= 0; i--) {
        fe_sqr(&zi, &zi);
        int bit = !(i == 4 || i == 2);
        if (bit) fe_mul(&zi, &zi, &p->Z);
    }
    fe_mul(&x, &p->X, &zi);
    fe_mul(&y, &p->Y, &zi);
    fe_to_bytes(enc, &y);
    enc[31] |= (uint8_t)((x.l[0] & 1) << 7);
}

/* Operand builders: limb-wise integer linear combinations kept non-negative as integers by adding k*p. */
static void lin2(int64_t out[16], int64_t ca, const fe_t *a, int64_t cb, const fe_t *b, int64_t kp) {
    for (int i = 0; i < 16; i++) out[i] = ca * a->l[i] + cb * b->l[i] + kp * P_LIMBS[i];
}
static void lin3(int64_t out[16], int64_t ca, const fe_t *a, int64_t cb, const fe_t *b, int64_t cc, const fe_t *c, int64_t kp) {
    for (int i = 0; i < 16; i++) out[i] = ca * a->l[i] + cb * b->l[i] + cc * c->l[i] + kp * P_LIMBS[i];
}
static void wit_out(fe_t *r, const fe_mul_witness_t *w) { memcpy(r->l, w->c, sizeof w->c); }

/*
 * One ladder row.  Multiplication slots (DESIGN.md "Ed25519 table"):
 *  add-2008-hwcd-3 (a = -1), P1 = res, P2 = temp:
 *   m0 A = (Y1-X1+p)(Y2-X2+p)   m1 B = (Y1+X1)(Y2+X2)   m2 U = T1*T2   m3 C = U*2d   m4 Dh = Z1*Z2
 *   E = B-A+p  F = 2Dh-C+p  G = 2Dh+C  H = B+A
 *   m5 X3 = E*F   m6 Y3 = G*H   m7 T3 = E*H   m8 Z3 = F*G
 *  dbl-2008-hwcd (a = -1), P = temp:
 *   m9 A = X^2   m10 B = Y^2   m11 Cz = Z^2   m12 S = (X+Y)^2
 *   E = S-A-B+2p  G = B-A+p  F = B-A-2Cz+3p  H = 2p-A-B
 *   m13 X3 = E*F  m14 Y3 = G*H  m15 T3 = E*H  m16 Z3 = F*G
 */
void ge_ladder_row(const ge_t *res, const ge_t *temp, ge_t *sum, ge_t *dbl, fe_mul_witness_t wit[17]) {
    init_consts();
    int64_t u[16], v[16];
    fe_t A, B, U, C, Dh;
    lin2(u, 1, &res->Y, -1, &res->X, 1); lin2(v, 1, &temp->Y, -1, &temp->X, 1);
    fe_mul_gadget(u, v, &wit[0]); wit_out(&A, &wit[0]);
    lin2(u, 1, &res->Y, 1, &res->X, 0); lin2(v, 1, &temp->Y, 1, &temp->X, 0);
    fe_mul_gadget(u, v, &wit[1]); wit_out(&B, &wit[1]);
    fe_mul_gadget(res->T.l, temp->T.l, &wit[2]); wit_out(&U, &wit[2]);
    fe_mul_gadget(U.l, FE_2D.l, &wit[3]); wit_out(&C, &wit[3]);
    fe_mul_gadget(res->Z.l, temp->Z.l, &wit[4]); wit_out(&Dh, &wit[4]);
    int64_t E[16], F[16], G[16], H[16];
    lin2(E, 1, &B, -1, &A, 1);
    lin2(F, 2, &Dh, -1, &C, 1);
    lin2(G, 2, &Dh, 1, &C, 0);
    lin2(H, 1, &B, 1, &A, 0);
    fe_mul_gadget(E, F, &wit[5]); wit_out(&sum->X, &wit[5]);
    fe_mul_gadget(G, H, &wit[6]); wit_out(&sum->Y, &wit[6]);
    fe_mul_gadget(E, H, &wit[7]); wit_out(&sum->T, &wit[7]);
    fe_mul_gadget(F, G, &wit[8]); wit_out(&sum->Z, &wit[8]);
    fe_t A2, B2, Cz, S;
    fe_mul_gadget(temp->X.l, temp->X.l, &wit[9]); wit_out(&A2, &wit[9]);
    fe_mul_gadget(temp->Y.l, temp->Y.l, &wit[10]); wit_out(&B2, &wit[10]);
    fe_mul_gadget(temp->Z.l, temp->Z.l, &wit[11]); wit_out(&Cz, &wit[11]);
    lin2(u, 1, &temp->X, 1, &temp->Y, 0);
    fe_mul_gadget(u, u, &wit[12]); wit_out(&S, &wit[12]);
    lin3(E, 1, &S, -1, &A2, -1, &B2, 2);
    lin2(G, 1, &B2, -1, &A2, 1);
    lin3(F, 1, &B2, -1, &A2, -2, &Cz, 3);
    lin2(H, -1, &A2, -1, &B2, 2);
    fe_mul_gadget(E, F, &wit[13]); wit_out(&dbl->X, &wit[13]);
    fe_mul_gadget(G, H, &wit[14]); wit_out(&dbl->Y, &wit[14]);
    fe_mul_gadget(E, H, &wit[15]); wit_out(&dbl->T, &wit[15]);
    fe_mul_gadget(F, G, &wit[16]); wit_out(&dbl->Z, &wit[16]);
}

/* ---- joint (Straus) evaluation of [s]B + [h](-A), the row structure of the Ed25519 table (include/tmx_trace.h) ---- */
static void fe_invert(fe_t *r, const fe_t *a) { /* a^(p-2), p - 2 = 2^255 - 21 */
    fe_t acc = FE_ONE;
    for (int i = 254; i >= 0; i--) {
        fe_sqr(&acc, &acc);
        if (!(i == 4 || i == 2)) fe_mul(&acc, &acc, a);
    }
    *r = acc;
}

/* The four addends of a validator slot as the ED_ADD cells hold them: (y + x, y - x, 2 d x y) of O, B, -A, B - A, each
 * component a limb vector.  O and B are canonical constants; for -A and B - A the sums / differences are taken LIMB-WISE
 * (plus 2p where a difference could go negative), because that is the linear expression of committed cells the logic
 * table provides on the bus.  A must be affine (Z = 1).  xD, yD: affine coordinates of D = B - A. */
void ge_straus_table(const ge_t *A, ge_cached_t T[4], fe_t *xD, fe_t *yD) {
    init_consts();
    memset(T, 0, 4 * sizeof(ge_cached_t));
    T[0].ypx[0] = 1;
    T[0].ymx[0] = 1;
    fe_t t;
    ge_t B, nA, D, dummy;
    fe_mul_witness_t wit[17];
    ge_basepoint(&B);
    fe_add(&t, &B.Y, &B.X); memcpy(T[1].ypx, t.l, sizeof t.l);
    fe_sub(&t, &B.Y, &B.X); memcpy(T[1].ymx, t.l, sizeof t.l);
    fe_mul(&t, &FE_2D, &B.T); memcpy(T[1].t2d, t.l, sizeof t.l);
    fe_t tA;
    fe_mul(&tA, &FE_2D, &A->T);
    for (int i = 0; i < 16; i++) {
        T[2].ypx[i] = A->Y.l[i] - A->X.l[i] + 2 * P_LIMBS[i];
        T[2].ymx[i] = A->Y.l[i] + A->X.l[i];
        T[2].t2d[i] = 2 * P_LIMBS[i] - tA.l[i];
    }
    fe_sub(&nA.X, &FE_ZERO, &A->X);
    nA.Y = A->Y;
    nA.Z = FE_ONE;
    fe_sub(&nA.T, &FE_ZERO, &A->T);
    ge_ladder_row(&B, &nA, &D, &dummy, wit); /* D = B + (-A), projective */
    fe_t zi, tD;
    fe_invert(&zi, &D.Z);
    fe_mul(xD, &D.X, &zi);
    fe_mul(yD, &D.Y, &zi);
    fe_mul(&tD, xD, yD);
    fe_mul(&tD, &FE_2D, &tD);
    for (int i = 0; i < 16; i++) {
        T[3].ypx[i] = yD->l[i] + xD->l[i];
        T[3].ymx[i] = yD->l[i] - xD->l[i] + 2 * P_LIMBS[i];
        T[3].t2d[i] = tD.l[i];
    }
}

/* One row: out = 2 * acc + addend.  dbl-2008-hwcd (a = -1) then add-2008-hwcd-3 with Z2 = 1 and no T output:
 *   m0 A = X^2  m1 B = Y^2  m2 Cz = Z^2  m3 S = (X+Y)^2
 *   E = S-A-B+2p  G = B-A+p  F = B-A-2Cz+3p  H = 2p-A-B
 *   m4 X3 = E*F  m5 Y3 = G*H  m6 T3 = E*H  m7 Z3 = F*G
 *   m8 a = (Y3-X3+p)*(y-x)  m9 b = (Y3+X3)*(y+x)  m10 c = T3*2dxy
 *   E' = b-a+p  F' = 2Z3-c+p  G' = 2Z3+c  H' = b+a
 *   m11 X4 = E'*F'  m12 Y4 = G'*H'  m13 Z4 = F'*G' */
void ge_straus_row(const fe_t acc[3], const ge_cached_t *add, fe_t out[3], fe_mul_witness_t wit[14]) {
    init_consts();
    int64_t u[16], E[16], F[16], G[16], H[16];
    fe_t A2, B2, Cz, S, X3, Y3, T3, Z3, a, b, c;
    fe_mul_gadget(acc[0].l, acc[0].l, &wit[0]); wit_out(&A2, &wit[0]);
    fe_mul_gadget(acc[1].l, acc[1].l, &wit[1]); wit_out(&B2, &wit[1]);
    fe_mul_gadget(acc[2].l, acc[2].l, &wit[2]); wit_out(&Cz, &wit[2]);
    lin2(u, 1, &acc[0], 1, &acc[1], 0);
    fe_mul_gadget(u, u, &wit[3]); wit_out(&S, &wit[3]);
    lin3(E, 1, &S, -1, &A2, -1, &B2, 2);
    lin2(G, 1, &B2, -1, &A2, 1);
    lin3(F, 1, &B2, -1, &A2, -2, &Cz, 3);
    lin2(H, -1, &A2, -1, &B2, 2);
    fe_mul_gadget(E, F, &wit[4]); wit_out(&X3, &wit[4]);
    fe_mul_gadget(G, H, &wit[5]); wit_out(&Y3, &wit[5]);
    fe_mul_gadget(E, H, &wit[6]); wit_out(&T3, &wit[6]);
    fe_mul_gadget(F, G, &wit[7]); wit_out(&Z3, &wit[7]);
    lin2(u, 1, &Y3, -1, &X3, 1);
    fe_mul_gadget(u, add->ymx, &wit[8]); wit_out(&a, &wit[8]);
    lin2(u, 1, &Y3, 1, &X3, 0);
    fe_mul_gadget(u, add->ypx, &wit[9]); wit_out(&b, &wit[9]);
    fe_mul_gadget(T3.l, add->t2d, &wit[10]); wit_out(&c, &wit[10]);
    lin2(E, 1, &b, -1, &a, 1);
    lin2(F, 2, &Z3, -1, &c, 1);
    lin2(G, 2, &Z3, 1, &c, 0);
    lin2(H, 1, &b, 1, &a, 0);
    fe_mul_gadget(E, F, &wit[11]); wit_out(&out[0], &wit[11]);
    fe_mul_gadget(G, H, &wit[12]); wit_out(&out[1], &wit[12]);
    fe_mul_gadget(F, G, &wit[13]); wit_out(&out[2], &wit[13]);
}

/* [s]B + [h](-A) by the row recurrence (most significant bits first); result projective (X : Y : Z) */
void ge_straus(const uint8_t s[32], const uint8_t h[32], const ge_t *A, fe_t out[3]) {
    ge_cached_t T[4];
    fe_t xD, yD, acc[3], nxt[3];
    fe_mul_witness_t wit[14];
    ge_straus_table(A, T, &xD, &yD);
    acc[0] = FE_ZERO; acc[1] = FE_ONE; acc[2] = FE_ONE;
    for (int r = 0; r < 256; r++) {
        int j = 255 - r;
        int sel = ((s[j >> 3] >> (j & 7)) & 1) + 2 * ((h[j >> 3] >> (j & 7)) & 1);
        ge_straus_row(acc, &T[sel], nxt, wit);
        memcpy(acc, nxt, sizeof acc);
    }
    memcpy(out, acc, sizeof acc);
}
/* does the projective (X : Y : Z) equal the affine point R?  (X = xR Z, Y = yR Z) */
int ge_projective_equals_affine(const fe_t q[3], const ge_t *R) {
    fe_t l;
    fe_mul(&l, &R->X, &q[2]);
    if (!fe_eq(&l, &q[0])) return 0;
    fe_mul(&l, &R->Y, &q[2]);
    return fe_eq(&l, &q[1]);
}

void ge_scalarmult(ge_t *r, const uint8_t scalar[32], const ge_t *p) {
    ge_t res, temp = *p, sum, dbl;
    fe_mul_witness_t wit[17];
    ge_identity(&res);
    for (int i = 0; i < 256; i++) {
        ge_ladder_row(&res, &temp, &sum, &dbl, wit);
        if ((scalar[i >> 3] >> (i & 7)) & 1) res = sum;
        temp = dbl;
    }
    *r = res;
}

int ge_equal_projective(const ge_t *a, const ge_t *b) {
    fe_t l, r;
    fe_mul(&l, &a->X, &b->Z); fe_mul(&r, &b->X, &a->Z);
    if (!fe_eq(&l, &r)) return 0;
    fe_mul(&l, &a->Y, &b->Z); fe_mul(&r, &b->Y, &a->Z);
    return fe_eq(&l, &r);
}

/* l = 2^252 + 27742317777372353535851937790883648493 */
static const uint64_t L_LIMBS[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};

static int ge4(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}

void sc_reduce512(uint8_t out[32], const uint8_t in[64]) {
    uint64_t r[4] = {0, 0, 0, 0};
    for (int bit = 511; bit >= 0; bit--) {
        /* r = 2r + bit (r < l < 2^253 so no overflow) */
        for (int i = 3; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
        r[0] = (r[0] << 1) | ((in[bit >> 3] >> (bit & 7)) & 1);
        if (ge4(r, L_LIMBS)) {
            unsigned __int128 borrow = 0;
            for (int i = 0; i < 4; i++) {
                unsigned __int128 s = (unsigned __int128)r[i] - L_LIMBS[i] - borrow;
                r[i] = (uint64_t)s;
                borrow = (s >> 64) & 1;
            }
        }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(r[i] >> (8 * j));
}

int sc_is_canonical(const uint8_t s[32]) {
    uint64_t v[4];
    for (int i = 0; i < 4; i++) {
        v[i] = 0;
        for (int j = 0; j < 8; j++) v[i] |= (uint64_t)s[8 * i + j] << (8 * j);
    }
    return !ge4(v, L_LIMBS);
}

int ed25519_verify(const uint8_t pk[32], const uint8_t sig[64], const uint8_t *msg, size_t len) {
    ge_t A, R, B, Ps, Ph, Q, dummy;
    fe_mul_witness_t wit[17];
    if (!sc_is_canonical(sig + 32)) return 0;
    if (ge_decompress(&A, pk) != 0) return 0;
    if (ge_decompress(&R, sig) != 0) return 0;
    uint8_t buf[64 + 256], dig[64], h[32];
    if (len > 256) return 0;
    memcpy(buf, sig, 32);
    memcpy(buf + 32, pk, 32);
    memcpy(buf + 64, msg, len);
    sha512(buf, 64 + len, dig);
    sc_reduce512(h, dig);
    ge_basepoint(&B);
    ge_scalarmult(&Ps, sig + 32, &B);
    ge_scalarmult(&Ph, h, &A);
    ge_ladder_row(&Ph, &R, &Q, &dummy, wit); /* Q = Ph + R with the same addition law */
    return ge_equal_projective(&Ps, &Q);
}
