/*
 * tmx_trace.h -- column layout of the three witness tables (OUR arithmetisation; the reference only names the
 * gadgets: `curta_sha256_variable`, `curta_eddsa_verify_sigs_conditional` [REF circuits/builder/verify.rs:202,248],
 * whose AIRs live in the absent starkyx crate).  Shared by the CUDA kernels, the C++ host and the CPU oracle
 * so both sides fill exactly the same cells.  Every cell is a canonical Goldilocks element; tables are
 * column-major (col * n_rows + row), n_rows a power of two.  DESIGN.md section "Trace tables" explains each
 * group and the constraints it serves.
 */
#ifndef TMX_TRACE_H
#define TMX_TRACE_H

/* ---------------- SHA-256: one row per round, 64 rows per 64-byte chunk ---------------- */
#define S256_A 0      /* 32 bits of a (LSB first), state BEFORE the round */
#define S256_B 32
#define S256_C 64
#define S256_E 96
#define S256_F 128
#define S256_G 160
#define S256_D 192    /* packed */
#define S256_H 193
#define S256_AN 194   /* 32 bits of the new a */
#define S256_EN 226   /* 32 bits of the new e */
#define S256_W 258    /* 16 packed words: w[j] = W_{t-15+j} (0 if t-15+j < 0) */
#define S256_WB14 274 /* 32 bits of w[14] */
#define S256_WB1 306  /* 32 bits of w[1]  */
#define S256_CV 338   /* 8 chaining-value words of this chunk */
#define S256_CA 346   /* 3 carry bits: T1 + T2 = an + 2^32 * ca */
#define S256_CE 349   /* 3 carry bits: d + T1 = en + 2^32 * ce */
#define S256_CW 352   /* 2 carry bits of the schedule sum s1(w[14]) + w[9] + s0(w[1]) + w[0] = ws + 2^32 * cw (all rows) */
#define S256_DG 354   /* 8 digest words, last round of the chunk only */
#define S256_DC 362   /* 8 digest carry bits, last round only */
#define S256_WS 370   /* low 32 bits of the schedule sum; equals the next row's w[15] on rows 15..62 */
#define S256_COLS 371
#define S256_ROUNDS 64

/* ---------------- SHA-512: one row per round, 128 rows per 128-byte chunk, 2 chunks per validator --------
 * Rows 0..79 of a chunk are the 80 rounds; rows 80..127 CONTINUE the round function and the message schedule with
 * round constant 0.  They cost nothing (the table was padded from 80 to a power of two anyway) and make the chunk
 * period a power of two, so the round constants and the selectors are periodic columns and every round / schedule
 * relation is enforced in the proof; the digest is taken on row 79. */
#define S512_A 0      /* 64 bits each */
#define S512_B 64
#define S512_C 128
#define S512_E 192
#define S512_F 256
#define S512_G 320
#define S512_D 384    /* lo, hi (32-bit halves) */
#define S512_H 386
#define S512_AN 388   /* 64 bits */
#define S512_EN 452
#define S512_W 516    /* 16 words x (lo, hi) */
#define S512_WB14 548 /* 64 bits */
#define S512_WB1 612
#define S512_CV 676   /* 8 words x (lo, hi) */
#define S512_CA 692   /* 3 bits carry of the low half, then 3 bits carry of the high half */
#define S512_CE 698
#define S512_CW 704   /* 2 + 2 bits */
#define S512_DG 708   /* 8 digest words x (lo, hi): zero before round 79, the digest from row 79 to the end of the chunk */
#define S512_DC 724   /* 8 words x (carry lo, carry hi), last round only */
#define S512_WS 740   /* lo, hi of the schedule sum; equals the next row's w[15] on rows 15..126 */
#define S512_TWO 742  /* 1 on all 256 rows of a validator slot whose message has two blocks (the second chunk chains from the
                         first); 0 when the second chunk is an unused compression from the IV */
#define S512_COLS 743
#define S512_ROUNDS 80
#define S512_ROWS_PER_CHUNK 128
#define S512_ROWS_PER_VALIDATOR 256

/* ---------------- Ed25519: one row per double-and-add step, 2 x 256 rows per validator ---------------- */
#define ED_BIT 0
#define ED_RES 1      /* X, Y, Z, T of the accumulator, 16 x 16-bit limbs each */
#define ED_TMP 65     /* X, Y, Z, T of the running double */
#define ED_MUL 129    /* 17 multiplication gadgets x 48 columns: c[16], q[17], w'[15] (stored + ED_W_OFFSET) */
#define ED_MUL_STRIDE 48
#define ED_MUL_Q 16
#define ED_MUL_W 33
#define ED_MUL_NW 15  /* the limb equations are checked in pairs: w'_k = carry out of limb 2k + 1, w'_15 = 0 */
#define ED_N_MUL 17
#define ED_COLS (ED_MUL + ED_N_MUL * ED_MUL_STRIDE) /* 945 */
#define ED_W_OFFSET (1 << 22)
#define ED_ROWS_PER_VALIDATOR 512
/* Row 0 of every 256-row ladder starts from res = O = (0, 1, 1, 0); the first ladder of a validator ([s]B, rows 0..255)
 * doubles the base point B = (BX, BY, 1, BX*BY), the second ([h]A, rows 256..511) an affine point (Z = 1).  16-bit limbs. */
#define ED_BASE_X_LIMBS {0xd51a, 0x8f25, 0x2d60, 0xc956, 0xa7b2, 0x9525, 0xc760, 0x692c, 0xdc5c, 0xfdd6, 0xe231, 0xc0a4, 0x53fe, 0xcd6e, 0x36d3, 0x2169}
#define ED_BASE_Y_LIMBS {0x6658, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666}
#define ED_BASE_T_LIMBS {0xdda3, 0xa5b7, 0x8ab3, 0x6dde, 0x52f5, 0x7751, 0x9f80, 0x20f0, 0xe37d, 0x64ab, 0x4e8e, 0x66ea, 0x7665, 0xd78b, 0x5f0f, 0x6787}

#endif
