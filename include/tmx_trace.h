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

/* ---------------- Ed25519: one row per scalar-bit PAIR, 256 rows per validator ----------------
 * Joint (Straus) evaluation of Q = [s]B + [h](-A), most significant bits first:
 *     acc' = 2 * acc + T[bs + 2 bh],   T = {O, B, -A, B - A}  (affine, cached as (y + x, y - x, 2 d x y))
 * The addend of the row is looked up on the bus (ED_BUS_ADDEND: validator id, selector, 48 limbs) from the four
 * entries the logic table provides per validator.  dbl-2008-hwcd (8 multiplications, T3 kept for the addition) then
 * add-2008-hwcd-3 with Z2 = 1 and no T output (6 multiplications): 14 multiplication gadgets per row.
 * Every gadget cell is range checked on the bus: c, q, wlo in [0, 2^16), whi in [0, 2^11).
 * Field elements are 16 x 16-bit limbs, little endian. */
#define ED_BS 0       /* bit of s */
#define ED_BH 1       /* bit of h */
#define ED_SACC_S 2   /* bits of s seen so far inside the current 16-row group (MSB first); 0 on the group's first row */
#define ED_SACC_H 3
#define ED_ACC 4      /* X, Y, Z of the accumulator BEFORE the row's step */
#define ED_ADD 52     /* the row's addend: y + x, y - x, 2 d x y */
#define ED_MUL 100    /* 14 multiplication gadgets x 63 columns */
#define ED_MUL_STRIDE 63
#define ED_MUL_Q 16   /* q[17] */
#define ED_MUL_WLO 33 /* low 16 bits of (carry + ED_W_OFFSET), 15 carries (the limb equations are checked in pairs) */
#define ED_MUL_WHI 48 /* high bits of the same, below 2^11 */
#define ED_MUL_NW 15
#define ED_N_MUL 14
#define ED_COLS (ED_MUL + ED_N_MUL * ED_MUL_STRIDE) /* 982 */
#define ED_W_OFFSET (1 << 22)
#define ED_ROWS_PER_VALIDATOR 256
/* gadget slots */
#define ED_G_A 0   /* X * X */
#define ED_G_B 1   /* Y * Y */
#define ED_G_CZ 2  /* Z * Z */
#define ED_G_S 3   /* (X + Y)^2 */
#define ED_G_X3 4  /* E * F  (doubled point) */
#define ED_G_Y3 5  /* G * H */
#define ED_G_T3 6  /* E * H */
#define ED_G_Z3 7  /* F * G */
#define ED_G_AA 8  /* (Y3 - X3) * (y - x) */
#define ED_G_BB 9  /* (Y3 + X3) * (y + x) */
#define ED_G_CC 10 /* T3 * 2dxy */
#define ED_G_X4 11 /* E' * F' */
#define ED_G_Y4 12 /* G' * H' */
#define ED_G_Z4 13 /* F' * G' */
/* constant (preprocessed) columns of the Ed25519 table, fixed by the circuit shape */
#define EDK_VID 0     /* validator slot of the row (row / 256) */
#define EDK_ACTIVE 1  /* 1 on the rows of slots < n_max */
#define EDK_SEND16 2  /* active and last row of a 16-row group: a scalar limb is complete */
#define EDK_LIMB 3    /* index of the scalar limb of the row's group: 15 - (row % 256) / 16 */
#define EDK_LAST 4    /* active and last row of the slot: the result leaves on the bus */
#define EDK_COLS 5
#define ED_BASE_X_LIMBS {0xd51a, 0x8f25, 0x2d60, 0xc956, 0xa7b2, 0x9525, 0xc760, 0x692c, 0xdc5c, 0xfdd6, 0xe231, 0xc0a4, 0x53fe, 0xcd6e, 0x36d3, 0x2169}
#define ED_BASE_Y_LIMBS {0x6658, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666, 0x6666}
#define ED_BASE_T_LIMBS {0xdda3, 0xa5b7, 0x8ab3, 0x6dde, 0x52f5, 0x7751, 0x9f80, 0x20f0, 0xe37d, 0x64ab, 0x4e8e, 0x66ea, 0x7665, 0xd78b, 0x5f0f, 0x6787}

/* constant columns of the SHA-256 table */
#define S256K_FIRST 0  /* row 0 of a chunk that starts a message: chaining value = IV */
#define S256K_LINK 1   /* row 63 of a chunk whose successor continues the message: next chaining value = this digest */
#define S256K_CID 2    /* chunk index (row / 64) */
#define S256K_MSG 3    /* used chunk, row 15: the 16 message words are received from the bus */
#define S256K_DIG 4    /* used chunk, row 63: the digest is sent on the bus */
#define S256K_COLS 5
/* constant columns of the SHA-512 table */
#define S512K_VID 0    /* validator slot (row / 256) */
#define S512K_CHUNK 1  /* 0 / 1: chunk inside the slot */
#define S512K_MSG 2    /* active slot, row 15 of a chunk */
#define S512K_DIG0 3   /* active slot, row 79 of the first chunk */
#define S512K_DIG1 4   /* active slot, row 79 of the second chunk */
#define S512K_COLS 5

/* ---------------- range table: 2^16 rows, row t provides the value t ---------------- */
#define RG_M16 0  /* multiplicity of t among the 16-bit lookups */
#define RG_M11 1  /* ... among the 11-bit lookups (zero on rows >= 2^11) */
#define RG_M8 2   /* ... among the 8-bit lookups (zero on rows >= 2^8) */
#define RG_M1 3   /* ... among the 1-bit lookups (zero on rows >= 2) */
#define RG_COLS 4
#define RGK_T 0
#define RGK_S11 1
#define RGK_S8 2
#define RGK_S1 3
#define RGK_COLS 4
#define RG_LOG_ROWS 16

/* ---------------- bus tags (first element of every message fingerprint) ---------------- */
#define BUS_R16 1      /* (v) */
#define BUS_R11 2      /* (v) */
#define BUS_R8 3       /* (v) */
#define BUS_R1 4       /* (v) */
#define BUS_MSG256 5   /* (chunk, w0..w15) */
#define BUS_DIG256 6   /* (chunk, d0..d7) */
#define BUS_MSG512 7   /* (slot, chunk, two, 32 halves) */
#define BUS_DIG512 8   /* (slot, 16 halves) */
#define BUS_ADDEND 9   /* (slot, selector, 48 limbs) */
#define BUS_SCALAR 10  /* (slot, which, limb index, limb) */
#define BUS_EDRES 11   /* (slot, X[16], Y[16], Z[16]) */
#define BUS_NODE 12    /* (node id, 8 big-endian words, enabled): value of a Merkle node */
#define BUS_KEY 13     /* (8 words): public key of a target validator that signed */
#define BUS_PKSIG 14   /* (slot, signed * 8 key words, signed): target validator -> its signature row */
#define BUS_GLOB 15    /* (2 height words LE, 8 header words, 2 round words LE, round == 0) */
#define BUS_FE 16      /* (wire id, 16 limbs): a curve25519 field element between logic rows */
#define BUS_BIT 17     /* (wire id, bit) */
#define BUS_DIGB 18    /* (slot, 16 little-endian words of the SHA-512 digest) */
#define BUS_PUB 19     /* (kind, ...): provided by the VERIFIER from the public input / output */
#define PUB_GLOB 1     /* (height LE words [2], target / next header words [8]) */
#define PUB_HEIGHT 2   /* (height as 9 septets) */
#define PUB_PREV 3     /* (previous header words [8]), step only */

#define TMX_N_TABLES 5
#define TMX_T_SHA256 0
#define TMX_T_SHA512 1
#define TMX_T_ED 2
#define TMX_T_LOGIC 3
#define TMX_T_RANGE 4

#endif
