// Logic table: the plain-gate gadgets of verify_skip / verify_step [REF circuits/builder/verify.rs:137-563,
// validator.rs:73-253, shared.rs:43-215, voting.rs:29-110] and the glue between the hash / signature tables, as one more AIR on
// the shared bus.  Also the table-generic accessors (rows, columns, helper counts, constant columns, AIR dispatch) the prover,
// the verifier and the circuit builder use for all five tables.
//
// A row of the table is one GADGET INSTANCE; its type is fixed by the circuit shape (constant selector columns), its cells
// are laid out per type in a shared pool of 25 groups x 16 cells (+ 24 free cells).  Every group of every row is range checked
// on the bus with the bit width its row type needs (constant columns LGK_TAG / LGK_USE), so byte, limb and flag cells need no
// constraints of their own.  Values move between rows (and to / from the hash and curve tables) only as bus messages, with
// ids and multiplicities in constant columns: the wiring is part of the circuit, not of the witness.
//
//   H1    one-chunk SHA-256 message of variable length: validator leaf 00 || marshal(pubkey, power) [validator.rs:185-229,
//         shared.rs:67-156], and the header-proof leaves (hash field, chain id, height) [verify.rs:180-222, shared.rs:169-207];
//         validator rows also carry enabled / signed flags, the running voting-power sums and, on a set's last row, the
//         threshold check [voting.rs:31-109, verify.rs:439-467], and the public-key matching of trusted against signed target
//         validators as a lookup [verify.rs:392-418]
//   H2    two-chunk SHA-256 message of fixed layout: inner node 01 || L || R of the validator-set trees (with the
//         promote-left rule for disabled children [validator.rs:231-252]) and of the header proofs [shared.rs:43-65], and the
//         72-byte last-block-id leaf of the step circuit [verify.rs:137-154]
//   SIG   SHA-512 message R || A || M of one validator with variable-length padding, the dummy substitution for unsigned
//         slots and the sign-bytes checks [validator.rs:73-183]
//   SC    scalars of one signature: h = SHA-512 digest mod l, s < l, h < l, limbs to the Ed25519 table
//   XS    canonical x / y and sign of the decompressed A and R
//   MUL   two generic multiplications c = U * V + W (mod 2^255 - 19) per row, operands = small linear combinations of up to
//         six field elements received from the bus: decompression equations, addend B - A, final check [s]B - [h]A == R
//   EDIO  per validator: provides the four addends to the Ed25519 table, receives its result
//   GLOB  height / header / round shared by all signature rows;  CFE  constant field elements
#pragma once
#include "air.cuh"

namespace tmx {

// ------------------------------------------------------------------------------------------ layout
constexpr int LG_GROUPS = 25, LG_GROUP = 16, LG_FREE = LG_GROUPS * LG_GROUP, LG_NFREE = 24, LG_COLS = LG_FREE + LG_NFREE;
// constant columns
constexpr int LGK_TAG = 0;                    // [25] range tag of the group's cells on this row
constexpr int LGK_USE = LGK_TAG + LG_GROUPS;  // [25] 1: the group's cells are looked up
constexpr int LGK_SEL = LGK_USE + LG_GROUPS;  // row type selectors
enum { LT_H1 = 0, LT_H2, LT_SIG, LT_SC, LT_XS, LT_MUL, LT_EDIO, LT_GLOB, LT_CFE, LT_COUNT };
constexpr int LGK_P = LGK_SEL + LT_COUNT;     // per-row parameters; every row type has its own index range, so a parameter is
                                              // zero on the rows of the other types and can gate a constraint by itself
constexpr int PH1 = 0, PH2 = 17, PSG = 30, PSC = 37, PXS = 38, PEI = 46, PCF = 61, PMU = 79;
constexpr int LG_NPARAM = PMU + 74;
constexpr int LGK_COLS = LGK_P + LG_NPARAM;

// ---- H1 cells / parameters
constexpr int H1_D = 0, H1_G7 = 55, H1_G72 = 64, H1_DB = 73, H1_IL = 112, H1_NZ = 168, H1_E = 176, H1_SGN = 177, H1_FLAG = 178;
constexpr int H1_TOT = 192, H1_TOT4 = 196, H1_SUM = 197, H1_SUM4 = 201, H1_DF = 202, H1_DF2 = 206;
constexpr int H1_GINV = LG_FREE, H1_GT = LG_FREE + 8, H1_MK = LG_FREE + 16;
enum { H1P_CID = PH1, H1P_SEND, H1P_FVAL, H1P_FHASH, H1P_FCHAIN, H1P_FHEIGHT, H1P_OUT_ID, H1P_OUT_MULT, H1P_IN_ID, H1P_IN_MULT,
       H1P_TARGET, H1P_TRUSTED, H1P_SETNEXT, H1P_SETFIRST, H1P_LAST2, H1P_LAST1, H1P_VID };
// ---- H2
constexpr int H2_MB = 0, H2_DB = 73, H2_EL = 112, H2_ER = 113, H2_FF = 114;
enum { H2P_CID = PH2, H2P_SEND, H2P_IS65, H2P_IS73, H2P_ID_L, H2P_RECV_L, H2P_ABS_L, H2P_ID_R, H2P_RECV_R, H2P_ABS_R, H2P_OUT_ID,
       H2P_OUT_MULT, H2P_PUBPREV };
// ---- SIG
constexpr int SG_D = 0, SG_DG = 188, SG_A31 = 252, SG_A31X2 = 253, SG_R31 = 254, SG_R31X2 = 255, SG_IL = 256, SG_SGN = 381, SG_RZ = 382,
              SG_SA = 383, SG_SR = 384;
enum { SGP_VID = PSG, SGP_YA_ID, SGP_YA_MULT, SGP_YR_ID, SGP_YR_MULT, SGP_SA_ID, SGP_SR_ID };
// ---- SC
constexpr int SC_S = 0, SC_HL = 32, SC_Q = 48, SC_WLO = 65, SC_WHI = 80, SC_DG = 96, SC_DS = 160, SC_CS = 176, SC_DH = 192, SC_CH = 208;
constexpr int SC_W_OFFSET = 1 << 21;
enum { SCP_VID = PSC };
// ---- XS
constexpr int XS_X = 0, XS_Y = 16, XS_DX = 32, XS_DY = 48, XS_CX = 64, XS_CY = 80, XS_PAR = 95, XS_STRIDE = 96, XS_XH = 192, XS_XH2 = 193,
              XS_SIGN = 208;
enum { XSP_X_ID = PXS, XSP_X_MULT, XSP_Y_ID, XSP_SIGN_ID };
constexpr int XSP_STRIDE = 4;
// ---- MUL (two gadgets per row)
constexpr int MU_STRIDE = 192, MU_IN = 0, MU_U = 96, MU_V = 112, MU_C = 128, MU_Q = 144, MU_WLO = 161, MU_WHI = 176;
enum { MUP_ACTIVE = PMU, MUP_IN_ID = PMU + 1, MUP_IN_RECV = PMU + 7, MUP_CU = PMU + 13, MUP_CV = PMU + 19, MUP_CW = PMU + 25, MUP_KPU = PMU + 31,
       MUP_KPV, MUP_KPW, MUP_OUT_ID, MUP_OUT_MULT, MUP_ASSERT };
constexpr int MUP_STRIDE = 37;
// ---- EDIO
constexpr int EI_YA = 0, EI_XA = 16, EI_T2A = 32, EI_XD = 48, EI_YD = 64, EI_T2D = 80, EI_XQ = 96, EI_YQ = 112, EI_ZQ = 128, EI_M = LG_FREE;
enum { EIP_VID = PEI, EIP_YA_ID, EIP_XA_ID, EIP_T2A_ID, EIP_T2D_ID, EIP_XD_ID, EIP_XD_MULT, EIP_YD_ID, EIP_YD_MULT, EIP_XQ_ID, EIP_XQ_MULT,
       EIP_YQ_ID, EIP_YQ_MULT, EIP_ZQ_ID, EIP_ZQ_MULT };
// ---- GLOB
constexpr int GB_HB = 0, GB_HDR = 8, GB_RB = 40, GB_RB7X2 = 48, GB_RZ = 64, GB_RINV = LG_FREE, GB_MG = LG_FREE + 1;
// ---- CFE
constexpr int CF_C = 0;
enum { CFP_LIMB = PCF, CFP_ID = PCF + 16, CFP_MULT = PCF + 17 };

// ids
TMX_HD uint64_t nid_tree(uint32_t set, uint32_t level, uint32_t index) { return ((uint64_t)(set + 1) << 24) | ((uint64_t)level << 16) | index; }
TMX_HD uint64_t nid_header(uint32_t proof, uint32_t level) { return ((uint64_t)0x7F << 24) | ((uint64_t)proof << 8) | level; }
TMX_HD uint64_t wid_slot(uint32_t slot, uint32_t local) { return ((uint64_t)0x40 << 24) | ((uint64_t)slot << 8) | local; }
TMX_HD uint64_t wid_const(uint32_t idx) { return ((uint64_t)0x50 << 24) | idx; }
// per-slot wires
enum { W_YA = 0, W_XA, W_YR, W_XR, W_XD, W_YD, W_XQ, W_YQ, W_ZQ, W_SA, W_SR, W_G0 = 16 /* outputs of gadget g: W_G0 + g */ };
// constant wires
enum { WC_ONE = 0, WC_D, WC_D2, WC_BYMX, WC_BYPX, WC_BT2D, WC_COUNT };
constexpr int LG_GADGETS_PER_SLOT = 22, LG_MUL_ROWS_PER_SLOT = 11;

// l = 2^252 + 27742317777372353535851937790883648493 in 16-bit limbs
TMX_HD uint64_t ell_limb(int i) {
    const uint64_t L[16] = {0xd3ed, 0x5cf5, 0x631a, 0x5812, 0x9cd6, 0xa2f7, 0xf9de, 0x14de, 0, 0, 0, 0, 0, 0, 0, 0x1000};
    return L[i];
}

// dummy key / signature of the unsigned slots (RFC 8032 key of the all-zero seed signing 32 zero bytes), as constants of the AIR
TMX_HD uint64_t dummy_pk_byte(int i) {
    const uint8_t t[32] = {0x3b, 0x6a, 0x27, 0xbc, 0xce, 0xb6, 0xa4, 0x2d, 0x62, 0xa3, 0xa8, 0xd0, 0x2a, 0x6f, 0x0d, 0x73,
                           0x65, 0x32, 0x15, 0x77, 0x1d, 0xe2, 0x43, 0xa6, 0x3a, 0xc0, 0x48, 0xa1, 0x8b, 0x59, 0xda, 0x29};
    return t[i];
}
TMX_HD uint64_t dummy_sig_byte(int i) {
    const uint8_t t[64] = {0x3d, 0xa1, 0xeb, 0xdf, 0xa9, 0x6e, 0xdd, 0x18, 0x1d, 0xbe, 0x36, 0x59, 0xd1, 0xc0, 0x51, 0xc4, 0x31, 0xf0, 0x56, 0xa5, 0xad, 0x6a,
                           0x97, 0xa6, 0x0d, 0x5c, 0xca, 0x10, 0x46, 0x04, 0x38, 0x78, 0x35, 0x46, 0x46, 0x1e, 0x31, 0x28, 0x5f, 0xc5, 0x9f, 0x91, 0xc7, 0x07,
                           0x26, 0x42, 0x74, 0x50, 0x61, 0xe2, 0x45, 0x1d, 0x5f, 0xf3, 0x3b, 0xcc, 0xd8, 0xc3, 0xc7, 0x4d, 0xab, 0xca, 0xf6, 0x0a};
    return t[i];
}
// canonical limbs of the base point's cached form: c = 0: By + Bx, 1: By - Bx, 2: 2 d Bx By
TMX_HD uint64_t ed_base_cached_limb(int c, int j) {
    const uint16_t t[3][16] = {
        {0x3b85, 0xf58c, 0x93c6, 0x2fbc, 0x0e19, 0xfb8c, 0x2dc6, 0xcf93, 0x42c2, 0x643d, 0x4898, 0x270b, 0xba65, 0x33d4, 0x9d3a, 0x07cf},
        {0x913e, 0xd740, 0x3905, 0x9d10, 0xbeb3, 0xd140, 0x9f05, 0xfd39, 0x8a09, 0x688f, 0x8434, 0xa5c1, 0x1267, 0x98f8, 0x2f92, 0x44fd},
        {0xaa68, 0x877a, 0x1205, 0xabc9, 0xc49e, 0xccaa, 0xe823, 0x26d9, 0x598c, 0xdd43, 0x7dcb, 0x5a1b, 0x65a8, 0x9f0c, 0x7b68, 0x6f11}};
    return t[c][j];
}
// limbs of the constant wires WC_*
TMX_HD uint64_t logic_const_wire_limb(int w, int j) {
    const uint16_t D[16] = {0x78a3, 0x1359, 0x4dca, 0x75eb, 0xd8ab, 0x4141, 0x0a4d, 0x0070, 0xe898, 0x7779, 0x4079, 0x8cc7, 0xfe73, 0x2b6f, 0x6cee, 0x5203};
    const uint16_t D2[16] = {0xf159, 0x26b2, 0x9b94, 0xebd6, 0xb156, 0x8283, 0x149a, 0x00e0, 0xd130, 0xeef3, 0x80f2, 0x198e, 0xfce7, 0x56df, 0xd9dc, 0x2406};
    if (w == 0) return j == 0;
    if (w == 1) return D[j];
    if (w == 2) return D2[j];
    if (w == 3) return ed_base_cached_limb(1, j);
    if (w == 4) return ed_base_cached_limb(0, j);
    return ed_base_cached_limb(2, j);
}

TMX_HD size_t logic_used_rows(AirShape sh) {
    const size_t np = air_pow2_at_least(sh.n_max), sets = sh.kind == 1 ? 2 : 1, proofs = sh.kind == 1 ? 4 : 5;
    return 1 + WC_COUNT + sets * ((size_t)sh.n_max + (np - 1)) + proofs * 5 + (size_t)sh.n_max * (4 + LG_MUL_ROWS_PER_SLOT);
}
#if TMX_BUS_LINKS
TMX_HD size_t logic_rows(AirShape sh) { return air_pow2_at_least(logic_used_rows(sh)); }
#else
TMX_HD size_t logic_rows(AirShape) { return 0; }
#endif
TMX_HD int logic_cols(AirShape sh) { return logic_rows(sh) ? LG_COLS : 0; }
TMX_HD int logic_const_cols(AirShape) { return LGK_COLS; }
constexpr int LG_HELPERS = LG_FREE / 2 + 5 + 4 + 7 + 17 + 3 + 8 + 7 + 1 + 1;  // range pairs, then H1, H2, SIG, SC, XS, MUL, EDIO, GLOB, CFE
TMX_HD int logic_helpers(AirShape) { return LG_HELPERS; }

// U * V + W = c + q * p over the integers on 16-bit limbs, limb equations in pairs with committed carries (cells at l[g0 + ..]);
// gate multiplies every equation.  Same shape as ed_mul_gadget plus the additive operand W.
template <class F, class Row, class Emit>
TMX_HD void logic_mul_gadget(const F U[16], const F V[16], const F W[16], const Row& l, int c0, int q0, int wlo0, int whi0, F gate, Emit& emit) {
    const F off = F::c(ED_W_OFFSET), two16 = F::c(1 << 16), two32 = F::c(1ULL << 32);
    F wprev = F::c(0);
    for (int K = 0; K < 16; K++) {
        F s = F::c(0);
        for (int h = 0; h < 2; h++) {
            const int k = 2 * K + h;
            F e = F::c(0);
            for (int i = 0; i < 16; i++) {
                const int j = k - i;
                if (j >= 0 && j < 16) e = e + U[i] * V[j];
            }
            if (k < 16) e = (e + W[k]) - l[c0 + k];
            for (int i = 0; i < 17; i++) {
                const int j = k - i;
                if (j >= 0 && j < 16) e = e - l[q0 + i] * F::c(p25519_limb(j));
            }
            s = h ? s + two16 * e : e;
        }
        if (K >= 1) s = s + wprev;
        if (K < 15) {
            wprev = (l[wlo0 + K] + two16 * l[whi0 + K]) - off;
            s = s - two32 * wprev;
        }
        emit(gate * s);
    }
}

// v < K for a 16-limb value v and a constant K: v + d + 1 = K with range-checked d and carry bits c (cells)
template <class F, class Row, class Limb, class KLimb, class Emit>
TMX_HD void logic_less_than(const Limb& v, const KLimb& kl, const Row& l, int d0, int c0, F gate, Emit& emit) {
    for (int i = 0; i < 16; i++) {
        const F cin = i == 0 ? F::c(1) : l[c0 + i - 1];
        F s = (v(i) + l[d0 + i] + cin) - F::c(kl(i));
        if (i < 15) s = s - F::c(1 << 16) * l[c0 + i];
        emit(gate * s);
    }
}

template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_logic(AirShape sh, const Row& l, const Row& n, const KRow& k, const Per&, Emit& emit, Bus& bus) {
    auto P = [&](int i) { return k[LGK_P + i]; };
    auto SEL = [&](int t) { return k[LGK_SEL + t]; };
    const F one = F::c(1), zero = F::c(0);
    auto be32 = [&](int c0) { return ((l[c0] * F::c(1 << 24) + l[c0 + 1] * F::c(1 << 16)) + l[c0 + 2] * F::c(1 << 8)) + l[c0 + 3]; };
    auto le32 = [&](int c0) { return ((l[c0 + 3] * F::c(1 << 24) + l[c0 + 2] * F::c(1 << 16)) + l[c0 + 1] * F::c(1 << 8)) + l[c0]; };
    auto le16 = [&](int c0) { return l[c0] + l[c0 + 1] * F::c(256); };
    auto limbs4 = [&](const Row& r, int c0) {
        return ((r[c0 + 3] * F::c(1ULL << 48) + r[c0 + 2] * F::c(1ULL << 32)) + r[c0 + 1] * F::c(1 << 16)) + r[c0];
    };
    // =========================================================================================== H1
    const F s1 = SEL(LT_H1), fval = P(H1P_FVAL), fhash = P(H1P_FHASH), fchain = P(H1P_FCHAIN), fheight = P(H1P_FHEIGHT);
    F h1w[16];
    {
        F il_sum = zero, len8 = zero, keep = zero;
        F byte[56];
        for (int j = 55; j >= 0; j--) {  // keep_j = [message length > j]
            const F il = l[H1_IL + j];
            il_sum = il_sum + il;
            len8 = len8 + il * F::c(8 * j);
            byte[j] = j < 55 ? keep * l[H1_D + j] + il * F::c(128) : il * F::c(128);
            keep = keep + il;
        }
        emit(s1 * (il_sum - one));
        for (int w = 0; w < 14; w++)
            h1w[w] = ((byte[4 * w] * F::c(1 << 24) + byte[4 * w + 1] * F::c(1 << 16)) + byte[4 * w + 2] * F::c(1 << 8)) + byte[4 * w + 3];
        h1w[14] = zero;
        h1w[15] = len8;
    }
    // varint of the value V = sum g_k 128^k [shared.rs:67-156]: byte k = g_k + 128 * [a higher septet is non-zero]
    F val = zero;
    for (int kk = 8; kk >= 0; kk--) val = val * F::c(128) + l[H1_G7 + kk];
    auto nz = [&](const Row& r, int kk) { return kk >= 1 && kk <= 8 ? r[H1_NZ + kk - 1] : F::c(0); };
    for (int kk = 0; kk < 9; kk++) emit(s1 * (l[H1_G72 + kk] - (l[H1_G7 + kk] + l[H1_G7 + kk])));  // septets: g and 2 g are bytes
    for (int kk = 1; kk <= 7; kk++) emit(s1 * (nz(l, kk + 1) * (one - nz(l, kk))));
    for (int kk = 1; kk <= 8; kk++) {
        emit(s1 * (l[H1_G7 + kk] * (one - nz(l, kk))));
        emit(s1 * (l[H1_GT + kk - 1] - l[H1_G7 + kk] * l[H1_GINV + kk - 1]));
        emit(s1 * ((nz(l, kk) - nz(l, kk + 1)) * (l[H1_GT + kk - 1] - one)));  // the highest non-zero septet is non-zero
    }
    auto vbyte = [&](int kk) { return l[H1_G7 + kk] + nz(l, kk + 1) * F::c(128); };
    // validator leaf: 00 0a 22 0a 20 <pubkey> 10 <varint>, 39..47 message bytes
    {
        const int fixed_pos[6] = {0, 1, 2, 3, 4, 37};
        const uint64_t fixed_val[6] = {0, 0x0a, 0x22, 0x0a, 0x20, 0x10};
        for (int i = 0; i < 6; i++) emit(fval * (l[H1_D + fixed_pos[i]] - F::c(fixed_val[i])));
        for (int kk = 0; kk < 9; kk++) emit(fval * (l[H1_D + 38 + kk] - vbyte(kk)));
        F in_range = zero;
        for (int p = 39; p <= 47; p++) in_range = in_range + l[H1_IL + p];
        emit(fval * (in_range - one));
    }
    // hash leaf of a header proof: 00 0a 20 <hash>
    emit(fhash * l[H1_D]);
    emit(fhash * (l[H1_D + 1] - F::c(0x0a)));
    emit(fhash * (l[H1_D + 2] - F::c(0x20)));
    emit(fhash * (l[H1_IL + 35] - one));
    // chain id leaf: 00 <enc>, enc[2 .. 2 + L) = the circuit's chain id [verify.rs:214-221]
    emit(fchain * l[H1_D]);
    for (int j = 0; j < (int)sh.chain_len && j < 50; j++) emit(fchain * (l[H1_D + 3 + j] - F::c((uint8_t)sh.chain[j])));
    // height leaf: 00 08 <varint(height)> [shared.rs:169-207]
    emit(fheight * l[H1_D]);
    emit(fheight * (l[H1_D + 1] - F::c(8)));
    for (int kk = 0; kk < 9; kk++) emit(fheight * (l[H1_D + 2 + kk] - vbyte(kk)));
    // flags and voting power [voting.rs:31-109]
    const F e = l[H1_E], sgn = l[H1_SGN], flag = l[H1_FLAG];
    emit((s1 - fval) * (e - one));
    emit(s1 * (sgn * (one - e)));
    // a trusted validator counts only from an enabled slot: the validators hash does not bind the fields of the padding slots
    // (the reference sums over every flagged slot, verify.rs:392-431 -- a padding slot with a signer's key and a large
    // power would meet the 1/3 threshold by itself)
    emit(s1 * (flag * (one - e)));
    emit(sgn * (s1 - P(H1P_TARGET)));
    emit(flag * (s1 - P(H1P_TRUSTED)));
    emit(s1 * (l[H1_MK] * (one - sgn)));
    emit(P(H1P_SETNEXT) * (n[H1_E] * (one - e)));
    emit(s1 * (l[H1_TOT4] - F::c(4) * l[H1_TOT + 3]));
    emit(s1 * (l[H1_SUM4] - F::c(4) * l[H1_SUM + 3]));
    emit(s1 * (l[H1_DF2] - F::c(2) * l[H1_DF + 3]));
    {
        F nval = zero;
        for (int kk = 8; kk >= 0; kk--) nval = nval * F::c(128) + n[H1_G7 + kk];
        const F tot = limbs4(l, H1_TOT), sum = limbs4(l, H1_SUM), ntot = limbs4(n, H1_TOT), nsum = limbs4(n, H1_SUM);
        emit(P(H1P_SETFIRST) * (tot - e * val));
        emit(P(H1P_SETFIRST) * (sum - (sgn + flag) * val));
        emit(P(H1P_SETNEXT) * ((ntot - tot) - n[H1_E] * nval));
        emit(P(H1P_SETNEXT) * ((nsum - sum) - (n[H1_SGN] + n[H1_FLAG]) * nval));
        // strictly more than 2/3 (target) or 1/3 (trusted) of the total [verify.rs:439-467]: 3 sum - num * total - 1 = diff >= 0
        const F df = limbs4(l, H1_DF), s3 = (sum + sum + sum) - one;
        emit(P(H1P_LAST2) * ((s3 - (tot + tot)) - df) + P(H1P_LAST1) * ((s3 - tot) - df));
    }
    F h1dw[8], pkw[8];
    for (int i = 0; i < 8; i++) {
        h1dw[i] = be32(H1_DB + 4 * i);
        pkw[i] = be32(H1_D + 5 + 4 * i);
    }
    // =========================================================================================== H2
    const F s2 = SEL(LT_H2), is65 = P(H2P_IS65), is73 = P(H2P_IS73);
    emit(s2 * (l[H2_MB] - is65));
    for (int j = 65; j < 73; j++) emit(is65 * l[H2_MB + j]);
    const F el = l[H2_EL], er = l[H2_ER], ff = l[H2_FF];
    emit(s2 * (ff - el * er));
    emit(P(H2P_ABS_L) * el);
    emit(P(H2P_ABS_R) * er);
    emit(((s2 - P(H2P_RECV_L)) - P(H2P_ABS_L)) * (el - one));
    emit(((s2 - P(H2P_RECV_R)) - P(H2P_ABS_R)) * (er - one));
    F h2a[16], h2b[16], h2dw[8], lw[8], rw[8];
    for (int w = 0; w < 16; w++) h2a[w] = be32(H2_MB + 4 * w);
    for (int w = 0; w < 16; w++) h2b[w] = zero;
    h2b[0] = be32(H2_MB + 64) + is65 * F::c(0x80 << 16);
    h2b[1] = be32(H2_MB + 68);
    h2b[2] = l[H2_MB + 72] * F::c(1 << 24) + is73 * F::c(0x80 << 16);
    h2b[15] = is65 * F::c(65 * 8) + is73 * F::c(73 * 8);
    for (int i = 0; i < 8; i++) {
        h2dw[i] = be32(H2_DB + 4 * i);
        lw[i] = be32(H2_MB + 1 + 4 * i);
        rw[i] = be32(H2_MB + 33 + 4 * i);
    }
    // =========================================================================================== SIG
    const F s3 = SEL(LT_SIG), ssgn = l[SG_SGN], rz = l[SG_RZ];
    F sg_half[2][32];  // [chunk][tuple order: 2 j = low half of word j (bytes 8 j + 4 ..), 2 j + 1 = high half (bytes 8 j ..)]
    F two_blocks = zero;
    {
        F il_sum = zero, keep = zero, len_lo = zero, len_hi = zero;
        // bytes 64 .. 188 of the padded message depend on the length p = 64 + len: byte_j = [p > j] D_j + 0x80 [p == j]
        F byte[256];
        for (int j = 255; j >= 189; j--) byte[j] = zero;
        for (int j = 188; j >= 64; j--) {
            const F il = l[SG_IL + j - 64];
            il_sum = il_sum + il;
            if (j >= 112) { two_blocks = two_blocks + il; len_hi = len_hi + il * F::c(8 * j); }
            else len_lo = len_lo + il * F::c(8 * j);
            byte[j] = j < 188 ? keep * l[SG_D + j] + il * F::c(128) : il * F::c(128);
            keep = keep + il;
        }
        for (int j = 63; j >= 0; j--) byte[j] = l[SG_D + j];
        emit(s3 * (il_sum - one));
        for (int c = 0; c < 2; c++)
            for (int w = 0; w < 16; w++) {
                const int b = 128 * c + 8 * w;
                sg_half[c][2 * w + 1] = ((byte[b] * F::c(1 << 24) + byte[b + 1] * F::c(1 << 16)) + byte[b + 2] * F::c(1 << 8)) + byte[b + 3];
                sg_half[c][2 * w] = ((byte[b + 4] * F::c(1 << 24) + byte[b + 5] * F::c(1 << 16)) + byte[b + 6] * F::c(1 << 8)) + byte[b + 7];
            }
        sg_half[0][30] = sg_half[0][30] + len_lo;  // bit length in the last word of the last block
        sg_half[1][30] = sg_half[1][30] + len_hi;
    }
    // unsigned slots verify the dummy key / signature over 32 zero bytes [conversion.rs:99-133]
    {
        const F unsigned_row = s3 * (one - ssgn);
        for (int j = 0; j < 32; j++) emit(unsigned_row * (l[SG_D + j] - F::c(dummy_sig_byte(j))));
        for (int j = 0; j < 32; j++) emit(unsigned_row * (l[SG_D + 32 + j] - F::c(dummy_pk_byte(j))));
        for (int j = 0; j < 32; j++) emit(unsigned_row * l[SG_D + 64 + j]);
        emit(unsigned_row * (l[SG_IL + 32] - one));
    }
    emit(s3 * (l[SG_D + 63] - (l[SG_A31] + l[SG_SA] * F::c(128))));
    emit(s3 * (l[SG_A31X2] - (l[SG_A31] + l[SG_A31])));
    emit(s3 * (l[SG_D + 31] - (l[SG_R31] + l[SG_SR] * F::c(128))));
    emit(s3 * (l[SG_R31X2] - (l[SG_R31] + l[SG_R31])));
    // sign bytes of a signed validator: precommit, then height / round / header hash through the GLOB message [validator.rs:73-183]
    emit(s3 * (ssgn * (l[SG_D + 64 + 1] - F::c(8))));
    emit(s3 * (ssgn * (l[SG_D + 64 + 2] - F::c(2))));
    // =========================================================================================== SC
    const F s4 = SEL(LT_SC);
    {
        // digest = q * l + h on 16-bit limbs (the digest is a little-endian 512-bit integer), limb equations in pairs
        const F off = F::c(SC_W_OFFSET), two16 = F::c(1 << 16), two32 = F::c(1ULL << 32);
        F wprev = zero;
        for (int K = 0; K < 16; K++) {
            F s = zero;
            for (int h = 0; h < 2; h++) {
                const int kk = 2 * K + h;
                F ev = le16(SC_DG + 2 * kk);
                if (kk < 16) ev = ev - l[SC_HL + kk];
                for (int i = 0; i < 17; i++) {
                    const int j = kk - i;
                    if (j >= 0 && j < 16 && ell_limb(j)) ev = ev - l[SC_Q + i] * F::c(ell_limb(j));
                }
                s = h ? s + two16 * ev : ev;
            }
            if (K >= 1) s = s + wprev;
            if (K < 15) {
                wprev = (l[SC_WLO + K] + two16 * l[SC_WHI + K]) - off;
                s = s - two32 * wprev;
            }
            emit(s4 * s);
        }
        logic_less_than<F>([&](int i) { return le16(SC_S + 2 * i); }, ell_limb, l, SC_DS, SC_CS, s4, emit);
        logic_less_than<F>([&](int i) { return l[SC_HL + i]; }, ell_limb, l, SC_DH, SC_CH, s4, emit);
    }
    // =========================================================================================== XS
    const F s5 = SEL(LT_XS);
    for (int a = 0; a < 2; a++) {
        const int o = a * XS_STRIDE;
        logic_less_than<F>([&](int i) { return l[o + XS_X + i]; }, p25519_limb, l, o + XS_DX, o + XS_CX, s5, emit);
        logic_less_than<F>([&](int i) { return l[o + XS_Y + i]; }, p25519_limb, l, o + XS_DY, o + XS_CY, s5, emit);
        emit(s5 * (l[o + XS_X] - ((l[XS_XH + 2 * a] + l[XS_XH + 2 * a]) + l[o + XS_PAR])));
        emit(s5 * (l[XS_XH2 + 2 * a] - (l[XS_XH + 2 * a] + l[XS_XH + 2 * a])));
        emit(s5 * (l[o + XS_PAR] - l[XS_SIGN + a]));
    }
    // =========================================================================================== MUL
    const F s7 = SEL(LT_MUL);
    for (int g = 0; g < 2; g++) {
        const int o = g * MU_STRIDE, q = g * MUP_STRIDE;
        F U[16], V[16], W[16];
        for (int i = 0; i < 16; i++) {
            const F pl = F::c(p25519_limb(i));
            F u = P(q + MUP_KPU) * pl, v = P(q + MUP_KPV) * pl, w = P(q + MUP_KPW) * pl;
            for (int s = 0; s < 6; s++) {
                const F in = l[o + MU_IN + 16 * s + i];
                u = u + P(q + MUP_CU + s) * in;
                v = v + P(q + MUP_CV + s) * in;
                w = w + P(q + MUP_CW + s) * in;
            }
            emit(s7 * (l[o + MU_U + i] - u));
            emit(s7 * (l[o + MU_V + i] - v));
            U[i] = l[o + MU_U + i];
            V[i] = l[o + MU_V + i];
            W[i] = w;
        }
        logic_mul_gadget<F>(U, V, W, l, o + MU_C, o + MU_Q, o + MU_WLO, o + MU_WHI, s7, emit);
        for (int i = 0; i < 16; i++) emit(P(q + MUP_ASSERT) * l[o + MU_C + i]);
    }
    // =========================================================================================== GLOB
    const F s8 = SEL(LT_GLOB);
    {
        F srb = zero;
        for (int i = 0; i < 8; i++) srb = srb + l[GB_RB + i];
        emit(s8 * (l[GB_RB7X2] - (l[GB_RB + 7] + l[GB_RB + 7])));  // round is a non-negative int64 [validator.rs:73-78]
        emit(s8 * ((l[GB_RZ] + srb * l[GB_RINV]) - one));
        emit(s8 * (l[GB_RZ] * srb));
    }
    // =========================================================================================== CFE
    const F s9 = SEL(LT_CFE);
    for (int i = 0; i < 16; i++) emit(s9 * (l[CF_C + i] - P(CFP_LIMB + i)));

    // =========================================================================================== bus
    // range checks: every cell of every group, with the row type's tag
    {
        LookupPairs<F, Bus> rc(bus);
        for (int g = 0; g < LG_GROUPS; g++) {
            const F tag = k[LGK_TAG + g], m = zero - k[LGK_USE + g];
            for (int i = 0; i < LG_GROUP; i++) rc.push(tag, m, l[g * LG_GROUP + i]);
        }
        rc.flush();
    }
    const F tNODE = bus_tag<F>(BUS_NODE), tFE = bus_tag<F>(BUS_FE), tMSG = bus_tag<F>(BUS_MSG256), tDIG = bus_tag<F>(BUS_DIG256);
    // ---- H1 (5)
    bus.one(tMSG, P(H1P_SEND), 17, [&](int i) { return i == 0 ? P(H1P_CID) : h1w[i - 1]; });
    bus.two(tDIG, zero - P(H1P_SEND), 9, [&](int i) { return i == 0 ? P(H1P_CID) : h1dw[i - 1]; },
            tNODE, P(H1P_OUT_MULT), 10, [&](int i) { return i == 0 ? P(H1P_OUT_ID) : (i == 9 ? e : h1dw[i - 1]); });
    bus.two(tNODE, zero - P(H1P_IN_MULT), 10, [&](int i) { return i == 0 ? P(H1P_IN_ID) : (i == 9 ? one : be32(H1_D + 3 + 4 * (i - 1))); },
            bus_tag<F>(BUS_KEY), s1 * l[H1_MK], 8, [&](int i) { return pkw[i]; });
    bus.two(bus_tag<F>(BUS_KEY), zero - s1 * flag, 8, [&](int i) { return pkw[i]; },
            bus_tag<F>(BUS_PUB), zero - fheight, 10, [&](int i) { return i == 0 ? F::c(PUB_HEIGHT) : l[H1_G7 + i - 1]; });
    bus.one(bus_tag<F>(BUS_PKSIG), P(H1P_TARGET), 10, [&](int i) { return i == 0 ? P(H1P_VID) : (i == 9 ? sgn : sgn * pkw[i - 1]); });
    // ---- H2 (4)
    bus.two(tMSG, P(H2P_SEND), 17, [&](int i) { return i == 0 ? P(H2P_CID) : h2a[i - 1]; },
            tMSG, P(H2P_SEND), 17, [&](int i) { return i == 0 ? P(H2P_CID) + one : h2b[i - 1]; });
    bus.two(tDIG, zero - P(H2P_SEND), 9, [&](int i) { return i == 0 ? P(H2P_CID) + one : h2dw[i - 1]; },
            tNODE, zero - P(H2P_RECV_L), 10, [&](int i) { return i == 0 ? P(H2P_ID_L) : (i == 9 ? el : lw[i - 1]); });
    bus.two(tNODE, zero - P(H2P_RECV_R), 10, [&](int i) { return i == 0 ? P(H2P_ID_R) : (i == 9 ? er : rw[i - 1]); },
            bus_tag<F>(BUS_PUB), zero - P(H2P_PUBPREV), 9, [&](int i) { return i == 0 ? F::c(PUB_PREV) : be32(H2_MB + 3 + 4 * (i - 1)); });
    bus.one(tNODE, P(H2P_OUT_MULT), 10, [&](int i) { return i == 0 ? P(H2P_OUT_ID) : (i == 9 ? el : lw[i - 1] + ff * (h2dw[i - 1] - lw[i - 1])); });
    // ---- SIG (8)
    const F vid3 = P(SGP_VID);
    for (int c = 0; c < 2; c++)
        bus.one(bus_tag<F>(BUS_MSG512), s3, 35,
                [&](int i) { return i == 0 ? vid3 : (i == 1 ? F::c((uint64_t)c) : (i == 2 ? two_blocks : sg_half[c][i - 3])); });
    bus.two(bus_tag<F>(BUS_DIG512), zero - s3, 17,
            [&](int i) { return i == 0 ? vid3 : ((i - 1) & 1 ? be32(SG_DG + 8 * ((i - 1) >> 1)) : be32(SG_DG + 8 * ((i - 1) >> 1) + 4)); },
            bus_tag<F>(BUS_DIGB), s3, 17, [&](int i) { return i == 0 ? vid3 : le32(SG_DG + 4 * (i - 1)); });
    bus.one(bus_tag<F>(BUS_GLOB), zero - s3 * ssgn, 13, [&](int i) {
        const int m0 = SG_D + 64;
        if (i < 2) return le32(m0 + 4 + 4 * i);
        if (i < 10) return rz * be32(m0 + 16 + 4 * (i - 2)) + (one - rz) * be32(m0 + 25 + 4 * (i - 2));
        if (i < 12) return (one - rz) * le32(m0 + 13 + 4 * (i - 10));
        return rz;
    });
    bus.one(bus_tag<F>(BUS_PKSIG), zero - s3, 10, [&](int i) { return i == 0 ? vid3 : (i == 9 ? ssgn : ssgn * be32(SG_D + 32 + 4 * (i - 1))); });
    bus.two(tFE, P(SGP_YA_MULT), 17, [&](int i) { return i == 0 ? P(SGP_YA_ID) : (i == 16 ? l[SG_D + 62] + l[SG_A31] * F::c(256) : le16(SG_D + 32 + 2 * (i - 1))); },
            tFE, P(SGP_YR_MULT), 17, [&](int i) { return i == 0 ? P(SGP_YR_ID) : (i == 16 ? l[SG_D + 30] + l[SG_R31] * F::c(256) : le16(SG_D + 2 * (i - 1))); });
    bus.two(bus_tag<F>(BUS_BIT), s3, 2, [&](int i) { return i == 0 ? P(SGP_SA_ID) : l[SG_SA]; },
            bus_tag<F>(BUS_BIT), s3, 2, [&](int i) { return i == 0 ? P(SGP_SR_ID) : l[SG_SR]; });
    // ---- SC (17)
    const F vid4 = P(SCP_VID);
    for (int i2 = 0; i2 < 16; i2 += 2)
        bus.two(bus_tag<F>(BUS_SCALAR), zero - s4, 4, [&](int i) { return i == 0 ? vid4 : (i == 1 ? zero : (i == 2 ? F::c((uint64_t)i2) : le16(SC_S + 2 * i2))); },
                bus_tag<F>(BUS_SCALAR), zero - s4, 4, [&](int i) { return i == 0 ? vid4 : (i == 1 ? zero : (i == 2 ? F::c((uint64_t)i2 + 1) : le16(SC_S + 2 * i2 + 2))); });
    for (int i2 = 0; i2 < 16; i2 += 2)
        bus.two(bus_tag<F>(BUS_SCALAR), zero - s4, 4, [&](int i) { return i == 0 ? vid4 : (i == 1 ? one : (i == 2 ? F::c((uint64_t)i2) : l[SC_HL + i2])); },
                bus_tag<F>(BUS_SCALAR), zero - s4, 4, [&](int i) { return i == 0 ? vid4 : (i == 1 ? one : (i == 2 ? F::c((uint64_t)i2 + 1) : l[SC_HL + i2 + 1])); });
    bus.one(bus_tag<F>(BUS_DIGB), zero - s4, 17, [&](int i) { return i == 0 ? vid4 : le32(SC_DG + 4 * (i - 1)); });
    // ---- XS (3)
    for (int a = 0; a < 2; a++) {
        const int o = a * XS_STRIDE, q = a * XSP_STRIDE;
        bus.two(tFE, P(q + XSP_X_MULT), 17, [&](int i) { return i == 0 ? P(q + XSP_X_ID) : l[o + XS_X + i - 1]; },
                tFE, zero - s5, 17, [&](int i) { return i == 0 ? P(q + XSP_Y_ID) : l[o + XS_Y + i - 1]; });
    }
    bus.two(bus_tag<F>(BUS_BIT), zero - s5, 2, [&](int i) { return i == 0 ? P(XSP_SIGN_ID) : l[XS_SIGN]; },
            bus_tag<F>(BUS_BIT), zero - s5, 2, [&](int i) { return i == 0 ? P(XSP_SIGN_ID + XSP_STRIDE) : l[XS_SIGN + 1]; });
    // ---- MUL (8)
    for (int g = 0; g < 2; g++) {
        const int o = g * MU_STRIDE, q = g * MUP_STRIDE;
        for (int s = 0; s < 6; s += 2)
            bus.two(tFE, zero - P(q + MUP_IN_RECV + s), 17, [&](int i) { return i == 0 ? P(q + MUP_IN_ID + s) : l[o + MU_IN + 16 * s + i - 1]; },
                    tFE, zero - P(q + MUP_IN_RECV + s + 1), 17, [&](int i) { return i == 0 ? P(q + MUP_IN_ID + s + 1) : l[o + MU_IN + 16 * (s + 1) + i - 1]; });
        bus.one(tFE, P(q + MUP_OUT_MULT), 17, [&](int i) { return i == 0 ? P(q + MUP_OUT_ID) : l[o + MU_C + i - 1]; });
    }
    // ---- EDIO (7)
    const F s6 = SEL(LT_EDIO), vid6 = P(EIP_VID);
    auto fe_recv = [&](int idp, int c0) { return [&, idp, c0](int i) { return i == 0 ? P(idp) : l[c0 + i - 1]; }; };
    bus.two(tFE, zero - s6, 17, fe_recv(EIP_YA_ID, EI_YA), tFE, zero - s6, 17, fe_recv(EIP_XA_ID, EI_XA));
    bus.two(tFE, zero - s6, 17, fe_recv(EIP_T2A_ID, EI_T2A), tFE, zero - s6, 17, fe_recv(EIP_T2D_ID, EI_T2D));
    bus.two(tFE, P(EIP_XD_MULT), 17, fe_recv(EIP_XD_ID, EI_XD), tFE, P(EIP_YD_MULT), 17, fe_recv(EIP_YD_ID, EI_YD));
    bus.two(bus_tag<F>(BUS_EDRES), zero - s6, 49, [&](int i) { return i == 0 ? vid6 : l[EI_XQ + i - 1]; }, tFE, P(EIP_XQ_MULT), 17, fe_recv(EIP_XQ_ID, EI_XQ));
    bus.two(tFE, P(EIP_YQ_MULT), 17, fe_recv(EIP_YQ_ID, EI_YQ), tFE, P(EIP_ZQ_MULT), 17, fe_recv(EIP_ZQ_ID, EI_ZQ));
    {
        auto addend = [&](int which) {
            return [&, which](int i) {
                if (i == 0) return vid6;
                if (i == 1) return F::c((uint64_t)which);
                const int c = (i - 2) >> 4, j = (i - 2) & 15;  // component (y + x, y - x, 2dxy), limb
                const F p2 = F::c(2 * p25519_limb(j));
                if (which == 0) return F::c(c < 2 && j == 0 ? 1 : 0);
                if (which == 1) return F::c(ed_base_cached_limb(c, j));
                if (which == 2) return c == 0 ? (l[EI_YA + j] - l[EI_XA + j]) + p2 : (c == 1 ? l[EI_YA + j] + l[EI_XA + j] : p2 - l[EI_T2A + j]);
                return c == 0 ? l[EI_YD + j] + l[EI_XD + j] : (c == 1 ? (l[EI_YD + j] - l[EI_XD + j]) + p2 : l[EI_T2D + j]);
            };
        };
        const F tADD = bus_tag<F>(BUS_ADDEND);
        bus.two(tADD, s6 * l[EI_M], 50, addend(0), tADD, s6 * l[EI_M + 1], 50, addend(1));
        bus.two(tADD, s6 * l[EI_M + 2], 50, addend(2), tADD, s6 * l[EI_M + 3], 50, addend(3));
    }
    // ---- GLOB (1), CFE (1)
    auto glob_tuple = [&](int i) {
        if (i < 2) return le32(GB_HB + 4 * i);
        if (i < 10) return be32(GB_HDR + 4 * (i - 2));
        if (i < 12) return le32(GB_RB + 4 * (i - 10));
        return l[GB_RZ];
    };
    bus.two(bus_tag<F>(BUS_PUB), zero - s8, 11, [&](int i) { return i == 0 ? F::c(PUB_GLOB) : glob_tuple(i - 1); },
            bus_tag<F>(BUS_GLOB), s8 * l[GB_MG], 13, glob_tuple);
    bus.one(tFE, P(CFP_MULT), 17, [&](int i) { return i == 0 ? P(CFP_ID) : l[CF_C + i - 1]; });
}

// The verifier's own bus terms: the messages that tie the tables to the public input (trusted height / header, target
// height) and to the public output (the proven header).  Defined in logic_plan.cu.
gl2 logic_public_terms(AirShape sh, uint64_t skip_max, const uint8_t* input, const uint8_t* out32, gl2 beta, gl2 gamma);
uint64_t logic_const_value(int kc, size_t row, AirShape sh);

// ---- table-generic accessors ----
TMX_HD size_t air_table_rows(int t, AirShape sh) { return t == AIR_LOGIC ? logic_rows(sh) : air_rows(t, sh); }
TMX_HD int air_table_cols(int t, AirShape sh) { return t == AIR_LOGIC ? logic_cols(sh) : air_cols(t); }
TMX_HD int air_table_const_cols(int t, AirShape sh) { return t == AIR_LOGIC ? logic_const_cols(sh) : air_const_cols(t); }
TMX_HD int air_table_helpers(int t, AirShape sh) { return t == AIR_LOGIC ? logic_helpers(sh) : air_helpers(t); }
TMX_HD int air_table_aux_cols(int t, AirShape sh) { return air_table_cols(t, sh) ? 2 * (air_table_helpers(t, sh) + 1) : 0; }
inline uint64_t air_table_const_value(int t, int kc, size_t row, AirShape sh) {
    return t == AIR_LOGIC ? logic_const_value(kc, row, sh) : air_const_value(t, kc, row, sh);
}
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_eval_any(int table, AirShape sh, const Row& l, const Row& n, const KRow& k, const Per& per, Emit& emit, Bus& bus) {
    if (table == AIR_LOGIC) air_logic<F>(sh, l, n, k, per, emit, bus);
    else air_eval<F>(table, l, n, k, per, emit, bus);
}

}  // namespace tmx
