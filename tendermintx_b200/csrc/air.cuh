// Constraint systems (AIR) of the SHA-256, SHA-512, Ed25519, logic and range tables, generic over the field: the
// quotient kernels (K5) instantiate them on base-field LDE values, the verifier on extension-field openings at zeta,
// the circuit builder on symbolic nodes (circuit_def.cu: that is how the build artefact gets its constraint DAG).
//
// The arithmetisation is this repo's own -- upstream the equivalent lives in starkyx's SHA-256 / SHA-512 /
// Ed25519 AIRs and plonky2's gates, reached by the reference through `curta_sha256_variable`,
// `curta_eddsa_verify_sigs_conditional` and the plain-gate gadgets of circuits/builder
// [REF circuits/builder/verify.rs:202,248-259].  Rules: every constraint has degree <= 3 in (trace, constant,
// periodic) columns; a constraint that reads the next row carries a selector that is zero on the last row of the
// table; a table first emits its own constraints, then declares its bus interactions (bus.one / bus.two, in a
// fixed order: it defines the helper columns of the second commitment round); emission order defines the alpha powers.
//
// Bus (logUp): a message is (tag, v_0 .. v_{k-1}) with fingerprint f = gamma + tag + beta v_0 + beta^2 v_1 + ...
// in the quadratic extension; a table row contributes multiplicity / f for each of its interactions (positive: sends /
// provides, negative: receives / looks up) and the sum over all rows of all tables plus the verifier's public
// terms must vanish.
#pragma once
#include "gl.cuh"
#include "../../include/tmx_trace.h"
#include <vector>
#include <algorithm>

namespace tmx {

// ---- field wrappers with operators ----
// (plain aggregates on purpose: no constructors, so local arrays of them carry no hidden initialisation)
struct FB {
    gl v;
    TMX_HD static FB mk(gl x) { FB r; r.v = x; return r; }
    TMX_HD static FB c(uint64_t x) { return mk((gl)x); }
};
TMX_HD FB operator+(FB a, FB b) { return FB::mk(gl_add(a.v, b.v)); }
TMX_HD FB operator-(FB a, FB b) { return FB::mk(gl_sub(a.v, b.v)); }
TMX_HD FB operator*(FB a, FB b) { return FB::mk(gl_mul(a.v, b.v)); }

struct FE {
    gl2 v;
    TMX_HD static FE mk(gl2 x) { FE r; r.v = x; return r; }
    TMX_HD static FE c(uint64_t x) { return mk(gl2_from((gl)x)); }
};
TMX_HD FE operator+(FE a, FE b) { return FE::mk(gl2_add(a.v, b.v)); }
TMX_HD FE operator-(FE a, FE b) { return FE::mk(gl2_sub(a.v, b.v)); }
TMX_HD FE operator*(FE a, FE b) { return FE::mk(gl2_mul(a.v, b.v)); }

// Horner accumulator over the two constraint challenges: acc <- acc * alpha + c
// acc * alpha + c
template <class F>
TMX_HD F horner_step(F acc, F alpha, F c) { return acc * alpha + c; }
// base field (quotient kernels): one lazily reduced multiply-add; the accumulator stays in [0, 2^64) and whoever reads it
// multiplies it once more (by 1 / Z_H), which canonicalises
TMX_HD FB horner_step(FB acc, FB alpha, FB c) { return FB::mk(gl_mac_nc(acc.v, alpha.v, c.v)); }
template <class F>
struct ConstraintAcc {
    F acc0, acc1, alpha0, alpha1;
    TMX_HD void operator()(F c) {
        acc0 = horner_step(acc0, alpha0, c);
        acc1 = horner_step(acc1, alpha1, c);
    }
};

template <class F>
TMX_HD F is_bool(F x) { return x * (x - F::c(1)); }
template <class F>
TMX_HD F xor3(F a, F b, F c) {
    F ab = a * b;
    F pairs = ab + b * c + c * a;
    return (a + b + c) - (pairs + pairs) + F::c(4) * (ab * c);
}
// sum_i 2^i row[col0 + i]
template <class F, class Row>
TMX_HD F pack_bits(const Row& r, int col0, int nbits) {
    F acc = F::c(0);
    for (int i = nbits - 1; i >= 0; i--) acc = acc + acc + r[col0 + i];
    return acc;
}

// two-component algebra F[X] / (X^2 - 7) over any of the field wrappers: with F = FB it is the quadratic extension, with
// F = FE (openings at zeta of the two component polynomials of an extension-valued column) the same identities hold
// component-wise, which is all the constraint check needs
template <class F>
struct Ext2 {
    F a0, a1;
};
template <class F>
TMX_HD Ext2<F> e2_mk(F a0, F a1) { Ext2<F> r; r.a0 = a0; r.a1 = a1; return r; }
template <class F>
TMX_HD Ext2<F> e2_add(Ext2<F> a, Ext2<F> b) { return e2_mk<F>(a.a0 + b.a0, a.a1 + b.a1); }
template <class F>
TMX_HD Ext2<F> e2_sub(Ext2<F> a, Ext2<F> b) { return e2_mk<F>(a.a0 - b.a0, a.a1 - b.a1); }
// 7 x (the non-residue of the extension)
template <class F>
TMX_HD F f_mul7(F x) { return F::c(7) * x; }
TMX_HD FB f_mul7(FB x) { return FB::mk(gl_mul7(x.v)); }
template <class F>
TMX_HD Ext2<F> e2_mul(Ext2<F> a, Ext2<F> b) {
    return e2_mk<F>(a.a0 * b.a0 + f_mul7(a.a1 * b.a1), a.a0 * b.a1 + a.a1 * b.a0);
}
template <class F>
TMX_HD Ext2<F> e2_scale(Ext2<F> a, F s) { return e2_mk<F>(a.a0 * s, a.a1 * s); }

// fingerprint of (tag, tup(0) .. tup(len - 1)): gamma + tag + beta (v_0 + beta (v_1 + ...))
// (Horner from the last element; its first step multiplies beta by a base-field value: two products, not an extension product)
template <class F, class Tup>
TMX_HD Ext2<F> bus_fingerprint(Ext2<F> beta, Ext2<F> gamma, F tag, int len, const Tup& tup) {
    Ext2<F> acc = e2_scale<F>(beta, tup(len - 1));
    for (int i = len - 2; i >= 0; i--) {
        acc.a0 = acc.a0 + tup(i);
        acc = e2_mul<F>(acc, beta);
    }
    acc.a0 = acc.a0 + tag;
    return e2_add<F>(acc, gamma);
}

// single-value lookups (range checks) are declared through this pairing buffer: two per helper column
template <class F, class Bus>
struct LookupPairs {
    Bus& bus;
    bool have, unit0;
    F tag0, m0, v0;
    TMX_HD LookupPairs(Bus& b) : bus(b), have(false), unit0(false) { tag0 = m0 = v0 = F::c(0); }
    // multiplicity m (p - 1 = one lookup); `unit`: m is the constant p - 1 (two such lookups share the cheaper
    // bus.two_lookups: same helper constraint, no products with the multiplicities)
    TMX_HD void push(F tag, F m, F v, bool unit = false) {
        if (!have) {
            have = true;
            unit0 = unit;
            tag0 = tag;
            m0 = m;
            v0 = v;
            return;
        }
        have = false;
        const F a = v0, b = v;
        if (unit0 && unit) bus.two_lookups(tag0, a, tag, b);
        else bus.two(tag0, m0, 1, [&](int) { return a; }, tag, m, 1, [&](int) { return b; });
    }
    TMX_HD void push(int tag, F v) { push(F::c((uint64_t)tag), F::c(0xFFFFFFFF00000000ULL), v, true); }
    TMX_HD void flush() {
        if (!have) return;
        have = false;
        const F a = v0;
        bus.one(tag0, m0, 1, [&](int) { return a; });
    }
};
// tag constants as field values
template <class F>
TMX_HD F bus_tag(int t) { return F::c((uint64_t)t); }

constexpr int AIR_SHA256 = TMX_T_SHA256, AIR_SHA512 = TMX_T_SHA512, AIR_ED25519 = TMX_T_ED, AIR_LOGIC = TMX_T_LOGIC,
              AIR_RANGE = TMX_T_RANGE;
// Circuit shape the constant columns and the shape-dependent constraints depend on.
struct AirShape {
    uint32_t kind, n_max;
    uint32_t chain_len;
    char chain[52];
};
inline AirShape air_shape(uint32_t kind, uint32_t n_max, const char* chain, size_t chain_len) {
    AirShape sh;
    sh.kind = kind;
    sh.n_max = n_max;
    sh.chain_len = (uint32_t)(chain_len > 50 ? 50 : chain_len);
    for (int i = 0; i < 52; i++) sh.chain[i] = i < (int)sh.chain_len ? chain[i] : 0;
    return sh;
}
constexpr int AIR_MAX_PERIODIC = 16;

// Does 64-row chunk c of the SHA-256 table continue the message of chunk c - 1?  Layout (witness_jobs.cuh): per
// validator set n_max one-chunk leaf hashes, then np - 1 two-chunk inner nodes; then the header proofs, each a leaf (two
// chunks for the 72-byte last-block-id leaf of the step circuit, else one) and four two-chunk inner nodes; then one-chunk
// padding messages.
TMX_HD bool sha256_chunk_continues(AirShape sh, size_t c) {
    size_t np = 1;
    while (np < sh.n_max) np *= 2;
    const size_t set_chunks = sh.n_max + 2 * (np - 1), sets = sh.kind == 1 /* TMX_KIND_SKIP */ ? 2 : 1;
    if (c < sets * set_chunks) {
        const size_t local = c % set_chunks;
        return local >= sh.n_max && ((local - sh.n_max) & 1);
    }
    size_t h = c - sets * set_chunks;
    const int n_proofs = sh.kind == 1 ? 4 : 5;
    for (int k = 0; k < n_proofs; k++) {
        const size_t leaf = (sh.kind == 0 && k == 3) ? 2 : 1, len = leaf + 8;
        if (h < len) return h < leaf ? h == 1 : ((h - leaf) & 1) != 0;
        h -= len;
    }
    return false;
}

// ------------------------------------------------------------------------------------------ SHA-256
// per = {K_t, is_last_round, not_last_round, schedule_active (rounds 15..62), first_round} with period 64; k = the
// table's constant columns (S256K_*: message boundaries, chunk index, bus multiplicities -- fixed by the circuit shape)
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_sha256(const Row& l, const Row& n, const KRow& k, const Per& per, Emit& emit, Bus& bus) {
    const F K = per[0], LAST = per[1], NOTLAST = per[2], SCHED = per[3], FIRSTROW = per[4];
    const F two32 = F::c(1ULL << 32);
    for (int i = S256_A; i < S256_D; i++) emit(is_bool<F>(l[i]));
    for (int i = S256_AN; i < S256_W; i++) emit(is_bool<F>(l[i]));
    for (int i = S256_WB14; i < S256_CV; i++) emit(is_bool<F>(l[i]));
    for (int i = S256_CA; i < S256_DG; i++) emit(is_bool<F>(l[i]));
    for (int i = 0; i < 8; i++) emit(is_bool<F>(l[S256_DC + i]));
    // round function: Sigma1(e), Ch(e,f,g), Sigma0(a), Maj(a,b,c) as degree-3 expressions of the bit columns
    F S1 = F::c(0), CH = F::c(0), S0 = F::c(0), MJ = F::c(0);
    for (int i = 31; i >= 0; i--) {
        const F ei = l[S256_E + i], fi = l[S256_F + i], gi = l[S256_G + i];
        const F ai = l[S256_A + i], bi = l[S256_B + i], ci = l[S256_C + i];
        const F s1 = xor3<F>(l[S256_E + ((i + 6) & 31)], l[S256_E + ((i + 11) & 31)], l[S256_E + ((i + 25) & 31)]);
        const F ch = ei * fi + (F::c(1) - ei) * gi;
        const F s0 = xor3<F>(l[S256_A + ((i + 2) & 31)], l[S256_A + ((i + 13) & 31)], l[S256_A + ((i + 22) & 31)]);
        const F ab = ai * bi;
        const F mj = (ab + ai * ci + bi * ci) - F::c(2) * (ab * ci);
        S1 = S1 + S1 + s1;
        CH = CH + CH + ch;
        S0 = S0 + S0 + s0;
        MJ = MJ + MJ + mj;
    }
    const F T1 = l[S256_H] + S1 + (CH + K) + l[S256_W + 15];
    const F T2 = S0 + MJ;
    const F an = pack_bits<F>(l, S256_AN, 32), en = pack_bits<F>(l, S256_EN, 32);
    const F ca = pack_bits<F>(l, S256_CA, 3), ce = pack_bits<F>(l, S256_CE, 3), cw = pack_bits<F>(l, S256_CW, 2);
    emit((an + two32 * ca) - (T1 + T2));
    emit((en + two32 * ce) - (l[S256_D] + T1));
    for (int i = 0; i < 32; i++) {
        emit(NOTLAST * (n[S256_A + i] - l[S256_AN + i]));
        emit(NOTLAST * (n[S256_B + i] - l[S256_A + i]));
        emit(NOTLAST * (n[S256_C + i] - l[S256_B + i]));
        emit(NOTLAST * (n[S256_E + i] - l[S256_EN + i]));
        emit(NOTLAST * (n[S256_F + i] - l[S256_E + i]));
        emit(NOTLAST * (n[S256_G + i] - l[S256_F + i]));
    }
    const F pc = pack_bits<F>(l, S256_C, 32), pg = pack_bits<F>(l, S256_G, 32);
    emit(NOTLAST * (n[S256_D] - pc));
    emit(NOTLAST * (n[S256_H] - pg));
    for (int j = 0; j < 15; j++) emit(NOTLAST * (n[S256_W + j] - l[S256_W + j + 1]));
    for (int j = 0; j < 8; j++) emit(NOTLAST * (n[S256_CV + j] - l[S256_CV + j]));
    // message schedule
    emit(pack_bits<F>(l, S256_WB14, 32) - l[S256_W + 14]);
    emit(pack_bits<F>(l, S256_WB1, 32) - l[S256_W + 1]);
    F s1 = F::c(0), s0 = F::c(0);
    for (int i = 31; i >= 0; i--) {
        const F hi1 = i + 10 < 32 ? l[S256_WB14 + i + 10] : F::c(0);
        const F hi0 = i + 3 < 32 ? l[S256_WB1 + i + 3] : F::c(0);
        s1 = s1 + s1 + xor3<F>(l[S256_WB14 + ((i + 17) & 31)], l[S256_WB14 + ((i + 19) & 31)], hi1);
        s0 = s0 + s0 + xor3<F>(l[S256_WB1 + ((i + 7) & 31)], l[S256_WB1 + ((i + 18) & 31)], hi0);
    }
    emit((l[S256_WS] + two32 * cw) - ((s1 + l[S256_W + 9]) + (s0 + l[S256_W])));
    emit(SCHED * (n[S256_W + 15] - l[S256_WS]));
    // digest words on the last round, zero elsewhere
    const F fin[8] = {an, pack_bits<F>(l, S256_A, 32), pack_bits<F>(l, S256_B, 32), pc,
                      en, pack_bits<F>(l, S256_E, 32), pack_bits<F>(l, S256_F, 32), pg};
    for (int j = 0; j < 8; j++) {
        emit(LAST * ((l[S256_DG + j] + two32 * l[S256_DC + j]) - (l[S256_CV + j] + fin[j])));
        emit(NOTLAST * l[S256_DG + j]);
        emit(NOTLAST * l[S256_DC + j]);
    }
    // message chaining: a message starts from the IV, a continuation chunk from the previous chunk's digest
    const uint32_t IV[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (int j = 0; j < 8; j++) emit(k[S256K_FIRST] * (l[S256_CV + j] - F::c(IV[j])));
    for (int j = 0; j < 8; j++) emit(k[S256K_LINK] * (n[S256_CV + j] - l[S256_DG + j]));
    // the compression starts from the chunk's chaining value: working variables of round 0 = CV
    {
        const F st[8] = {fin[1], fin[2], fin[3], l[S256_D], fin[5], fin[6], fin[7], l[S256_H]};
        for (int j = 0; j < 8; j++) emit(FIRSTROW * (st[j] - l[S256_CV + j]));
    }
    // bus: the chunk's 16 message words (all in the schedule window on row 15) are received, its digest is sent
    const F minus_msg = F::c(0) - k[S256K_MSG];
    bus.two(
        bus_tag<F>(BUS_MSG256), minus_msg, 17, [&](int i) { return i == 0 ? k[S256K_CID] : l[S256_W + i - 1]; },
        bus_tag<F>(BUS_DIG256), k[S256K_DIG], 9, [&](int i) { return i == 0 ? k[S256K_CID] : l[S256_DG + i - 1]; });
}

// ------------------------------------------------------------------------------------------ SHA-512
// per = {K_t low half, K_t high half, is_round_79, not_round_79, not_chunk_end, schedule_active (rows 15..126),
// first_row_of_the_validator_slot, before_round_79, rows_79_to_126 (digest carried), last_row_of_the_first_chunk,
// not_last_row_of_the_slot}, period 256 (two 128-row chunks).  The digest columns are zero before round 79 and constant
// from row 79 to the end of the chunk, so the slot's second chunk can chain from them: its chaining value is the first
// chunk's digest when the slot's TWO flag is set (two-block message), else the IV (unused second compression).
// 64-bit words are (lo, hi) pairs of 32-bit field elements with an explicit carry from lo to hi.  Rows 80..127 of a chunk
// continue the round function with round constant 0 (include/tmx_trace.h), so only what reads the next row, the digest
// and the schedule hand-over need a selector.
// per[11] = first row of a chunk (the working variables of round 0 equal the chaining value); k = constant columns S512K_*
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_sha512(const Row& l, const Row& n, const KRow& kc, const Per& per, Emit& emit, Bus& bus) {
    const F KLO = per[0], KHI = per[1], LAST = per[2], NOTLAST = per[3], NOTEND = per[4], SCHED = per[5];
    const F two32 = F::c(1ULL << 32);
    for (int i = S512_A; i < S512_D; i++) emit(is_bool<F>(l[i]));
    for (int i = S512_AN; i < S512_W; i++) emit(is_bool<F>(l[i]));
    for (int i = S512_WB14; i < S512_CV; i++) emit(is_bool<F>(l[i]));
    for (int i = S512_CA; i < S512_DG; i++) emit(is_bool<F>(l[i]));
    for (int i = 0; i < 16; i++) emit(is_bool<F>(l[S512_DC + i]));
    // round function
    F S1[2] = {F::c(0), F::c(0)}, CH[2] = {F::c(0), F::c(0)}, S0[2] = {F::c(0), F::c(0)}, MJ[2] = {F::c(0), F::c(0)};
    for (int i = 63; i >= 0; i--) {
        const int h = i >> 5;
        const F ei = l[S512_E + i], fi = l[S512_F + i], gi = l[S512_G + i];
        const F ai = l[S512_A + i], bi = l[S512_B + i], ci = l[S512_C + i];
        const F s1 = xor3<F>(l[S512_E + ((i + 14) & 63)], l[S512_E + ((i + 18) & 63)], l[S512_E + ((i + 41) & 63)]);
        const F ch = ei * fi + (F::c(1) - ei) * gi;
        const F s0 = xor3<F>(l[S512_A + ((i + 28) & 63)], l[S512_A + ((i + 34) & 63)], l[S512_A + ((i + 39) & 63)]);
        const F ab = ai * bi;
        const F mj = (ab + ai * ci + bi * ci) - F::c(2) * (ab * ci);
        S1[h] = S1[h] + S1[h] + s1;
        CH[h] = CH[h] + CH[h] + ch;
        S0[h] = S0[h] + S0[h] + s0;
        MJ[h] = MJ[h] + MJ[h] + mj;
    }
    const F T1lo = l[S512_H] + S1[0] + (CH[0] + KLO) + l[S512_W + 30];
    const F T1hi = l[S512_H + 1] + S1[1] + (CH[1] + KHI) + l[S512_W + 31];
    const F an[2] = {pack_bits<F>(l, S512_AN, 32), pack_bits<F>(l, S512_AN + 32, 32)};
    const F en[2] = {pack_bits<F>(l, S512_EN, 32), pack_bits<F>(l, S512_EN + 32, 32)};
    const F ca[2] = {pack_bits<F>(l, S512_CA, 3), pack_bits<F>(l, S512_CA + 3, 3)};
    const F ce[2] = {pack_bits<F>(l, S512_CE, 3), pack_bits<F>(l, S512_CE + 3, 3)};
    const F cw[2] = {pack_bits<F>(l, S512_CW, 2), pack_bits<F>(l, S512_CW + 2, 2)};
    emit((an[0] + two32 * ca[0]) - (T1lo + (S0[0] + MJ[0])));
    emit((an[1] + two32 * ca[1]) - ((T1hi + (S0[1] + MJ[1])) + ca[0]));
    emit((en[0] + two32 * ce[0]) - (l[S512_D] + T1lo));
    emit((en[1] + two32 * ce[1]) - ((l[S512_D + 1] + T1hi) + ce[0]));
    // shift registers (not across the chunk boundary)
    for (int i = 0; i < 64; i++) {
        emit(NOTEND * (n[S512_A + i] - l[S512_AN + i]));
        emit(NOTEND * (n[S512_B + i] - l[S512_A + i]));
        emit(NOTEND * (n[S512_C + i] - l[S512_B + i]));
        emit(NOTEND * (n[S512_E + i] - l[S512_EN + i]));
        emit(NOTEND * (n[S512_F + i] - l[S512_E + i]));
        emit(NOTEND * (n[S512_G + i] - l[S512_F + i]));
    }
    const F pc[2] = {pack_bits<F>(l, S512_C, 32), pack_bits<F>(l, S512_C + 32, 32)};
    const F pg[2] = {pack_bits<F>(l, S512_G, 32), pack_bits<F>(l, S512_G + 32, 32)};
    for (int h = 0; h < 2; h++) {
        emit(NOTEND * (n[S512_D + h] - pc[h]));
        emit(NOTEND * (n[S512_H + h] - pg[h]));
    }
    for (int j = 0; j < 30; j++) emit(NOTEND * (n[S512_W + j] - l[S512_W + j + 2]));
    for (int j = 0; j < 16; j++) emit(NOTEND * (n[S512_CV + j] - l[S512_CV + j]));
    // message schedule
    emit(pack_bits<F>(l, S512_WB14, 32) - l[S512_W + 28]);
    emit(pack_bits<F>(l, S512_WB14 + 32, 32) - l[S512_W + 29]);
    emit(pack_bits<F>(l, S512_WB1, 32) - l[S512_W + 2]);
    emit(pack_bits<F>(l, S512_WB1 + 32, 32) - l[S512_W + 3]);
    F s1[2] = {F::c(0), F::c(0)}, s0[2] = {F::c(0), F::c(0)};
    for (int i = 63; i >= 0; i--) {
        const int h = i >> 5;
        const F hi1 = i + 6 < 64 ? l[S512_WB14 + i + 6] : F::c(0);
        const F hi0 = i + 7 < 64 ? l[S512_WB1 + i + 7] : F::c(0);
        s1[h] = s1[h] + s1[h] + xor3<F>(l[S512_WB14 + ((i + 19) & 63)], l[S512_WB14 + ((i + 61) & 63)], hi1);
        s0[h] = s0[h] + s0[h] + xor3<F>(l[S512_WB1 + ((i + 1) & 63)], l[S512_WB1 + ((i + 8) & 63)], hi0);
    }
    emit((l[S512_WS] + two32 * cw[0]) - ((s1[0] + l[S512_W + 18]) + (s0[0] + l[S512_W])));
    emit((l[S512_WS + 1] + two32 * cw[1]) - (((s1[1] + l[S512_W + 19]) + (s0[1] + l[S512_W + 1])) + cw[0]));
    emit(SCHED * (n[S512_W + 30] - l[S512_WS]));
    emit(SCHED * (n[S512_W + 31] - l[S512_WS + 1]));
    // digest words on round 79, zero elsewhere
    const F fin[8][2] = {{an[0], an[1]},
                         {pack_bits<F>(l, S512_A, 32), pack_bits<F>(l, S512_A + 32, 32)},
                         {pack_bits<F>(l, S512_B, 32), pack_bits<F>(l, S512_B + 32, 32)},
                         {pc[0], pc[1]},
                         {en[0], en[1]},
                         {pack_bits<F>(l, S512_E, 32), pack_bits<F>(l, S512_E + 32, 32)},
                         {pack_bits<F>(l, S512_F, 32), pack_bits<F>(l, S512_F + 32, 32)},
                         {pg[0], pg[1]}};
    for (int j = 0; j < 8; j++) {
        const F lo = l[S512_DG + 2 * j] + two32 * l[S512_DC + 2 * j];
        const F hi = l[S512_DG + 2 * j + 1] + two32 * l[S512_DC + 2 * j + 1];
        emit(LAST * (lo - (l[S512_CV + 2 * j] + fin[j][0])));
        emit(LAST * (hi - ((l[S512_CV + 2 * j + 1] + fin[j][1]) + l[S512_DC + 2 * j])));
        for (int k = 0; k < 2; k++) {
            emit(per[7] * l[S512_DG + 2 * j + k]);                                  // zero before round 79
            emit(per[8] * (n[S512_DG + 2 * j + k] - l[S512_DG + 2 * j + k]));       // carried to the end of the chunk
            emit(NOTLAST * l[S512_DC + 2 * j + k]);
        }
    }
    // chaining inside a validator slot: first chunk from the IV, second chunk from the first one's digest iff TWO
    const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                            0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    const F two = l[S512_TWO];
    emit(is_bool<F>(two));
    emit(per[10] * (n[S512_TWO] - two));
    for (int j = 0; j < 8; j++)
        for (int k = 0; k < 2; k++) {
            const F iv = F::c(k ? IV[j] >> 32 : (uint64_t)(uint32_t)IV[j]);
            emit(per[6] * (l[S512_CV + 2 * j + k] - iv));
            emit(per[9] * (n[S512_CV + 2 * j + k] - (iv + two * (l[S512_DG + 2 * j + k] - iv))));
        }
    // the compression starts from the chunk's chaining value
    {
        const F st[8][2] = {{fin[1][0], fin[1][1]}, {fin[2][0], fin[2][1]}, {fin[3][0], fin[3][1]}, {l[S512_D], l[S512_D + 1]},
                            {fin[5][0], fin[5][1]}, {fin[6][0], fin[6][1]}, {fin[7][0], fin[7][1]}, {l[S512_H], l[S512_H + 1]}};
        for (int j = 0; j < 8; j++)
            for (int h = 0; h < 2; h++) emit(per[11] * (st[j][h] - l[S512_CV + 2 * j + h]));
    }
    // bus: each chunk of an active slot receives its 16 message words (row 15); the digest of the slot's message leaves
    // from row 79 of the second chunk when the message has two blocks, else of the first
    const F minus_msg = F::c(0) - kc[S512K_MSG];
    const F dig_mult = kc[S512K_DIG0] * (F::c(1) - two) + kc[S512K_DIG1] * two;
    bus.two(
        bus_tag<F>(BUS_MSG512), minus_msg, 35,
        [&](int i) { return i == 0 ? kc[S512K_VID] : (i == 1 ? kc[S512K_CHUNK] : (i == 2 ? two : l[S512_W + i - 3])); },
        bus_tag<F>(BUS_DIG512), dig_mult, 17, [&](int i) { return i == 0 ? kc[S512K_VID] : l[S512_DG + i - 1]; });
}

// ------------------------------------------------------------------------------------------ Ed25519
TMX_HD uint64_t p25519_limb(int i) { return i == 0 ? 0xFFEDULL : (i == 15 ? 0x7FFFULL : 0xFFFFULL); }

// U * V = c + q * p with carries: the 32 limb equations of one multiplication gadget (cells from column g0), in 16
// pairs:  e_2K + 2^16 e_2K+1 + w_{K-1} = 2^32 w_K,  w_K = wlo_K + 2^16 whi_K - ED_W_OFFSET,  w_{-1} = w_15 = 0.
// Operand limbs are below 2^19 in magnitude and every committed cell is range checked on the bus (c, q, wlo < 2^16,
// whi < 2^11), so every term stays below 2^60: the equations hold over the integers, not just in F_p.
template <class F, class Row, class Emit>
TMX_HD void ed_mul_gadget(const F U[16], const F V[16], const Row& l, int g0, Emit& emit) {
    const F off = F::c(ED_W_OFFSET), two16 = F::c(1 << 16), two32 = F::c(1ULL << 32);
    F q[17];
    for (int i = 0; i < 17; i++) q[i] = l[g0 + ED_MUL_Q + i];
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
            if (k < 16) e = e - l[g0 + k];
            for (int i = 0; i < 17; i++) {
                const int j = k - i;
                if (j >= 0 && j < 16) e = e - q[i] * F::c(p25519_limb(j));
            }
            s = h ? s + two16 * e : e;
        }
        if (K >= 1) s = s + wprev;
        if (K < ED_MUL_NW) {
            wprev = (l[g0 + ED_MUL_WLO + K] + two16 * l[g0 + ED_MUL_WHI + K]) - off;
            s = s - two32 * wprev;
        }
        emit(s);
    }
}

// per = {not_block_end (row % 256 != 255), first row of a slot (row % 256 == 0), not last row of a 16-row group, first row
// of a 16-row group}, period 256; k = constant columns EDK_*
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_ed25519(const Row& l, const Row& n, const KRow& k, const Per& per, Emit& emit, Bus& bus) {
    const F NOTEND = per[0], FIRST = per[1], NOT16END = per[2], FIRST16 = per[3];
    const F bs = l[ED_BS], bh = l[ED_BH];
    emit(is_bool<F>(bs));
    emit(is_bool<F>(bh));
    // scalar limbs: inside a 16-row group acc' = 2 acc + bit, starting from 0
    emit(FIRST16 * l[ED_SACC_S]);
    emit(FIRST16 * l[ED_SACC_H]);
    const F limb_s = (l[ED_SACC_S] + l[ED_SACC_S]) + bs, limb_h = (l[ED_SACC_H] + l[ED_SACC_H]) + bh;
    emit(NOT16END * (n[ED_SACC_S] - limb_s));
    emit(NOT16END * (n[ED_SACC_H] - limb_h));
    const int X1 = ED_ACC, Y1 = ED_ACC + 16, Z1 = ED_ACC + 32;
    const int YPX = ED_ADD, YMX = ED_ADD + 16, T2D = ED_ADD + 32;
    auto G = [](int m) { return ED_MUL + m * ED_MUL_STRIDE; };
    F u[16], v[16], E[16], Fq[16], Gq[16], H[16];
    // ---- doubling (dbl-2008-hwcd, a = -1) ----
    for (int i = 0; i < 16; i++) u[i] = l[X1 + i];
    ed_mul_gadget<F>(u, u, l, G(ED_G_A), emit);
    for (int i = 0; i < 16; i++) u[i] = l[Y1 + i];
    ed_mul_gadget<F>(u, u, l, G(ED_G_B), emit);
    for (int i = 0; i < 16; i++) u[i] = l[Z1 + i];
    ed_mul_gadget<F>(u, u, l, G(ED_G_CZ), emit);
    for (int i = 0; i < 16; i++) u[i] = l[X1 + i] + l[Y1 + i];
    ed_mul_gadget<F>(u, u, l, G(ED_G_S), emit);
    for (int i = 0; i < 16; i++) {
        const F pl = F::c(p25519_limb(i)), p2 = pl + pl;
        const F A2 = l[G(ED_G_A) + i], B2 = l[G(ED_G_B) + i], Cz = l[G(ED_G_CZ) + i], S = l[G(ED_G_S) + i];
        const F ba = B2 - A2;
        E[i] = ((S - A2) - B2) + p2;
        Gq[i] = ba + pl;
        Fq[i] = (ba - (Cz + Cz)) + (p2 + pl);
        H[i] = (p2 - A2) - B2;
    }
    ed_mul_gadget<F>(E, Fq, l, G(ED_G_X3), emit);
    ed_mul_gadget<F>(Gq, H, l, G(ED_G_Y3), emit);
    ed_mul_gadget<F>(E, H, l, G(ED_G_T3), emit);
    ed_mul_gadget<F>(Fq, Gq, l, G(ED_G_Z3), emit);
    // ---- mixed addition of the row's addend (add-2008-hwcd-3 with Z2 = 1; the T output is never used) ----
    for (int i = 0; i < 16; i++) {
        u[i] = (l[G(ED_G_Y3) + i] - l[G(ED_G_X3) + i]) + F::c(p25519_limb(i));
        v[i] = l[YMX + i];
    }
    ed_mul_gadget<F>(u, v, l, G(ED_G_AA), emit);
    for (int i = 0; i < 16; i++) {
        u[i] = l[G(ED_G_Y3) + i] + l[G(ED_G_X3) + i];
        v[i] = l[YPX + i];
    }
    ed_mul_gadget<F>(u, v, l, G(ED_G_BB), emit);
    for (int i = 0; i < 16; i++) {
        u[i] = l[G(ED_G_T3) + i];
        v[i] = l[T2D + i];
    }
    ed_mul_gadget<F>(u, v, l, G(ED_G_CC), emit);
    for (int i = 0; i < 16; i++) {
        const F pl = F::c(p25519_limb(i));
        const F A = l[G(ED_G_AA) + i], B = l[G(ED_G_BB) + i], C = l[G(ED_G_CC) + i], Z3 = l[G(ED_G_Z3) + i];
        const F d2 = Z3 + Z3;
        E[i] = (B - A) + pl;
        Fq[i] = (d2 - C) + pl;
        Gq[i] = d2 + C;
        H[i] = B + A;
    }
    ed_mul_gadget<F>(E, Fq, l, G(ED_G_X4), emit);
    ed_mul_gadget<F>(Gq, H, l, G(ED_G_Y4), emit);
    ed_mul_gadget<F>(Fq, Gq, l, G(ED_G_Z4), emit);
    // ---- transitions inside a slot, accumulator = O = (0 : 1 : 1) on its first row ----
    const int out_slot[3] = {ED_G_X4, ED_G_Y4, ED_G_Z4};
    for (int co = 0; co < 3; co++)
        for (int i = 0; i < 16; i++) emit(NOTEND * (n[ED_ACC + 16 * co + i] - l[G(out_slot[co]) + i]));
    for (int co = 0; co < 3; co++)
        for (int i = 0; i < 16; i++) emit(FIRST * (l[ED_ACC + 16 * co + i] - F::c(co >= 1 && i == 0 ? 1 : 0)));
    // ---- bus: range checks of every gadget cell, then the addend lookup, the scalar limbs and the result ----
    LookupPairs<F, Bus> rc(bus);
    for (int m = 0; m < ED_N_MUL; m++) {
        for (int i = 0; i < ED_MUL_WHI; i++) rc.push(BUS_R16, l[G(m) + i]);
        for (int i = 0; i < ED_MUL_NW; i++) rc.push(BUS_R11, l[G(m) + ED_MUL_WHI + i]);
    }
    rc.flush();
    const F sel = bs + (bh + bh);
    bus.two(
        bus_tag<F>(BUS_ADDEND), F::c(0) - k[EDK_ACTIVE], 50, [&](int i) { return i == 0 ? k[EDK_VID] : (i == 1 ? sel : l[ED_ADD + i - 2]); },
        bus_tag<F>(BUS_SCALAR), k[EDK_SEND16], 4, [&](int i) { return i == 0 ? k[EDK_VID] : (i == 1 ? F::c(0) : (i == 2 ? k[EDK_LIMB] : limb_s)); });
    bus.two(
        bus_tag<F>(BUS_SCALAR), k[EDK_SEND16], 4, [&](int i) { return i == 0 ? k[EDK_VID] : (i == 1 ? F::c(1) : (i == 2 ? k[EDK_LIMB] : limb_h)); },
        bus_tag<F>(BUS_EDRES), k[EDK_LAST], 49, [&](int i) { return i == 0 ? k[EDK_VID] : l[G(out_slot[(i - 1) >> 4]) + ((i - 1) & 15)]; });
}

// ------------------------------------------------------------------------------------------ range table
// row t provides the value t (constant column RGK_T) to the 16-bit lookups with multiplicity M16, and on its first 2^11 /
// 2^8 / 2 rows to the 11-bit / 8-bit / 1-bit lookups
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_range(const Row& l, const Row&, const KRow& k, const Per&, Emit& emit, Bus& bus) {
    emit((F::c(1) - k[RGK_S11]) * l[RG_M11]);
    emit((F::c(1) - k[RGK_S8]) * l[RG_M8]);
    emit((F::c(1) - k[RGK_S1]) * l[RG_M1]);
    const F t = k[RGK_T];
    bus.two(bus_tag<F>(BUS_R16), l[RG_M16], 1, [&](int) { return t; }, bus_tag<F>(BUS_R11), l[RG_M11], 1, [&](int) { return t; });
    bus.two(bus_tag<F>(BUS_R8), l[RG_M8], 1, [&](int) { return t; }, bus_tag<F>(BUS_R1), l[RG_M1], 1, [&](int) { return t; });
}

// ------------------------------------------------------------------------------------------ dispatch and table shapes
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_eval(int table, const Row& l, const Row& n, const KRow& k, const Per& per, Emit& emit, Bus& bus) {
    if (table == AIR_SHA256) air_sha256<F>(l, n, k, per, emit, bus);
    else if (table == AIR_SHA512) air_sha512<F>(l, n, k, per, emit, bus);
    else if (table == AIR_ED25519) air_ed25519<F>(l, n, k, per, emit, bus);
    else if (table == AIR_RANGE) air_range<F>(l, n, k, per, emit, bus);
}

// helper columns of the second commitment round = number of bus.one / bus.two declarations of the table's AIR
constexpr int ED_HELPERS = ED_N_MUL * ED_MUL_STRIDE / 2 + 2;
TMX_HD int air_cols(int t) {
    return t == AIR_SHA256 ? S256_COLS : t == AIR_SHA512 ? S512_COLS : t == AIR_ED25519 ? ED_COLS : t == AIR_RANGE ? RG_COLS : 0;
}
TMX_HD int air_const_cols(int t) {
    return t == AIR_SHA256 ? S256K_COLS : t == AIR_SHA512 ? S512K_COLS : t == AIR_ED25519 ? EDK_COLS : t == AIR_RANGE ? RGK_COLS : 0;
}
TMX_HD int air_helpers(int t) { return t == AIR_SHA256 ? 1 : t == AIR_SHA512 ? 1 : t == AIR_ED25519 ? ED_HELPERS : t == AIR_RANGE ? 2 : 0; }
TMX_HD int air_aux_cols(int t) { return air_cols(t) ? 2 * (air_helpers(t) + 1) : 0; }  // helpers and the running sum, two components each
TMX_HD int air_n_periodic(int t) { return t == AIR_SHA256 ? 5 : t == AIR_SHA512 ? 12 : t == AIR_ED25519 ? 4 : 0; }
TMX_HD size_t air_period(int t) { return t == AIR_SHA256 ? 64 : t == AIR_SHA512 ? S512_ROWS_PER_VALIDATOR : t == AIR_ED25519 ? ED_ROWS_PER_VALIDATOR : 1; }

TMX_HD size_t air_pow2_at_least(size_t x) {
    size_t p = 1;
    while (p < x) p *= 2;
    return p;
}
TMX_HD size_t air_sha256_used_chunks(AirShape sh) {
    const size_t np = air_pow2_at_least(sh.n_max), set = (size_t)sh.n_max + 2 * (np - 1);
    return sh.kind == 1 ? 2 * set + 36 : set + 46;
}
// rows of table t for a circuit shape (0 = the table is absent)
TMX_HD size_t air_rows(int t, AirShape sh) {
    if (t == AIR_SHA256) return air_pow2_at_least(air_sha256_used_chunks(sh) * S256_ROUNDS);
    if (t == AIR_SHA512) return air_pow2_at_least((size_t)sh.n_max * S512_ROWS_PER_VALIDATOR);
    if (t == AIR_ED25519) return air_pow2_at_least((size_t)sh.n_max * ED_ROWS_PER_VALIDATOR);
    if (t == AIR_RANGE) return (size_t)1 << RG_LOG_ROWS;
    return 0;
}

// The cross-table messages (hash inputs / digests, scalar limbs, addends, results) have their counterparty in the logic
// table; until that table carries them their multiplicity columns are zero and only the range-check bus is live.
#ifndef TMX_BUS_LINKS
#define TMX_BUS_LINKS 1
#endif

// periodic pattern of column pc at row r of its period
TMX_HD uint64_t air_periodic_pattern(int table, int pc, size_t row, const uint32_t* k256_table, const uint64_t* k512_table) {
    if (table == AIR_SHA256) {
        const int r64 = (int)(row & 63);
        if (pc == 0) return k256_table[r64];
        if (pc == 1) return r64 == 63;
        if (pc == 2) return r64 != 63;
        if (pc == 3) return r64 >= 15 && r64 <= 62;
        return r64 == 0;
    }
    if (table == AIR_SHA512) {
        const int r = (int)(row & 255);
        const int rr = r % S512_ROWS_PER_CHUNK;  // row inside the chunk; r is the row inside the validator's two-chunk slot
        if (pc == 0) return rr < S512_ROUNDS ? (uint32_t)k512_table[rr] : 0;
        if (pc == 1) return rr < S512_ROUNDS ? k512_table[rr] >> 32 : 0;
        if (pc == 2) return rr == S512_ROUNDS - 1;
        if (pc == 3) return rr != S512_ROUNDS - 1;
        if (pc == 4) return rr != S512_ROWS_PER_CHUNK - 1;
        if (pc == 5) return rr >= 15 && rr <= S512_ROWS_PER_CHUNK - 2;
        if (pc == 6) return r == 0;
        if (pc == 7) return rr < S512_ROUNDS - 1;
        if (pc == 8) return rr >= S512_ROUNDS - 1 && rr <= S512_ROWS_PER_CHUNK - 2;
        if (pc == 9) return r == S512_ROWS_PER_CHUNK - 1;
        if (pc == 10) return r != S512_ROWS_PER_VALIDATOR - 1;
        return rr == 0;
    }
    const int r = (int)(row & 255);  // Ed25519
    if (pc == 0) return r != 255;
    if (pc == 1) return r == 0;
    if (pc == 2) return (r & 15) != 15;
    return (r & 15) == 0;
}

// value of constant column kc of table t at row `row` (fixed by the circuit shape; committed once per circuit, the cap of
// that commitment is part of the circuit digest)
TMX_HD uint64_t air_const_value(int table, int kc, size_t row, AirShape sh) {
    if (table == AIR_SHA256) {
        const int r64 = (int)(row & 63);
        const size_t chunk = row >> 6;
        const bool used = chunk < air_sha256_used_chunks(sh);
        if (kc == S256K_FIRST) return r64 == 0 && !sha256_chunk_continues(sh, chunk);
        if (kc == S256K_LINK) return r64 == 63 && sha256_chunk_continues(sh, chunk + 1);
        if (kc == S256K_CID) return chunk;
        if (kc == S256K_MSG) return TMX_BUS_LINKS && used && r64 == 15;
        return TMX_BUS_LINKS && used && r64 == 63 && !sha256_chunk_continues(sh, chunk + 1);
    }
    if (table == AIR_SHA512) {
        const int r = (int)(row & 255);
        const size_t slot = row >> 8;
        const bool active = slot < sh.n_max;
        if (kc == S512K_VID) return slot;
        if (kc == S512K_CHUNK) return r >> 7;
        if (kc == S512K_MSG) return TMX_BUS_LINKS && active && (r & 127) == 15;
        if (kc == S512K_DIG0) return TMX_BUS_LINKS && active && r == S512_ROUNDS - 1;
        return TMX_BUS_LINKS && active && r == S512_ROWS_PER_CHUNK + S512_ROUNDS - 1;
    }
    if (table == AIR_ED25519) {
        const int r = (int)(row & 255);
        const size_t slot = row >> 8;
        const bool active = TMX_BUS_LINKS && slot < sh.n_max;
        if (kc == EDK_VID) return slot;
        if (kc == EDK_ACTIVE) return active;
        if (kc == EDK_SEND16) return active && (r & 15) == 15;
        if (kc == EDK_LIMB) return 15 - (r >> 4);
        return active && r == 255;
    }
    if (table == AIR_RANGE) {
        if (kc == RGK_T) return row;
        if (kc == RGK_S11) return row < 2048;
        if (kc == RGK_S8) return row < 256;
        return row < 2;
    }
    return 0;
}

// Host NTT (in place, natural order in and out) for the public columns: iterative radix-2, O(n log n).
inline void air_host_ntt(std::vector<gl>& a, bool inverse) {
    const size_t n = a.size();
    const unsigned lg = ilog2(n);
    for (size_t i = 0; i < n; i++) {
        const size_t j = bitrev32((uint32_t)i, lg);
        if (i < j) std::swap(a[i], a[j]);
    }
    for (unsigned s = 1; s <= lg; s++) {
        const size_t half = (size_t)1 << (s - 1);
        gl w = gl_root_of_unity(s);
        if (inverse) w = gl_inv(w);
        std::vector<gl> tw(half);
        tw[0] = 1;
        for (size_t k = 1; k < half; k++) tw[k] = gl_mul(tw[k - 1], w);
        for (size_t base = 0; base < n; base += 2 * half)
            for (size_t k = 0; k < half; k++) {
                const gl u = a[base + k], v = gl_mul(a[base + k + half], tw[k]);
                a[base + k] = gl_add(u, v);
                a[base + k + half] = gl_sub(u, v);
            }
    }
    if (inverse) {
        const gl ninv = gl_inv((gl)n);
        for (auto& x : a) x = gl_mul(x, ninv);
    }
}

// coefficients of the interpolant of periodic column pc over one period (P values)
inline std::vector<gl> air_periodic_coeffs(int table, int pc, const uint32_t* k256_table, const uint64_t* k512_table) {
    const size_t P = air_period(table);
    std::vector<gl> c(P);
    for (size_t r = 0; r < P; r++) c[r] = (gl)air_periodic_pattern(table, pc, r, k256_table, k512_table);
    air_host_ntt(c, true);
    return c;
}

// Host: values of the periodic columns on the LDE coset, [nper][2P], indexed by (natural LDE index mod 2P).
// Column pc is the interpolant s of its one-period pattern composed with x -> x^(n/P); on the coset
// x_j = 7 w_m^j this only depends on j mod 2P: s(7^(n/P) w_2P^j).
inline std::vector<gl> air_periodic_lde_table(int table, unsigned log_n, const uint32_t* k256_table, const uint64_t* k512_table) {
    const size_t n = (size_t)1 << log_n, P = air_period(table);
    const int nper = air_n_periodic(table);
    std::vector<gl> tab((size_t)nper * 2 * P);
    const gl sh = gl_pow(GL_GEN, n / P);
    for (int pc = 0; pc < nper; pc++) {
        std::vector<gl> coef = air_periodic_coeffs(table, pc, k256_table, k512_table);
        coef.resize(2 * P, 0);
        gl s = 1;
        for (size_t k = 0; k < P; k++) {  // s(sh * x): scale coefficient k by sh^k, then a plain NTT of size 2P
            coef[k] = gl_mul(coef[k], s);
            s = gl_mul(s, sh);
        }
        air_host_ntt(coef, false);
        for (size_t j = 0; j < 2 * P; j++) tab[(size_t)pc * 2 * P + j] = coef[j];
    }
    return tab;
}

}  // namespace tmx
