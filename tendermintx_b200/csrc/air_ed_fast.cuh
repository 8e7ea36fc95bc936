// Factored evaluation of the Ed25519 table's random-linear-combination of constraints (K5 fast path).
//
// air_ed25519() in air.cuh is the definition: 1 + 17 * 16 + 128 + 208 constraints folded by Horner's rule in the
// challenge alpha.  Evaluated literally that costs ~9,000 field multiplications per LDE point, almost all of them
// the limb products U_i V_j and q_i p_j of the 17 multiplication gadgets.  The 16 paired limb equations of a gadget
// (pair K = limb equations 2K and 2K + 1, the second weighted by 2^16) carry CONSECUTIVE powers of alpha, so with
// the half-limb tildes  Xe = sum_i X_2i alpha^(7-i),  Xo = sum_i X_2i+1 alpha^(7-i),  X^ = Xe + 2^16 Xo:
//
//   sum_K alpha^(15-K) s_K = alpha U^ V^ - alpha^8 C^ - Q^' P^ + (1 - 2^32 alpha) (Uo Vo - Qo Po + W~)
//
//   Q^' = sum_{i<=8} q_2i alpha^(8-i) + 2^16 alpha Qo  (q has 17 limbs),  W~ = sum_k (w'_k - ED_W_OFFSET) alpha^(14-k).
//
// Operands that are limb-wise linear combinations of earlier cells (E = B - A + p, ...) have (X^, Xo) that are the same
// linear combinations, so every trace cell enters one Horner chain per challenge: ~1,300 multiply-adds per
// (point, challenge).  The result is the SAME field element as the literal evaluation (polynomial identity over
// F_p), so proof bytes do not change; tests pin it against air_ed25519() on the host (tmx_host_air_ed25519) and
// against the oracle's proofs on the GPU.
#pragma once
#include "air.cuh"

namespace tmx {

struct EdFastConsts {
    gl a, a8, a15, a16, a32, a48, a64, a128;
    gl base_tilde[4];  // full tildes of the base point's X, Y, Z (= 1), T limbs (block initialisation constraints)
    gl p_hat, p_odd;        // p = 2^255 - 19 in 16-bit limbs
    gl twod_hat, twod_odd;  // the curve constant 2d
    gl w_off;               // ED_W_OFFSET * sum_{e<=14} a^e
    gl one_m_232a;          // 1 - 2^32 a
};

// half-limb tildes of 16 constant limbs: e = sum_i x_2i a^(7-i), o = sum_i x_2i+1 a^(7-i)
inline void ed_half_tildes(const uint64_t x[16], gl a, gl* hat, gl* odd) {
    gl e = 0, o = 0;
    for (int i = 0; i < 8; i++) {
        e = gl_add(gl_mul(e, a), (gl)x[2 * i]);
        o = gl_add(gl_mul(o, a), (gl)x[2 * i + 1]);
    }
    *hat = gl_add(e, gl_mul((gl)1 << 16, o));
    *odd = o;
}

inline EdFastConsts ed_fast_consts(gl a) {
    static const uint64_t TWOD[16] = {0xF159, 0x26B2, 0x9B94, 0xEBD6, 0xB156, 0x8283, 0x149A, 0x00E0,
                                      0xD130, 0xEEF3, 0x80F2, 0x198E, 0xFCE7, 0x56DF, 0xD9DC, 0x2406};
    static const uint64_t BXL[16] = ED_BASE_X_LIMBS, BYL[16] = ED_BASE_Y_LIMBS, BTL[16] = ED_BASE_T_LIMBS;
    EdFastConsts k;
    k.a = a;
    k.a8 = gl_pow(a, 8);
    k.a15 = gl_pow(a, 15);
    k.a16 = gl_pow(a, 16);
    k.a32 = gl_pow(a, 32);
    k.a48 = gl_pow(a, 48);
    k.a64 = gl_pow(a, 64);
    k.a128 = gl_pow(a, 128);
    {
        gl bx = 0, by = 0, bt = 0;
        for (int i = 0; i < 16; i++) {
            bx = gl_add(gl_mul(bx, a), (gl)BXL[i]);
            by = gl_add(gl_mul(by, a), (gl)BYL[i]);
            bt = gl_add(gl_mul(bt, a), (gl)BTL[i]);
        }
        k.base_tilde[0] = bx;
        k.base_tilde[1] = by;
        k.base_tilde[2] = k.a15;  // limbs (1, 0, ..., 0)
        k.base_tilde[3] = bt;
    }
    uint64_t pl[16];
    for (int i = 0; i < 16; i++) pl[i] = p25519_limb(i);
    ed_half_tildes(pl, a, &k.p_hat, &k.p_odd);
    ed_half_tildes(TWOD, a, &k.twod_hat, &k.twod_odd);
    gl sum = 0;
    for (int e = 0; e <= 14; e++) sum = gl_add(gl_mul(sum, a), 1);
    k.w_off = gl_mul((gl)ED_W_OFFSET, sum);
    k.one_m_232a = gl_sub(1, gl_mul((gl)1 << 32, a));
    return k;
}

// A gadget operand as the fold sees it: hat = Xe + 2^16 Xo and odd = Xo, with the half-limb tildes
// Xe = sum_i X_2i a^(7-i), Xo = sum_i X_2i+1 a^(7-i).  Linear combinations of operands act component-wise.
struct EdOp {
    gl hat, odd;
};
TMX_HD EdOp ed_op_add(EdOp x, EdOp y) { return EdOp{gl_add(x.hat, y.hat), gl_add(x.odd, y.odd)}; }
TMX_HD EdOp ed_op_sub(EdOp x, EdOp y) { return EdOp{gl_sub(x.hat, y.hat), gl_sub(x.odd, y.odd)}; }

// Row: operator[](int col) -> FB (canonical cell).  per = {not_block_end, first row of [s]B, first row of [h]A}.
// Returns the Horner-folded constraint value for challenge k.a.
template <class Row>
TMX_HD gl ed25519_constraints_fast(const Row& l, const Row& n, const gl per[3], const EdFastConsts& k) {
    const gl notend = per[0], s0 = per[1], h0 = per[2];
    const gl a = k.a;
    const gl two16 = (gl)1 << 16;
    auto G = [](int m) { return ED_MUL + m * ED_MUL_STRIDE; };
    const gl bit = l[ED_BIT].v;
    gl acc = gl_mul(bit, gl_sub(bit, 1));
    // half-limb tildes of the 16 cells starting at col0
    auto halves = [&](int col0) {
        gl e = l[col0].v, o = l[col0 + 1].v;
#pragma unroll
        for (int i = 1; i < 8; i++) {
            e = gl_mac_nc(e, a, l[col0 + 2 * i].v);
            o = gl_mac_nc(o, a, l[col0 + 2 * i + 1].v);
        }
        o = gl_canon(o);
        return EdOp{gl_canon(gl_mac_nc(o, two16, e)), o};
    };
    // ---- pass 1: accumulator / running-double coordinates (full tildes for the block initialisation, operand form for
    //      the gadgets), the gadget outputs they select from, transitions ----
    const int sum_slot[4] = {5, 6, 8, 7}, dbl_slot[4] = {13, 14, 16, 15};  // X, Y, Z, T
    gl R[4], S[4], T = 0;
    EdOp Ro[4], So[4];
    gl Cs[4], Cd[4];  // hats of the product limbs of the sum / double gadgets
#pragma unroll
    for (int co = 0; co < 4; co++) {
        gl r_t = 0, s_t = 0;
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const gl r = l[ED_RES + 16 * co + i].v, t = l[ED_TMP + 16 * co + i].v;
            const gl s = l[G(sum_slot[co]) + i].v, d = l[G(dbl_slot[co]) + i].v;
            const gl nr = n[ED_RES + 16 * co + i].v, nt = n[ED_TMP + 16 * co + i].v;
            r_t = gl_mac_nc(r_t, a, r);
            s_t = gl_mac_nc(s_t, a, t);
            T = gl_mac_nc(T, a, gl_sub(nr, gl_add(r, gl_mul(bit, gl_sub(s, r)))));
            T = gl_mac_nc(T, a, gl_sub(nt, d));
        }
        R[co] = gl_canon(r_t);
        S[co] = gl_canon(s_t);
        Ro[co] = halves(ED_RES + 16 * co);
        So[co] = halves(ED_TMP + 16 * co);
        Cs[co] = halves(G(sum_slot[co])).hat;
        Cd[co] = halves(G(dbl_slot[co])).hat;
    }
    // ---- pass 2: the 17 gadgets in emission order; 16 paired limb equations each:
    //   sum_K a^(15-K) s_K = a U^ V^ - a^8 C^ - Q^' P^ + (1 - 2^32 a) (Uo Vo - Qo Po + W~),
    //   Q^' = sum_{i<=8} q_2i a^(8-i) + 2^16 a Qo,  W~ = sum_{k<=14} (w'_k - ED_W_OFFSET) a^(14-k)
    auto tilde = [&](int col0, int len) {
        gl t = l[col0].v;
#pragma unroll 4
        for (int i = 1; i < len; i++) t = gl_mac_nc(t, a, l[col0 + i].v);
        return gl_canon(t);
    };
    const EdOp Pp{k.p_hat, k.p_odd};
    // have_c: the hat of the product limbs is already known from pass 1; returns the product as an operand
    auto gadget = [&](EdOp u, EdOp v, int g, bool have_c, gl c_hat) {
        EdOp c{c_hat, 0};
        if (!have_c) c = halves(G(g));
        gl qe = l[G(g) + ED_MUL_Q].v, qo = l[G(g) + ED_MUL_Q + 1].v;
#pragma unroll
        for (int i = 1; i < 8; i++) {
            qe = gl_mac_nc(qe, a, l[G(g) + ED_MUL_Q + 2 * i].v);
            qo = gl_mac_nc(qo, a, l[G(g) + ED_MUL_Q + 2 * i + 1].v);
        }
        qe = gl_mac_nc(qe, a, l[G(g) + ED_MUL_Q + 16].v);
        qo = gl_canon(qo);
        const gl q_hat = gl_add(gl_canon(qe), gl_mul(two16, gl_mul(a, qo)));
        const gl w_t = gl_sub(tilde(G(g) + ED_MUL_W, ED_MUL_NW), k.w_off);
        gl s = gl_mul(a, gl_mul(u.hat, v.hat));
        s = gl_sub(s, gl_mul(k.a8, c.hat));
        s = gl_sub(s, gl_mul(q_hat, Pp.hat));
        const gl inner = gl_add(gl_sub(gl_mul(u.odd, v.odd), gl_mul(qo, Pp.odd)), w_t);
        s = gl_add(s, gl_mul(k.one_m_232a, inner));
        acc = gl_add(gl_mul(acc, k.a16), s);
        return c;
    };
    const EdOp P2 = ed_op_add(Pp, Pp), P3 = ed_op_add(P2, Pp);
    const EdOp X1 = Ro[0], Y1 = Ro[1], Z1 = Ro[2], T1 = Ro[3], X2 = So[0], Y2 = So[1], Z2 = So[2], T2 = So[3];
    const EdOp A = gadget(ed_op_add(ed_op_sub(Y1, X1), Pp), ed_op_add(ed_op_sub(Y2, X2), Pp), 0, false, 0);
    const EdOp B = gadget(ed_op_add(Y1, X1), ed_op_add(Y2, X2), 1, false, 0);
    const EdOp TT = gadget(T1, T2, 2, false, 0);
    const EdOp C = gadget(TT, EdOp{k.twod_hat, k.twod_odd}, 3, false, 0);
    const EdOp Dh = gadget(Z1, Z2, 4, false, 0);
    {
        const EdOp d2 = ed_op_add(Dh, Dh);
        const EdOp E = ed_op_add(ed_op_sub(B, A), Pp), Fq = ed_op_add(ed_op_sub(d2, C), Pp), Gq = ed_op_add(d2, C), H = ed_op_add(B, A);
        gadget(E, Fq, 5, true, Cs[0]);
        gadget(Gq, H, 6, true, Cs[1]);
        gadget(E, H, 7, true, Cs[3]);
        gadget(Fq, Gq, 8, true, Cs[2]);
    }
    const EdOp A2 = gadget(X2, X2, 9, false, 0);
    const EdOp B2 = gadget(Y2, Y2, 10, false, 0);
    const EdOp Cz = gadget(Z2, Z2, 11, false, 0);
    const EdOp xy = ed_op_add(X2, Y2);
    const EdOp Sq = gadget(xy, xy, 12, false, 0);
    {
        const EdOp ba = ed_op_sub(B2, A2);
        const EdOp E = ed_op_add(ed_op_sub(ed_op_sub(Sq, A2), B2), P2), Gq = ed_op_add(ba, Pp);
        const EdOp Fq = ed_op_add(ed_op_sub(ba, ed_op_add(Cz, Cz)), P3), H = ed_op_sub(ed_op_sub(P2, A2), B2);
        gadget(E, Fq, 13, true, Cd[0]);
        gadget(Gq, H, 14, true, Cd[1]);
        gadget(E, H, 15, true, Cd[3]);
        gadget(Fq, Gq, 16, true, Cd[2]);
    }
    // ---- the 128 transition constraints share the periodic selector ----
    acc = gl_add(gl_mul(acc, k.a128), gl_mul(notend, gl_canon(T)));
    // ---- block initialisation: 64 limbs of res - O, 64 of temp - B (both on S0), 64 of res - O and 16 of temp.Z - 1 (H0);
    //      sum over (co, i) of a^(63 - 16 co - i) x = sum over co of a^(48 - 16 co) tilde(x_co) ----
    const gl zero = 0;
    const gl o_tilde[4] = {zero, k.a15, k.a15, zero};  // O = (0, 1, 1, 0), the limbs of 1 are (1, 0, ..., 0)
    const gl wgt[4] = {k.a48, k.a32, k.a16, 1};
    gl res_o = 0, tmp_b = 0;
#pragma unroll
    for (int co = 0; co < 4; co++) {
        res_o = gl_add(res_o, gl_mul(wgt[co], gl_sub(R[co], o_tilde[co])));
        tmp_b = gl_add(tmp_b, gl_mul(wgt[co], gl_sub(S[co], k.base_tilde[co])));
    }
    acc = gl_add(gl_mul(acc, k.a64), gl_mul(s0, res_o));
    acc = gl_add(gl_mul(acc, k.a64), gl_mul(s0, tmp_b));
    acc = gl_add(gl_mul(acc, k.a64), gl_mul(h0, res_o));
    acc = gl_add(gl_mul(acc, k.a16), gl_mul(h0, gl_sub(S[2], k.a15)));
    return acc;
}

}  // namespace tmx
