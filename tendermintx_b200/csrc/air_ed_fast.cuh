// Factored evaluation of the Ed25519 table's random-linear-combination of constraints (K5 fast path).
//
// air_ed25519() in air.cuh is the definition: 1 + 17 * 32 + 128 + 208 constraints folded by Horner's rule in the
// challenge alpha.  Evaluated literally that costs ~9,000 field multiplications per LDE point, almost all of them
// the limb products U_i V_j and q_i p_j of the 17 multiplication gadgets.  Because the 32 limb equations of a
// gadget carry CONSECUTIVE powers of alpha, their combination is a product of two polynomial evaluations:
//
//   sum_k alpha^(31-k) s_k  =  alpha U~ V~  -  alpha^16 C~  -  Q~ P~  +  (1 - 2^16 alpha) W~
//
//   X~ = sum_i X_i alpha^(len-1-i)   (Horner over the limbs, 15 / 16 / 30 multiply-adds),
//   W~ over w_m - ED_W_OFFSET, P~ and the 2d constant's tilde depend on alpha only.
//
// Operands that are limb-wise linear combinations of earlier cells (E = B - A + p, ...) have tildes that are the same
// linear combinations of tildes, so every trace cell enters exactly one Horner chain per challenge: ~1,500
// multiply-adds per (point, challenge) instead of ~5,000.  The result is the SAME field element as the literal
// evaluation (polynomial identity over F_p), so proof bytes do not change; tests pin it against air_ed25519()
// on the host (tmx_host_air_ed25519) and against the oracle's proofs on the GPU.
#pragma once
#include "air.cuh"

namespace tmx {

struct EdFastConsts {
    gl a, a15, a16, a32, a48, a64, a128;
    gl base_tilde[4];  // tildes of the base point's X, Y, Z (= 1), T limbs: the start of every [s]B ladder
    gl p_tilde;      // sum_j p_j a^(15-j), p = 2^255 - 19 in 16-bit limbs
    gl twod_tilde;   // same for the curve constant 2d
    gl w_off;        // ED_W_OFFSET * sum_{e<=30} a^e
    gl one_m_216a;   // 1 - 2^16 a
};

inline EdFastConsts ed_fast_consts(gl a) {
    static const uint64_t TWOD[16] = {0xF159, 0x26B2, 0x9B94, 0xEBD6, 0xB156, 0x8283, 0x149A, 0x00E0,
                                      0xD130, 0xEEF3, 0x80F2, 0x198E, 0xFCE7, 0x56DF, 0xD9DC, 0x2406};
    EdFastConsts k;
    k.a = a;
    k.a15 = gl_pow(a, 15);
    k.a16 = gl_pow(a, 16);
    k.a32 = gl_pow(a, 32);
    k.a48 = gl_pow(a, 48);
    k.a64 = gl_pow(a, 64);
    k.a128 = gl_pow(a, 128);
    {
        static const uint64_t BXL[16] = ED_BASE_X_LIMBS, BYL[16] = ED_BASE_Y_LIMBS, BTL[16] = ED_BASE_T_LIMBS;
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
    gl pt = 0, tt = 0, s = 0;
    for (int i = 0; i < 16; i++) {
        pt = gl_add(gl_mul(pt, a), (gl)p25519_limb(i));
        tt = gl_add(gl_mul(tt, a), (gl)TWOD[i]);
    }
    for (int e = 0; e <= 30; e++) s = gl_add(gl_mul(s, a), 1);
    k.p_tilde = pt;
    k.twod_tilde = tt;
    k.w_off = gl_mul((gl)ED_W_OFFSET, s);
    k.one_m_216a = gl_sub(1, gl_mul((gl)1 << 16, a));
    return k;
}

// Row: operator[](int col) -> FB (canonical cell).  per = {not_block_end, first row of [s]B, first row of [h]A}.
// Returns the Horner-folded constraint value for challenge k.a.
template <class Row>
TMX_HD gl ed25519_constraints_fast(const Row& l, const Row& n, const gl per[3], const EdFastConsts& k) {
    const gl notend = per[0], s0 = per[1], h0 = per[2];
    const gl a = k.a;
    auto G = [](int m) { return ED_MUL + m * ED_MUL_STRIDE; };
    const gl bit = l[ED_BIT].v;
    gl acc = gl_mul(bit, gl_sub(bit, 1));
    // ---- pass 1: accumulator / running-double coordinates, the gadget outputs they select from, transitions ----
    const int sum_slot[4] = {5, 6, 8, 7}, dbl_slot[4] = {13, 14, 16, 15};  // X, Y, Z, T
    gl R[4], S[4], Cs[4], Cd[4], T = 0;
#pragma unroll
    for (int co = 0; co < 4; co++) {
        gl r_t = 0, s_t = 0, cs_t = 0, cd_t = 0;
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const gl r = l[ED_RES + 16 * co + i].v, t = l[ED_TMP + 16 * co + i].v;
            const gl s = l[G(sum_slot[co]) + i].v, d = l[G(dbl_slot[co]) + i].v;
            const gl nr = n[ED_RES + 16 * co + i].v, nt = n[ED_TMP + 16 * co + i].v;
            r_t = gl_mac_nc(r_t, a, r);
            s_t = gl_mac_nc(s_t, a, t);
            cs_t = gl_mac_nc(cs_t, a, s);
            cd_t = gl_mac_nc(cd_t, a, d);
            T = gl_mac_nc(T, a, gl_sub(nr, gl_add(r, gl_mul(bit, gl_sub(s, r)))));
            T = gl_mac_nc(T, a, gl_sub(nt, d));
        }
        R[co] = gl_canon(r_t);
        S[co] = gl_canon(s_t);
        Cs[co] = gl_canon(cs_t);
        Cd[co] = gl_canon(cd_t);
    }
    // ---- pass 2: the 17 gadgets in emission order ----
    auto tilde = [&](int col0, int len) {
        gl t = l[col0].v;
#pragma unroll 4
        for (int i = 1; i < len; i++) t = gl_mac_nc(t, a, l[col0 + i].v);
        return gl_canon(t);
    };
    // c_known < 0: read the product limbs; otherwise their tilde is already known from pass 1
    auto gadget = [&](gl u, gl v, int g, bool have_c, gl c_t) {
        if (!have_c) c_t = tilde(G(g), 16);
        const gl q_t = tilde(G(g) + ED_MUL_Q, 17);
        const gl w_t = gl_sub(tilde(G(g) + ED_MUL_W, 31), k.w_off);
        gl s = gl_mul(a, gl_mul(u, v));
        s = gl_sub(s, gl_mul(k.a16, c_t));
        s = gl_sub(s, gl_mul(q_t, k.p_tilde));
        s = gl_add(s, gl_mul(k.one_m_216a, w_t));
        acc = gl_add(gl_mul(acc, k.a32), s);
        return c_t;
    };
    const gl P = k.p_tilde, P2 = gl_add(P, P), P3 = gl_add(P2, P);
    const gl X1 = R[0], Y1 = R[1], Z1 = R[2], T1 = R[3], X2 = S[0], Y2 = S[1], Z2 = S[2], T2 = S[3];
    const gl A = gadget(gl_add(gl_sub(Y1, X1), P), gl_add(gl_sub(Y2, X2), P), 0, false, 0);
    const gl B = gadget(gl_add(Y1, X1), gl_add(Y2, X2), 1, false, 0);
    const gl TT = gadget(T1, T2, 2, false, 0);
    const gl C = gadget(TT, k.twod_tilde, 3, false, 0);
    const gl Dh = gadget(Z1, Z2, 4, false, 0);
    {
        const gl d2 = gl_add(Dh, Dh);
        const gl E = gl_add(gl_sub(B, A), P), Fq = gl_add(gl_sub(d2, C), P), Gq = gl_add(d2, C), H = gl_add(B, A);
        gadget(E, Fq, 5, true, Cs[0]);
        gadget(Gq, H, 6, true, Cs[1]);
        gadget(E, H, 7, true, Cs[3]);
        gadget(Fq, Gq, 8, true, Cs[2]);
    }
    const gl A2 = gadget(X2, X2, 9, false, 0);
    const gl B2 = gadget(Y2, Y2, 10, false, 0);
    const gl Cz = gadget(Z2, Z2, 11, false, 0);
    const gl xy = gl_add(X2, Y2);
    const gl Sq = gadget(xy, xy, 12, false, 0);
    {
        const gl ba = gl_sub(B2, A2);
        const gl E = gl_add(gl_sub(gl_sub(Sq, A2), B2), P2), Gq = gl_add(ba, P);
        const gl Fq = gl_add(gl_sub(ba, gl_add(Cz, Cz)), P3), H = gl_sub(gl_sub(P2, A2), B2);
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
