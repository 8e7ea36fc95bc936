// Logic table, host side: the row schedule of a circuit shape (which gadget instance sits on which row, its message ids and
// multiplicities: the constant columns), the netlist of the per-validator curve25519 multiplications, the witness of the
// table, and the verifier's public bus terms.  The constraints are in logic.cuh (air_logic); this file decides what they are
// applied to.  Host code only: the table has a few thousand rows.
//
// Follows the call structure of verify_skip / verify_step [REF circuits/builder/verify.rs:469-563]: verify_trusted_validators
// (361-437), verify_header (224-334), verify_prev_header_in_header / ..._next_validators_hash (137-178).
#include "logic_plan.cuh"
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace tmx {

// ------------------------------------------------------------------------------------------ netlist of one validator slot
// wire index: 0 .. 63 slot-local (W_*), 64 + c constant wire c
struct Term {
    int wire, coef;
};
struct Operand {
    Term t[3];
    int nt, kp;
};
struct Gadget {
    Operand U, V, W;
    bool assert0;
};
static constexpr int CW = 64;
static Operand op(std::initializer_list<Term> ts, int kp = 0) {
    Operand o;
    o.nt = 0;
    o.kp = kp;
    for (const Term& t : ts) o.t[o.nt++] = t;
    return o;
}
static const std::vector<Gadget>& netlist() {
    static const std::vector<Gadget> g = [] {
        const int G = W_G0;
        std::vector<Gadget> v;
        auto add = [&](Operand U, Operand V, Operand W = op({}), bool a = false) { v.push_back(Gadget{U, V, W, a}); };
        // decompression of A: x^2 (d y^2 + 1) - (y^2 - 1) = 0 [RFC 8032 5.1.3]
        add(op({{W_YA, 1}}), op({{W_YA, 1}}));                                                                   // 0  y^2
        add(op({{W_XA, 1}}), op({{W_XA, 1}}));                                                                   // 1  x^2
        add(op({{G + 0, 1}}), op({{CW + WC_D, 1}}));                                                            // 2  d y^2
        add(op({{G + 1, 1}}), op({{G + 2, 1}, {CW + WC_ONE, 1}}), op({{CW + WC_ONE, 1}, {G + 0, -1}}, 1), true);  // 3
        // decompression of R
        add(op({{W_YR, 1}}), op({{W_YR, 1}}));                                                                   // 4
        add(op({{W_XR, 1}}), op({{W_XR, 1}}));                                                                   // 5
        add(op({{G + 4, 1}}), op({{CW + WC_D, 1}}));                                                            // 6
        add(op({{G + 5, 1}}), op({{G + 6, 1}, {CW + WC_ONE, 1}}), op({{CW + WC_ONE, 1}, {G + 4, -1}}, 1), true);  // 7
        // cached form of -A: t = x y, 2 d t
        add(op({{W_XA, 1}}), op({{W_YA, 1}}));                                                                   // 8  tA
        add(op({{G + 8, 1}}), op({{CW + WC_D2, 1}}));                                                           // 9  2 d tA
        // D = B + (-A), mixed addition of two affine points (add-2008-hwcd-3, Z1 = Z2 = 1)
        add(op({{CW + WC_BYMX, 1}}), op({{W_YA, 1}, {W_XA, 1}}));                                                // 10 a = (By - Bx)(yA + xA)
        add(op({{CW + WC_BYPX, 1}}), op({{W_YA, 1}, {W_XA, -1}}, 1));                                            // 11 b = (By + Bx)(yA - xA)
        add(op({{CW + WC_BT2D, 1}}), op({{G + 8, 1}}));                                                         // 12 cc = 2 d tB tA  (C = -cc)
        add(op({{G + 11, 1}, {G + 10, -1}}, 1), op({{CW + WC_ONE, 2}, {G + 12, 1}}));                            // 13 X3 = (b - a)(2 + cc)
        add(op({{CW + WC_ONE, 2}, {G + 12, -1}}, 1), op({{G + 11, 1}, {G + 10, 1}}));                            // 14 Y3 = (2 - cc)(b + a)
        add(op({{CW + WC_ONE, 2}, {G + 12, 1}}), op({{CW + WC_ONE, 2}, {G + 12, -1}}, 1));                       // 15 Z3 = (2 + cc)(2 - cc)
        add(op({{W_XD, 1}}), op({{G + 15, 1}}), op({{G + 13, -1}}, 1), true);                                    // 16 xD Z3 = X3
        add(op({{W_YD, 1}}), op({{G + 15, 1}}), op({{G + 14, -1}}, 1), true);                                    // 17 yD Z3 = Y3
        add(op({{W_XD, 1}}), op({{W_YD, 1}}));                                                                   // 18 tD
        add(op({{G + 18, 1}}), op({{CW + WC_D2, 1}}));                                                          // 19 2 d tD
        // result of the Ed25519 table: Q = [s]B + [h](-A) equals R:  xR ZQ = XQ, yR ZQ = YQ
        add(op({{W_XR, 1}}), op({{W_ZQ, 1}}), op({{W_XQ, -1}}, 1), true);                                        // 20
        add(op({{W_YR, 1}}), op({{W_ZQ, 1}}), op({{W_YQ, -1}}, 1), true);                                        // 21
        return v;
    }();
    return g;
}
// distinct wires of a gadget in first-use order (its input slots)
static int gadget_slots(const Gadget& g, int slots[6]) {
    int n = 0;
    for (const Operand* o : {&g.U, &g.V, &g.W})
        for (int i = 0; i < o->nt; i++) {
            bool seen = false;
            for (int s = 0; s < n; s++) seen |= slots[s] == o->t[i].wire;
            if (!seen) slots[n++] = o->t[i].wire;
        }
    return n;
}

// ------------------------------------------------------------------------------------------ plan
static uint64_t fneg(int c) { return c >= 0 ? (uint64_t)c : GL_P - (uint64_t)(-c); }

static void set_tags(LogicPlan& p, size_t row, std::initializer_list<std::pair<int, int>> groups_tag) {
    // default: unused groups are not looked up
    for (int g = 0; g < LG_GROUPS; g++) {
        p.k(LGK_TAG + g, row) = BUS_R16;
        p.k(LGK_USE + g, row) = 0;
    }
    int g = 0;
    for (auto& gt : groups_tag) {
        for (int i = 0; i < gt.first; i++, g++) {
            if (gt.second) {
                p.k(LGK_TAG + g, row) = (uint64_t)gt.second;
                p.k(LGK_USE + g, row) = 1;
            }
        }
    }
}

std::shared_ptr<const LogicPlan> logic_plan_get(AirShape sh) {
    static std::mutex mu;
    static std::map<std::string, std::shared_ptr<const LogicPlan>> cache;
    const std::string key = std::to_string(sh.kind) + "/" + std::to_string(sh.n_max) + "/" + std::string(sh.chain, sh.chain_len);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto pp = std::make_shared<LogicPlan>();
    LogicPlan& p = *pp;
    p.sh = sh;
    p.n_rows = air_pow2_at_least(logic_used_rows(sh));
    p.K.assign((size_t)LGK_COLS * p.n_rows, 0);
    const uint32_t N = sh.n_max, np = (uint32_t)air_pow2_at_least(N), log_np = ilog2(np);
    const uint32_t sets = sh.kind == TMX_KIND_SKIP ? 2 : 1, n_proofs = sh.kind == TMX_KIND_SKIP ? 4 : 5;
    const size_t set_chunks = (size_t)N + 2 * ((size_t)np - 1);
    for (size_t r = 0; r < p.n_rows; r++) set_tags(p, r, {});
    size_t row = 0;
    // ---- wire uses per slot (netlist + the rows that consume wires directly)
    int uses[64] = {0}, const_uses[WC_COUNT] = {0};
    for (const Gadget& g : netlist()) {
        int slots[6];
        const int ns = gadget_slots(g, slots);
        for (int s = 0; s < ns; s++) (slots[s] >= CW ? const_uses[slots[s] - CW] : uses[slots[s]])++;
    }
    uses[W_YA] += 2;       // EDIO, XS
    uses[W_XA] += 1;       // EDIO
    uses[W_YR] += 1;       // XS
    uses[W_G0 + 9] += 1;   // 2 d tA -> EDIO
    uses[W_G0 + 19] += 1;  // 2 d tD -> EDIO
    // ---- GLOB
    p.row_glob = row;
    p.k(LGK_SEL + LT_GLOB, row) = 1;
    set_tags(p, row, {{4, BUS_R8}, {1, BUS_R1}});
    row++;
    // ---- constant field elements
    p.row_cfe = row;
    for (int c = 0; c < WC_COUNT; c++, row++) {
        p.k(LGK_SEL + LT_CFE, row) = 1;
        set_tags(p, row, {{1, BUS_R16}});
        for (int j = 0; j < 16; j++) p.k(LGK_P + CFP_LIMB + j, row) = logic_const_wire_limb(c, j);
        p.k(LGK_P + CFP_ID, row) = wid_const(c);
        p.k(LGK_P + CFP_MULT, row) = (uint64_t)const_uses[c] * N;
    }
    // ---- validator leaves
    auto h1_tags = [&](size_t r) { set_tags(p, r, {{7, BUS_R8}, {5, BUS_R1}, {1, BUS_R16}}); };
    const uint32_t root_uses = sh.kind == TMX_KIND_SKIP ? 1 : 2;
    for (uint32_t s = 0; s < sets; s++) {
        p.row_leaf[s] = row;
        const bool target = s == sets - 1;
        for (uint32_t i = 0; i < N; i++, row++) {
            p.k(LGK_SEL + LT_H1, row) = 1;
            h1_tags(row);
            p.k(LGK_P + H1P_CID, row) = s * set_chunks + i;
            p.k(LGK_P + H1P_SEND, row) = 1;
            p.k(LGK_P + H1P_FVAL, row) = 1;
            p.k(LGK_P + H1P_OUT_ID, row) = nid_tree(s, 0, i);
            p.k(LGK_P + H1P_OUT_MULT, row) = log_np == 0 ? root_uses : 1;
            p.k(LGK_P + H1P_TARGET, row) = target;
            p.k(LGK_P + H1P_TRUSTED, row) = !target;
            p.k(LGK_P + H1P_SETNEXT, row) = i + 1 < N;
            p.k(LGK_P + H1P_SETFIRST, row) = i == 0;
            p.k(LGK_P + H1P_LAST2, row) = target && i + 1 == N;
            p.k(LGK_P + H1P_LAST1, row) = !target && i + 1 == N;
            p.k(LGK_P + H1P_VID, row) = i;
        }
    }
    // ---- inner nodes of the validator-set trees
    auto h2_tags = [&](size_t r) { set_tags(p, r, {{7, BUS_R8}, {1, BUS_R1}}); };
    for (uint32_t s = 0; s < sets; s++) {
        p.row_inner[s] = row;
        for (uint32_t l = 1; l <= log_np; l++)
            for (uint32_t i = 0; i < (np >> l); i++, row++) {
                p.k(LGK_SEL + LT_H2, row) = 1;
                h2_tags(row);
                const size_t inner_off = (size_t)np - ((size_t)np >> (l - 1));
                p.k(LGK_P + H2P_CID, row) = s * set_chunks + N + 2 * (inner_off + i);
                p.k(LGK_P + H2P_SEND, row) = 1;
                p.k(LGK_P + H2P_IS65, row) = 1;
                const bool abs_l = l == 1 && 2 * i >= N, abs_r = l == 1 && 2 * i + 1 >= N;
                p.k(LGK_P + H2P_ID_L, row) = nid_tree(s, l - 1, 2 * i);
                p.k(LGK_P + H2P_RECV_L, row) = !abs_l;
                p.k(LGK_P + H2P_ABS_L, row) = abs_l;
                p.k(LGK_P + H2P_ID_R, row) = nid_tree(s, l - 1, 2 * i + 1);
                p.k(LGK_P + H2P_RECV_R, row) = !abs_r;
                p.k(LGK_P + H2P_ABS_R, row) = abs_r;
                p.k(LGK_P + H2P_OUT_ID, row) = nid_tree(s, l, i);
                p.k(LGK_P + H2P_OUT_MULT, row) = l == log_np ? root_uses : 1;
            }
    }
    // ---- header proofs: leaf, then four inner nodes; the verifier consumes each root
    p.row_hdr = row;
    for (uint32_t kp = 0; kp < n_proofs; kp++) {
        // which logical proof (witness_jobs.cuh header_proof_desc): 0 aux-valhash(skip) 1 valhash 2 chain 3 height 4 last-block-id
        // 5 aux-next-valhash(step)
        const int which = sh.kind == TMX_KIND_SKIP ? (int)kp : (int)kp + 1;
        const size_t chunk0 = sets * set_chunks + 9 * (size_t)kp + (sh.kind == TMX_KIND_STEP && kp > 3 ? 1 : 0);
        const unsigned index = which == 2 ? TMX_CHAIN_ID_INDEX : which == 3 ? TMX_BLOCK_HEIGHT_INDEX : which == 4 ? TMX_LAST_BLOCK_ID_INDEX
                               : which == 5 ? TMX_NEXT_VALIDATORS_HASH_INDEX : TMX_VALIDATORS_HASH_INDEX;
        size_t chunk = chunk0;
        if (which == 4) {  // 72-byte leaf: two chunks, fixed layout
            p.k(LGK_SEL + LT_H2, row) = 1;
            h2_tags(row);
            p.k(LGK_P + H2P_CID, row) = chunk;
            p.k(LGK_P + H2P_SEND, row) = 1;
            p.k(LGK_P + H2P_IS73, row) = 1;
            p.k(LGK_P + H2P_OUT_ID, row) = nid_header(kp, 0);
            p.k(LGK_P + H2P_OUT_MULT, row) = 1;
            p.k(LGK_P + H2P_PUBPREV, row) = 1;
            chunk += 2;
        } else {
            p.k(LGK_SEL + LT_H1, row) = 1;
            h1_tags(row);
            p.k(LGK_P + H1P_CID, row) = chunk;
            p.k(LGK_P + H1P_SEND, row) = 1;
            p.k(LGK_P + (which == 2 ? H1P_FCHAIN : which == 3 ? H1P_FHEIGHT : H1P_FHASH), row) = 1;
            p.k(LGK_P + H1P_OUT_ID, row) = nid_header(kp, 0);
            p.k(LGK_P + H1P_OUT_MULT, row) = 1;
            if (which == 0 || which == 1 || which == 5) {  // the hash in the leaf is a validator-set root
                const uint32_t set = which == 0 ? 0 : sets - 1;
                p.k(LGK_P + H1P_IN_ID, row) = nid_tree(set, log_np, 0);
                p.k(LGK_P + H1P_IN_MULT, row) = 1;
            }
            chunk += 1;
        }
        row++;
        for (uint32_t j = 1; j <= 4; j++, row++, chunk += 2) {
            p.k(LGK_SEL + LT_H2, row) = 1;
            h2_tags(row);
            const bool right = (index >> (j - 1)) & 1;  // the running node is the right child, the aunt is on the left
            p.k(LGK_P + H2P_CID, row) = chunk;
            p.k(LGK_P + H2P_SEND, row) = 1;
            p.k(LGK_P + H2P_IS65, row) = 1;
            p.k(LGK_P + (right ? H2P_ID_R : H2P_ID_L), row) = nid_header(kp, j - 1);
            p.k(LGK_P + (right ? H2P_RECV_R : H2P_RECV_L), row) = 1;
            p.k(LGK_P + H2P_OUT_ID, row) = nid_header(kp, j);
            p.k(LGK_P + H2P_OUT_MULT, row) = 1;
        }
    }
    // ---- per validator slot: SIG, SC, XS, EDIO, MUL rows
    p.row_slots = row;
    for (uint32_t i = 0; i < N; i++) {
        // SIG
        p.k(LGK_SEL + LT_SIG, row) = 1;
        set_tags(p, row, {{16, BUS_R8}, {9, BUS_R1}});
        p.k(LGK_P + SGP_VID, row) = i;
        p.k(LGK_P + SGP_YA_ID, row) = wid_slot(i, W_YA);
        p.k(LGK_P + SGP_YA_MULT, row) = uses[W_YA];
        p.k(LGK_P + SGP_YR_ID, row) = wid_slot(i, W_YR);
        p.k(LGK_P + SGP_YR_MULT, row) = uses[W_YR];
        p.k(LGK_P + SGP_SA_ID, row) = wid_slot(i, W_SA);
        p.k(LGK_P + SGP_SR_ID, row) = wid_slot(i, W_SR);
        row++;
        // SC
        p.k(LGK_SEL + LT_SC, row) = 1;
        set_tags(p, row, {{2, BUS_R8}, {1, BUS_R16}, {2, BUS_R16}, {1, BUS_R8}, {4, BUS_R8}, {1, BUS_R16}, {1, BUS_R1}, {1, BUS_R16}, {1, BUS_R1}});
        p.k(LGK_P + SCP_VID, row) = i;
        row++;
        // XS
        p.k(LGK_SEL + LT_XS, row) = 1;
        set_tags(p, row, {{4, BUS_R16}, {2, BUS_R1}, {4, BUS_R16}, {2, BUS_R1}, {1, BUS_R16}, {1, BUS_R1}});
        for (int a = 0; a < 2; a++) {
            p.k(LGK_P + XSP_X_ID + a * XSP_STRIDE, row) = wid_slot(i, a ? W_XR : W_XA);
            p.k(LGK_P + XSP_X_MULT + a * XSP_STRIDE, row) = uses[a ? W_XR : W_XA];
            p.k(LGK_P + XSP_Y_ID + a * XSP_STRIDE, row) = wid_slot(i, a ? W_YR : W_YA);
            p.k(LGK_P + XSP_SIGN_ID + a * XSP_STRIDE, row) = wid_slot(i, a ? W_SR : W_SA);
        }
        row++;
        // EDIO
        p.k(LGK_SEL + LT_EDIO, row) = 1;
        set_tags(p, row, {{9, BUS_R16}});
        p.k(LGK_P + EIP_VID, row) = i;
        p.k(LGK_P + EIP_YA_ID, row) = wid_slot(i, W_YA);
        p.k(LGK_P + EIP_XA_ID, row) = wid_slot(i, W_XA);
        p.k(LGK_P + EIP_T2A_ID, row) = wid_slot(i, W_G0 + 9);
        p.k(LGK_P + EIP_T2D_ID, row) = wid_slot(i, W_G0 + 19);
        p.k(LGK_P + EIP_XD_ID, row) = wid_slot(i, W_XD);
        p.k(LGK_P + EIP_XD_MULT, row) = uses[W_XD];
        p.k(LGK_P + EIP_YD_ID, row) = wid_slot(i, W_YD);
        p.k(LGK_P + EIP_YD_MULT, row) = uses[W_YD];
        p.k(LGK_P + EIP_XQ_ID, row) = wid_slot(i, W_XQ);
        p.k(LGK_P + EIP_XQ_MULT, row) = uses[W_XQ];
        p.k(LGK_P + EIP_YQ_ID, row) = wid_slot(i, W_YQ);
        p.k(LGK_P + EIP_YQ_MULT, row) = uses[W_YQ];
        p.k(LGK_P + EIP_ZQ_ID, row) = wid_slot(i, W_ZQ);
        p.k(LGK_P + EIP_ZQ_MULT, row) = uses[W_ZQ];
        row++;
        // MUL rows: two gadgets each
        const std::vector<Gadget>& nl = netlist();
        for (size_t g0 = 0; g0 < nl.size(); g0 += 2, row++) {
            p.k(LGK_SEL + LT_MUL, row) = 1;
            set_tags(p, row, {{8, 0}, {3, BUS_R16}, {1, BUS_R11}, {8, 0}, {3, BUS_R16}, {1, BUS_R11}});
            for (int h = 0; h < 2 && g0 + h < nl.size(); h++) {
                const Gadget& g = nl[g0 + h];
                const int q = LGK_P + h * MUP_STRIDE;
                int slots[6];
                const int ns = gadget_slots(g, slots);
                p.k(q + MUP_ACTIVE, row) = 1;
                for (int s = 0; s < ns; s++) {
                    p.k(q + MUP_IN_ID + s, row) = slots[s] >= CW ? wid_const(slots[s] - CW) : wid_slot(i, slots[s]);
                    p.k(q + MUP_IN_RECV + s, row) = 1;
                }
                const Operand* ops[3] = {&g.U, &g.V, &g.W};
                const int cbase[3] = {MUP_CU, MUP_CV, MUP_CW}, kbase[3] = {MUP_KPU, MUP_KPV, MUP_KPW};
                for (int o = 0; o < 3; o++) {
                    for (int t = 0; t < ops[o]->nt; t++)
                        for (int s = 0; s < ns; s++)
                            if (slots[s] == ops[o]->t[t].wire) p.k(q + cbase[o] + s, row) = fneg(ops[o]->t[t].coef);
                    p.k(q + kbase[o], row) = fneg(ops[o]->kp);
                }
                p.k(q + MUP_OUT_ID, row) = wid_slot(i, W_G0 + (int)(g0 + h));
                p.k(q + MUP_OUT_MULT, row) = uses[W_G0 + (int)(g0 + h)];
                p.k(q + MUP_ASSERT, row) = g.assert0;
            }
        }
    }
    p.used_rows = row;
    cache[key] = pp;
    return pp;
}

uint64_t logic_const_value(int kc, size_t row, AirShape sh) {
    static thread_local std::shared_ptr<const LogicPlan> last;
    if (!last || last->sh.kind != sh.kind || last->sh.n_max != sh.n_max || last->sh.chain_len != sh.chain_len ||
        memcmp(last->sh.chain, sh.chain, sh.chain_len))
        last = logic_plan_get(sh);
    return last->K[(size_t)kc * last->n_rows + row];
}

// ------------------------------------------------------------------------------------------ public terms
static gl2 fingerprint(gl2 beta, gl2 gamma, uint64_t tag, const std::vector<uint64_t>& v) {
    gl2 acc = gl2_from(0);
    for (size_t i = v.size(); i-- > 0;) {
        acc.a0 = gl_add(acc.a0, v[i] % GL_P);
        acc = gl2_mul(acc, beta);
    }
    acc.a0 = gl_add(acc.a0, tag);
    return gl2_add(acc, gamma);
}
static uint64_t be64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}
static uint64_t be32w(const uint8_t* p) { return ((uint64_t)p[0] << 24) | ((uint64_t)p[1] << 16) | ((uint64_t)p[2] << 8) | p[3]; }

// The verifier's side of the bus [REF circuits/skip.rs:119-133, step.rs:106-117 for the public input / output layout]:
//   - consumes the root of every header proof: the trusted header (skip) / the proven header (out32) / the previous header
//   - provides the height as varint septets to the height leaf, the previous header to the last-block-id leaf (step), and
//     (height, proven header) to the GLOB row that the sign-bytes checks read.
gl2 logic_public_terms(AirShape sh, uint64_t, const uint8_t* input, const uint8_t* out32, gl2 beta, gl2 gamma) {
    if (!logic_rows(sh)) return gl2_from(0);
    const bool skip = sh.kind == TMX_KIND_SKIP;
    const uint64_t height = skip ? be64(input + 40) : be64(input) + 1;
    const uint8_t* other = input + 8;  // trusted header (skip) / previous header (step)
    gl2 sum = gl2_from(0);
    auto term = [&](int sign, uint64_t tag, const std::vector<uint64_t>& v) {
        const gl2 inv = gl2_inv(fingerprint(beta, gamma, tag, v));
        sum = sign > 0 ? gl2_add(sum, inv) : gl2_sub(sum, inv);
    };
    auto node = [&](uint32_t proof, const uint8_t* hdr) {
        std::vector<uint64_t> v{nid_header(proof, 4)};
        for (int i = 0; i < 8; i++) v.push_back(be32w(hdr + 4 * i));
        v.push_back(1);
        term(-1, BUS_NODE, v);
    };
    if (skip) {
        node(0, other);
        for (uint32_t kp = 1; kp < 4; kp++) node(kp, out32);
    } else {
        for (uint32_t kp = 0; kp < 4; kp++) node(kp, out32);
        node(4, other);
        std::vector<uint64_t> v{PUB_PREV};
        for (int i = 0; i < 8; i++) v.push_back(be32w(other + 4 * i));
        term(+1, BUS_PUB, v);
    }
    {
        std::vector<uint64_t> v{PUB_HEIGHT};
        for (int kk = 0; kk < 9; kk++) v.push_back((height >> (7 * kk)) & 0x7F);
        term(+1, BUS_PUB, v);
    }
    {
        std::vector<uint64_t> v{PUB_GLOB, height & 0xFFFFFFFFULL, height >> 32};
        for (int i = 0; i < 8; i++) v.push_back(be32w(out32 + 4 * i));
        term(+1, BUS_PUB, v);
    }
    return sum;
}

// ------------------------------------------------------------------------------------------ witness
namespace {

struct Filler {
    const LogicPlan& p;
    gl* t;
    gl& c(int col, size_t row) { return t[(size_t)col * p.n_rows + row]; }
};

void sha256_bytes(const uint8_t* msg, int len, uint8_t out[32]) {
    uint8_t buf[192];
    Sha256Hist hs;
    const int nb = sha256_pad_blocks(msg, len, buf);
    uint32_t st[8];
    for (int k = 0; k < 8; k++) st[k] = iv256(k);
    for (int b = 0; b < nb; b++) sha256_compress_hist(st, buf + 64 * b, &hs, st);
    sha256_state_to_bytes(st, out);
}

void limbs_of(const fe256& x, int32_t l[16]) { fe256_to_limbs(x, l); }

// v + d + 1 = K on 16-bit limbs: d limbs and the 15 carries
void less_than_cells(const int32_t v[16], uint64_t (*kl)(int), gl* d, gl* c, size_t stride) {
    int64_t borrow = 0;
    int32_t dl[16];
    for (int i = 0; i < 16; i++) {  // d = K - 1 - v
        int64_t s = (int64_t)kl(i) - (i == 0) - v[i] + borrow;
        borrow = s >> 16;
        dl[i] = (int32_t)(s & 0xFFFF);
    }
    int64_t carry = 1;
    for (int i = 0; i < 16; i++) {
        d[(size_t)i * stride] = (gl)dl[i];
        const int64_t s = (int64_t)v[i] + dl[i] + carry;
        carry = s >> 16;
        if (i < 15) c[(size_t)i * stride] = (gl)carry;
    }
}

}  // namespace

// Fills the logic table ([LG_COLS][n_rows], column-major, zero-initialised by the caller) for one proof.  Returns 0, or a
// check id when the inputs cannot satisfy the table (the caller's pre-check reports the same condition first).
int logic_fill_trace(const LogicPlan& p, const uint8_t* input, const uint8_t* blob, const EdSlotInfo* slots, gl* trace, bool force) {
    Filler f{p, trace};
    int status = 0;
    // an unsatisfiable input: stop (honest prover) or remember the first failing check and keep filling (force: the tests' cheating
    // provers, and the prover itself, which reports the pre-check's verdict first)
#define LOGIC_FAIL(id)                   \
    do {                                 \
        if (!status) status = (id);      \
        if (!force) return status;       \
    } while (0)
    const AirShape& sh = p.sh;
    const uint32_t N = sh.n_max, np = (uint32_t)air_pow2_at_least(N), log_np = ilog2(np);
    const bool skip = sh.kind == TMX_KIND_SKIP;
    const uint32_t sets = skip ? 2 : 1, n_proofs = skip ? 4 : 5;
    const tmx_offchain_head* h = blob_head(blob);
    const tmx_validator* vals = blob_validators(blob);
    const tmx_hash_field* tf = blob_hash_fields(blob, N);
    const uint64_t height = skip ? be64(input + 40) : be64(input) + 1;
    const size_t n = p.n_rows;
    auto put_bytes = [&](int col0, size_t row, const uint8_t* b, int len) {
        for (int i = 0; i < len; i++) f.c(col0 + i, row) = b[i];
    };
    auto put_limbs = [&](int col0, size_t row, const int32_t l[16]) {
        for (int i = 0; i < 16; i++) f.c(col0 + i, row) = (gl)(int64_t)l[i];
    };
    // varint cells of value v at an H1 row
    auto put_varint = [&](size_t row, uint64_t v, uint8_t bytes[9]) {
        int last = 0;
        for (int k = 0; k < 9; k++)
            if ((v >> (7 * k)) & 0x7F) last = k;
        for (int k = 0; k < 9; k++) {
            const uint64_t g = (v >> (7 * k)) & 0x7F;
            f.c(H1_G7 + k, row) = g;
            f.c(H1_G72 + k, row) = 2 * g;
            bytes[k] = (uint8_t)(g | (k < last ? 0x80 : 0));
            if (k >= 1) {
                f.c(H1_NZ + k - 1, row) = k <= last;
                f.c(H1_GINV + k - 1, row) = g ? gl_inv(g) : 0;
                f.c(H1_GT + k - 1, row) = g ? 1 : 0;
            }
        }
    };
    // ---- signed keys, matching, multiplicities
    const uint32_t tset = sets - 1;
    std::vector<uint64_t> key_mult(N, 0);
    std::vector<uint8_t> flag(N, 0);
    if (skip)
        for (uint32_t j = 0; j < N && j < h->nb_trusted; j++)  // enabled slots only: padding never counts (logic.cuh)
            for (uint32_t i = 0; i < N; i++)
                if (vals[i].is_signed && !memcmp(vals[i].pubkey, tf[j].pubkey, 32)) {
                    flag[j] = 1;  // the lookup is answered by the first signed target validator with this key
                    key_mult[i]++;
                    break;
                }
    uint64_t n_signed = 0;
    for (uint32_t i = 0; i < N; i++) n_signed += vals[i].is_signed != 0;
    // ---- GLOB
    {
        const size_t r = p.row_glob;
        for (int i = 0; i < 8; i++) f.c(GB_HB + i, r) = (height >> (8 * i)) & 0xFF;
        put_bytes(GB_HDR, r, h->header, 32);
        uint64_t srb = 0;
        for (int i = 0; i < 8; i++) {
            const uint64_t b = (h->round >> (8 * i)) & 0xFF;
            f.c(GB_RB + i, r) = b;
            srb += b;
        }
        f.c(GB_RB7X2, r) = 2 * ((h->round >> 56) & 0xFF);
        f.c(GB_RZ, r) = srb == 0;
        f.c(GB_RINV, r) = srb ? gl_inv(srb) : 0;
        f.c(GB_MG, r) = n_signed;
    }
    for (int c = 0; c < WC_COUNT; c++)
        for (int j = 0; j < 16; j++) f.c(CF_C + j, p.row_cfe + c) = logic_const_wire_limb(c, j);
    // ---- validator sets: leaves, then trees
    std::vector<uint8_t> node((size_t)2 * np * 32), en((size_t)2 * np);
    uint8_t roots[2][32];
    for (uint32_t s = 0; s < sets; s++) {
        const bool trusted = skip && s == 0;
        const uint32_t nb = trusted ? h->nb_trusted : h->nb_validators;
        uint64_t tot = 0, sum = 0;
        for (uint32_t i = 0; i < np; i++) {
            uint8_t* nd = &node[(size_t)i * 32];
            if (i >= N) {
                memset(nd, 0, 32);
                en[i] = 0;
                continue;
            }
            const size_t r = p.row_leaf[s] + i;
            const uint8_t* pk = trusted ? tf[i].pubkey : vals[i].pubkey;
            const uint64_t power = trusted ? tf[i].voting_power : vals[i].voting_power;
            const uint32_t blen = trusted ? tf[i].validator_byte_length : vals[i].validator_byte_length;
            if (power >> 63) LOGIC_FAIL(15);
            if (blen < 38 || blen > 46) LOGIC_FAIL(17);
            uint8_t msg[56] = {0};
            uint8_t vb[9];
            put_varint(r, power, vb);
            msg[0] = 0; msg[1] = 0x0a; msg[2] = 0x22; msg[3] = 0x0a; msg[4] = 0x20;
            memcpy(msg + 5, pk, 32);
            msg[37] = 0x10;
            memcpy(msg + 38, vb, 9);
            put_bytes(H1_D, r, msg, 47);
            const int plen = 1 + (int)(blen > 46 ? 46 : blen);
            f.c(H1_IL + plen, r) = 1;
            sha256_bytes(msg, plen, nd);
            put_bytes(H1_DB, r, nd, 32);
            en[i] = i < nb;
            const bool sgn = !trusted && vals[i].is_signed;
            const bool fl = trusted && flag[i];
            f.c(H1_E, r) = en[i];
            f.c(H1_SGN, r) = sgn;
            f.c(H1_FLAG, r) = fl;
            f.c(H1_MK, r) = sgn ? key_mult[i] : 0;
            if (en[i]) tot += power;
            if (sgn || fl) sum += power;
            if (tot >> 62 || sum >> 62) LOGIC_FAIL(14);
            for (int k = 0; k < 4; k++) {
                f.c(H1_TOT + k, r) = (tot >> (16 * k)) & 0xFFFF;
                f.c(H1_SUM + k, r) = (sum >> (16 * k)) & 0xFFFF;
            }
            f.c(H1_TOT4, r) = 4 * ((tot >> 48) & 0xFFFF);
            f.c(H1_SUM4, r) = 4 * ((sum >> 48) & 0xFFFF);
            if (i + 1 == N) {
                const uint64_t lhs = 3 * sum, rhs = (trusted ? 1 : 2) * tot;
                if (lhs <= rhs) LOGIC_FAIL(trusted ? 4 : 8);
                const uint64_t df = lhs > rhs ? lhs - rhs - 1 : 0;
                for (int k = 0; k < 4; k++) f.c(H1_DF + k, r) = (df >> (16 * k)) & 0xFFFF;
                f.c(H1_DF2, r) = 2 * ((df >> 48) & 0xFFFF);
            }
        }
        // tree: level l nodes at [off_l, off_l + np >> l)
        size_t row = p.row_inner[s], off = 0;
        for (uint32_t l = 1; l <= log_np; l++) {
            const size_t cnt = np >> l, noff = off + (np >> (l - 1));
            for (size_t i = 0; i < cnt; i++, row++) {
                const uint8_t* L = &node[(off + 2 * i) * 32];
                const uint8_t* R = L + 32;
                const uint8_t eL = en[off + 2 * i], eR = en[off + 2 * i + 1];
                uint8_t msg[65], dg[32];
                msg[0] = 1;
                memcpy(msg + 1, L, 32);
                memcpy(msg + 33, R, 32);
                sha256_bytes(msg, 65, dg);
                put_bytes(H2_MB, row, msg, 65);
                put_bytes(H2_DB, row, dg, 32);
                f.c(H2_EL, row) = eL;
                f.c(H2_ER, row) = eR;
                f.c(H2_FF, row) = eL && eR;
                memcpy(&node[(noff + i) * 32], eL && eR ? dg : L, 32);
                en[noff + i] = eL;
            }
            off = noff;
        }
        memcpy(roots[s], &node[off * 32], 32);
    }
    // ---- header proofs
    {
        WitnessArgs wa;
        memset(&wa, 0, sizeof wa);
        wa.blob = blob;
        wa.kind = sh.kind;
        wa.n_max = N;
        wa.np = np;
        wa.log_np = log_np;
        size_t row = p.row_hdr;
        for (uint32_t kp = 0; kp < n_proofs; kp++) {
            HeaderProofDesc d;
            header_proof_desc(wa, kp, &d);
            uint8_t cur[32];
            const int which = skip ? (int)kp : (int)kp + 1;
            // leaf
            sha256_bytes(d.leaf_msg, d.leaf_len, cur);
            if (which == 4) {
                put_bytes(H2_MB, row, d.leaf_msg, 73);
                put_bytes(H2_DB, row, cur, 32);
                f.c(H2_EL, row) = 1;
                f.c(H2_ER, row) = 1;
                f.c(H2_FF, row) = 1;
            } else {
                if (d.leaf_len > 55) LOGIC_FAIL(17);
                put_bytes(H1_D, row, d.leaf_msg, d.leaf_len < 55 ? 55 : d.leaf_len);
                f.c(H1_IL + d.leaf_len, row) = 1;
                put_bytes(H1_DB, row, cur, 32);
                f.c(H1_E, row) = 1;
                if (which == 3) {
                    uint8_t vb[9];
                    put_varint(row, h->height_proof.height, vb);
                }
            }
            row++;
            for (int j = 1; j <= 4; j++, row++) {
                const uint8_t* aunt = d.aunts[j - 1];
                const bool right = (d.index >> (j - 1)) & 1;
                uint8_t msg[65];
                msg[0] = 1;
                memcpy(msg + 1, right ? aunt : cur, 32);
                memcpy(msg + 33, right ? cur : aunt, 32);
                sha256_bytes(msg, 65, cur);
                put_bytes(H2_MB, row, msg, 65);
                put_bytes(H2_DB, row, cur, 32);
                f.c(H2_EL, row) = 1;
                f.c(H2_ER, row) = 1;
                f.c(H2_FF, row) = 1;
            }
        }
    }
    // ---- validator slots
    const std::vector<Gadget>& nl = netlist();
    for (uint32_t i = 0; i < N; i++) {
        size_t row = p.row_slots + (size_t)i * (4 + LG_MUL_ROWS_PER_SLOT);
        const EdSlotInfo& e = slots[i];
        EdTriple t;
        effective_triple(vals + i, &t);
        const bool sgn = vals[i].is_signed != 0;
        // SIG
        {
            put_bytes(SG_D, row, t.sig, 32);
            put_bytes(SG_D + 32, row, t.pk, 32);
            put_bytes(SG_D + 64, row, t.msg, TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX);
            f.c(SG_IL + t.len, row) = 1;
            put_bytes(SG_DG, row, e.digest, 64);
            f.c(SG_A31, row) = t.pk[31] & 0x7F;
            f.c(SG_A31X2, row) = 2 * (t.pk[31] & 0x7F);
            f.c(SG_SA, row) = t.pk[31] >> 7;
            f.c(SG_R31, row) = t.sig[31] & 0x7F;
            f.c(SG_R31X2, row) = 2 * (t.sig[31] & 0x7F);
            f.c(SG_SR, row) = t.sig[31] >> 7;
            f.c(SG_SGN, row) = sgn;
            f.c(SG_RZ, row) = sgn && h->round == 0;
        }
        row++;
        // SC: digest = q l + h, s < l, h < l
        {
            put_bytes(SC_S, row, t.sig + 32, 32);
            put_bytes(SC_DG, row, e.digest, 64);
            int32_t hl[16], sl[16];
            for (int k = 0; k < 16; k++) {
                hl[k] = (int32_t)((e.h[k >> 2] >> (16 * (k & 3))) & 0xFFFF);
                sl[k] = t.sig[32 + 2 * k] | (t.sig[32 + 2 * k + 1] << 8);
                f.c(SC_HL + k, row) = (gl)hl[k];
            }
            const int64_t linv = 0x7d1b;  // 0xd3ed * 0x7d1b = 1 (mod 2^16)? computed below if not
            int64_t inv = 1;
            for (int it = 0; it < 5; it++) inv = (inv * (2 - (int64_t)ell_limb(0) * inv)) & 0xFFFF;  // Newton iteration mod 2^16
            (void)linv;
            int64_t q[17] = {0}, carry = 0;
            for (int k = 0; k < 32; k++) {
                int64_t s = (int64_t)(e.digest[2 * k] | (e.digest[2 * k + 1] << 8)) - (k < 16 ? hl[k] : 0) + carry;
                for (int a = 0; a < 17 && a < k; a++)
                    if (k - a < 16) s -= q[a] * (int64_t)ell_limb(k - a);
                if (k < 17) {
                    q[k] = ((s & 0xFFFF) * inv) & 0xFFFF;
                    s -= q[k] * (int64_t)ell_limb(0);
                    f.c(SC_Q + k, row) = (gl)q[k];
                }
                if (s & 0xFFFF) LOGIC_FAIL(5);
                carry = s >> 16;
                if ((k & 1) && k < 31) {
                    const int64_t w = carry + SC_W_OFFSET;
                    if (w < 0 || w >= (1 << 24)) LOGIC_FAIL(5);
                    f.c(SC_WLO + (k >> 1), row) = (gl)(w & 0xFFFF);
                    f.c(SC_WHI + (k >> 1), row) = (gl)(w >> 16);
                }
            }
            if (carry) LOGIC_FAIL(5);
            if (!sc_lt_l(e.s)) LOGIC_FAIL(5);
            less_than_cells(sl, ell_limb, &f.c(SC_DS, row), &f.c(SC_CS, row), n);
            less_than_cells(hl, ell_limb, &f.c(SC_DH, row), &f.c(SC_CH, row), n);
        }
        row++;
        // wires of the slot
        int32_t wire[64][16];
        memset(wire, 0, sizeof wire);
        limbs_of(e.yA, wire[W_YA]); limbs_of(e.xA, wire[W_XA]); limbs_of(e.yR, wire[W_YR]); limbs_of(e.xR, wire[W_XR]);
        limbs_of(e.xD, wire[W_XD]); limbs_of(e.yD, wire[W_YD]);
        limbs_of(e.QX, wire[W_XQ]); limbs_of(e.QY, wire[W_YQ]); limbs_of(e.QZ, wire[W_ZQ]);
        // XS
        for (int a = 0; a < 2; a++) {
            const int o = a * XS_STRIDE;
            const int32_t* x = wire[a ? W_XR : W_XA];
            const int32_t* y = wire[a ? W_YR : W_YA];
            put_limbs(o + XS_X, row, x);
            put_limbs(o + XS_Y, row, y);
            less_than_cells(x, p25519_limb, &f.c(o + XS_DX, row), &f.c(o + XS_CX, row), n);
            less_than_cells(y, p25519_limb, &f.c(o + XS_DY, row), &f.c(o + XS_CY, row), n);
            f.c(o + XS_PAR, row) = x[0] & 1;
            f.c(XS_XH + 2 * a, row) = x[0] >> 1;
            f.c(XS_XH2 + 2 * a, row) = 2 * (x[0] >> 1);
            f.c(XS_SIGN + a, row) = (a ? t.sig[31] : t.pk[31]) >> 7;
        }
        row++;
        // MUL rows first (their outputs feed EDIO), the EDIO row sits before them
        const size_t row_edio = row;
        row++;
        for (size_t g0 = 0; g0 < nl.size(); g0 += 2, row++)
            for (int hh = 0; hh < 2; hh++) {
                const int o = hh * MU_STRIDE;
                int32_t U[16] = {0}, V[16] = {0}, W[16] = {0}, c[16];
                if (g0 + hh < nl.size()) {
                    const Gadget& g = nl[g0 + hh];
                    int sl[6];
                    const int ns = gadget_slots(g, sl);
                    for (int s = 0; s < ns; s++) {
                        int32_t cw[16];
                        const int32_t* v = wire[0];
                        if (sl[s] >= CW) {
                            for (int j = 0; j < 16; j++) cw[j] = (int32_t)logic_const_wire_limb(sl[s] - CW, j);
                            v = cw;
                        } else
                            v = wire[sl[s]];
                        put_limbs(o + MU_IN + 16 * s, row, v);
                    }
                    const Operand* ops[3] = {&g.U, &g.V, &g.W};
                    int32_t* outs[3] = {U, V, W};
                    for (int k = 0; k < 3; k++)
                        for (int j = 0; j < 16; j++) {
                            int64_t acc = (int64_t)ops[k]->kp * (int64_t)p25519_limb(j);
                            for (int tt = 0; tt < ops[k]->nt; tt++) {
                                const int w = ops[k]->t[tt].wire;
                                acc += (int64_t)ops[k]->t[tt].coef * (w >= CW ? (int64_t)logic_const_wire_limb(w - CW, j) : wire[w][j]);
                            }
                            outs[k][j] = (int32_t)acc;
                        }
                }
                for (int j = 0; j < 16; j++) {
                    f.c(o + MU_U + j, row) = U[j] >= 0 ? (gl)U[j] : GL_P - (gl)(-U[j]);
                    f.c(o + MU_V + j, row) = V[j] >= 0 ? (gl)V[j] : GL_P - (gl)(-V[j]);
                }
                mul_gadget_cells(U, V, &f.c(o + MU_C, row), n, c, W);
                if (g0 + hh < nl.size()) {
                    memcpy(wire[W_G0 + g0 + hh], c, sizeof c);
                    if (nl[g0 + hh].assert0)
                        for (int j = 0; j < 16; j++)
                            if (c[j]) LOGIC_FAIL(5);  // a curve equation or the final check [s]B - [h]A == R fails
                }
            }
        // EDIO
        {
            put_limbs(EI_YA, row_edio, wire[W_YA]);
            put_limbs(EI_XA, row_edio, wire[W_XA]);
            put_limbs(EI_T2A, row_edio, wire[W_G0 + 9]);
            put_limbs(EI_XD, row_edio, wire[W_XD]);
            put_limbs(EI_YD, row_edio, wire[W_YD]);
            put_limbs(EI_T2D, row_edio, wire[W_G0 + 19]);
            put_limbs(EI_XQ, row_edio, wire[W_XQ]);
            put_limbs(EI_YQ, row_edio, wire[W_YQ]);
            put_limbs(EI_ZQ, row_edio, wire[W_ZQ]);
            uint64_t cnt[4] = {0, 0, 0, 0};
            for (int j = 0; j < 256; j++) cnt[((e.s[j >> 6] >> (j & 63)) & 1) + 2 * ((e.h[j >> 6] >> (j & 63)) & 1)]++;
            for (int k = 0; k < 4; k++) f.c(EI_M + k, row_edio) = cnt[k];
        }
    }
    return status;
#undef LOGIC_FAIL
}

// slot infos on the host (tests and the oracle's input; the prover takes them from the GPU)
void logic_slot_infos_host(const uint8_t* blob, uint32_t n_max, std::vector<EdSlotInfo>& out) {
    out.assign(n_max, EdSlotInfo());
    std::vector<ge_acc_packed> acc(ED_ROWS_PER_VALIDATOR);
    Sha512Hist h5[2];
    for (uint32_t i = 0; i < n_max; i++) {
        EdTriple t;
        effective_triple(blob_validators(blob) + i, &t);
        uint8_t digest[64];
        sha512_validator_prepare(t, h5, digest);
        ge_cached51 tab[4];
        ge51 R;
        ed_slot_prepare(t, digest, &out[i], tab, &R);
        const ge_acc51 q = ed_straus_ladder(out[i].s, out[i].h, tab, acc.data());
        out[i].QX = fe_freeze(q.X); out[i].QY = fe_freeze(q.Y); out[i].QZ = fe_freeze(q.Z);
    }
}

}  // namespace tmx

using namespace tmx;

// The logic table of one proof computed entirely on the host (tests, and the first-round trace the CPU oracle proves with):
// returns the number of u64 cells ([cols][rows], column-major), copies at most `cap` of them.  *status = 0 or a check id; with
// `force` the table is filled past the first failing check (the tests' cheating provers commit to such tables).
extern "C" size_t tmx_logic_trace(uint32_t kind, uint32_t n_max, const char* chain_id, size_t chain_id_len, const uint8_t* input,
                                  const uint8_t* blob, int force, uint64_t* out, size_t cap, int* status) {
    if (!chain_id || !input || !blob || kind > 1 || n_max == 0 || n_max > 4096 || chain_id_len == 0 || chain_id_len > 50) return 0;
    const AirShape sh = air_shape(kind, n_max, chain_id, chain_id_len);
    if (!logic_rows(sh)) return 0;
    auto plan = logic_plan_get(sh);
    const size_t cells = (size_t)LG_COLS * plan->n_rows;
    if (!out) return cells;
    std::vector<EdSlotInfo> slots;
    logic_slot_infos_host(blob, n_max, slots);
    std::vector<gl> tr(cells, 0);
    const int rc = logic_fill_trace(*plan, input, blob, slots.data(), tr.data(), force != 0);
    if (status) *status = rc;
    memcpy(out, tr.data(), std::min(cap, cells) * sizeof(uint64_t));
    return cells;
}
