// Host-side verifier (the role of `circuit.verify()` [REF circuits/skip.rs:247, circuits/step.rs:226] for our
// multi-table STARK with a shared bus): transcript replay over both commitment rounds, the bus balance (the per-table
// totals and the verifier's own public-input terms sum to zero), per table the constraint identity at zeta with the
// shared AIR templates over the extension field (table constraints, helper-column constraints, running sum), the
// proof-of-work check, and per query the Merkle openings of the constant / first-round / second-round / quotient
// trees, the FRI batch combination, the arity-16 folds and the final polynomial.  CPU by nature (it is what a light
// client or the next recursion layer runs); independent of the GPU prover's code paths.
#include "stark.cuh"
#include "bus.cuh"
#include "circuit_def.cuh"
#include "logic.cuh"
#include <algorithm>
#include <memory>
#include <mutex>

namespace tmx {

namespace {

struct Reader {
    const gl* v;
    size_t n, pos;
    bool err;
    gl get() {
        if (pos >= n) { err = true; return 0; }
        return v[pos++];
    }
    const gl* take(size_t k) {
        if (pos + k > n) { err = true; return nullptr; }
        const gl* p = v + pos;
        pos += k;
        return p;
    }
    gl2 ext() {
        gl a = get(), b = get();
        return gl2_make(a, b);
    }
};

struct ExtRow {
    const FE* p;
    FE operator[](int c) const { return p[c]; }
};

void hash_row_host(const gl* row, size_t n, gl out[4]) { poseidon_hash_row(row, 1, n, out); }

bool merkle_check(const gl* leaf, size_t leaf_len, size_t index, const gl* sib, unsigned n_sib, const gl* cap) {
    gl cur[4], nxt[4];
    hash_row_host(leaf, leaf_len, cur);
    for (unsigned k = 0; k < n_sib; k++) {
        if (index & 1) poseidon_two_to_one(sib + 4 * k, cur, nxt);
        else poseidon_two_to_one(cur, sib + 4 * k, nxt);
        for (int i = 0; i < 4; i++) cur[i] = nxt[i];
        index >>= 1;
    }
    for (int i = 0; i < 4; i++)
        if (cur[i] != cap[4 * index + i]) return false;
    return true;
}

gl2 ext_horner(const gl2* c, size_t n, gl2 x) {
    gl2 acc = gl2_from(0);
    for (size_t i = n; i-- > 0;) acc = gl2_add(gl2_mul(acc, x), c[i]);
    return acc;
}

// Lagrange interpolation through the 16 coset points, evaluated at beta (plonky2 compute_evaluation)
gl2 fold_coset(gl x, unsigned within, const gl2 evals[16], gl2 beta) {
    const gl g = gl_root_of_unity(STARK_ARITY_BITS);
    gl2 ys[16];
    gl xs[16];
    for (unsigned i = 0; i < 16; i++) ys[bitrev32(i, 4)] = evals[i];
    gl cur = gl_mul(x, gl_pow(g, 16 - bitrev32(within, 4)));
    for (int i = 0; i < 16; i++) {
        xs[i] = cur;
        cur = gl_mul(cur, g);
    }
    gl2 acc = gl2_from(0);
    for (int i = 0; i < 16; i++) {
        gl2 num = gl2_from(1);
        gl den = 1;
        for (int j = 0; j < 16; j++)
            if (j != i) {
                num = gl2_mul(num, gl2_sub(beta, gl2_from(xs[j])));
                den = gl_mul(den, gl_sub(xs[i], xs[j]));
            }
        acc = gl2_add(acc, gl2_mul(ys[i], gl2_scale(num, gl_inv(den))));
    }
    return acc;
}

}  // namespace

// one table's part of the proof after both commitment rounds: quotient commitment, openings, constraint identity, FRI
static int verify_table(const CircuitDef& def, int table, const gl* cap_m, const gl* cap_a, gl2 beta, gl2 gamma, gl2 total,
                        Reader& r, Challenger& ch) {
    const TableDef& td = def.tables[table];
    const AirShape shape = air_shape(def.kind, def.n_max, def.chain_id.data(), def.chain_id.size());
    const size_t n = td.rows(), m = n << STARK_RATE_BITS;
    const size_t Kc = td.n_const, C = td.n_main, A = (size_t)td.n_aux(), CT = Kc + C + A;
    const unsigned k = td.log_n, km = k + STARK_RATE_BITS;
    const unsigned cap_h = std::min<unsigned>(km, STARK_CAP_HEIGHT);
    const size_t cap_n = (size_t)1 << cap_h;
    gl alpha[2];
    alpha[0] = ch.get();
    alpha[1] = ch.get();
    const gl* cap_q = r.take(4 * cap_n);
    if (r.err) return 1;
    ch.observe(cap_q, 4 * cap_n);
    const gl2 zeta = ch.get_ext();
    const gl2 zeta_next = gl2_scale(zeta, gl_root_of_unity(k));
    std::vector<FE> loc(CT), nxt(CT);
    gl2 quot[4];
    for (size_t c = 0; c < CT; c++) loc[c] = FE::mk(r.ext());
    for (size_t c = 0; c < CT; c++) nxt[c] = FE::mk(r.ext());
    for (int q = 0; q < 4; q++) quot[q] = r.ext();
    if (r.err) return 1;
    for (size_t c = 0; c < CT; c++) ch.observe_ext(loc[c].v);
    for (int q = 0; q < 4; q++) ch.observe_ext(quot[q]);
    for (size_t c = 0; c < CT; c++) ch.observe_ext(nxt[c].v);
    // constraint identity at zeta: (chunk0 + zeta^n chunk1) * (zeta^n - 1) == sum_i alpha^(M-1-i) C_i(zeta)
    {
        const size_t P = td.period;
        FE per[AIR_MAX_PERIODIC];
        for (int i = 0; i < AIR_MAX_PERIODIC; i++) per[i] = FE::c(0);
        const gl2 y = gl2_pow(zeta, n / P);
        for (uint32_t pc = 0; pc < td.n_per; pc++) {
            // interpolant of the column's one-period pattern, composed with x -> x^(n / P)
            std::vector<gl> cb(td.periodic.begin() + (size_t)pc * P, td.periodic.begin() + (size_t)(pc + 1) * P);
            air_host_ntt(cb, true);
            std::vector<gl2> coef(P);
            for (size_t kk = 0; kk < P; kk++) coef[kk] = gl2_from(cb[kk]);
            per[pc] = FE::mk(ext_horner(coef.data(), P, y));
        }
        ConstraintAcc<FE> acc;
        acc.acc0 = FE::c(0); acc.acc1 = FE::c(0);
        acc.alpha0 = FE::mk(gl2_from(alpha[0])); acc.alpha1 = FE::mk(gl2_from(alpha[1]));
        ExtRow kl{loc.data()}, l{loc.data() + Kc}, nn{nxt.data() + Kc}, al{loc.data() + Kc + C}, an{nxt.data() + Kc + C}, pp{per};
        const Ext2<FE> eb = e2_mk<FE>(FE::c(beta.a0), FE::c(beta.a1)), eg = e2_mk<FE>(FE::c(gamma.a0), FE::c(gamma.a1));
        BusCheck<FE, ExtRow, ConstraintAcc<FE>> bus(eb, eg, al, acc);
        air_eval_any<FE>(table, shape, l, nn, kl, pp, acc, bus);
        if (bus.h != (int)td.n_helpers) return 8;
        const gl ninv = gl_inv((gl)n);
        bus.finish(an, e2_mk<FE>(FE::c(gl_mul(total.a0, ninv)), FE::c(gl_mul(total.a1, ninv))));
        const gl2 zn = gl2_pow(zeta, n), zh = gl2_sub(zn, gl2_from(1));
        for (int i = 0; i < 2; i++) {
            const gl2 q = gl2_add(quot[2 * i], gl2_mul(zn, quot[2 * i + 1]));
            if (!gl2_eq(gl2_mul(q, zh), (i == 0 ? acc.acc0 : acc.acc1).v)) return 2;
        }
    }
    const gl2 fa = ch.get_ext();
    gl2 red[2] = {gl2_from(0), gl2_from(0)};
    for (size_t j = CT + 4; j-- > 0;) red[0] = gl2_add(gl2_mul(red[0], fa), j < CT ? loc[j].v : quot[j - CT]);
    for (size_t j = CT; j-- > 0;) red[1] = gl2_add(gl2_mul(red[1], fa), nxt[j].v);
    const unsigned n_layers = fri_num_layers(k);
    std::vector<const gl*> layer_caps(n_layers);
    std::vector<gl2> betas(n_layers);
    size_t rows = m;
    for (unsigned l = 0; l < n_layers; l++) {
        rows >>= STARK_ARITY_BITS;
        const unsigned lg = ilog2(rows);
        const size_t lcap = (size_t)1 << std::min<unsigned>(lg, STARK_CAP_HEIGHT);
        layer_caps[l] = r.take(4 * lcap);
        if (r.err) return 1;
        ch.observe(layer_caps[l], 4 * lcap);
        betas[l] = ch.get_ext();
    }
    const size_t final_len = (size_t)r.get();
    if (r.err || final_len != ((m >> (STARK_ARITY_BITS * n_layers)) >> STARK_RATE_BITS) || final_len > 64) return 3;
    gl2 fin[64];
    for (size_t i = 0; i < final_len; i++) {
        fin[i] = r.ext();
        ch.observe_ext(fin[i]);
    }
    const gl pow_witness = r.get();
    if (r.err) return 1;
    ch.observe(pow_witness);
    if ((ch.get() >> (64 - STARK_POW_BITS)) != 0) return 4;
    const gl2 a_c = gl2_pow(fa, CT);
    const unsigned n_sib = km - cap_h;
    std::vector<gl> row(CT + 4);
    for (int qi = 0; qi < STARK_NUM_QUERIES; qi++) {
        size_t x = (size_t)(ch.get() % m);
        const gl* row_k = Kc ? r.take(Kc) : nullptr;
        const gl* path_k = Kc ? r.take(4 * n_sib) : nullptr;
        const gl* row_m = r.take(C);
        const gl* path_m = r.take(4 * n_sib);
        const gl* row_a = r.take(A);
        const gl* path_a = r.take(4 * n_sib);
        const gl* row_q = r.take(4);
        const gl* path_q = r.take(4 * n_sib);
        if (r.err) return 1;
        if (Kc && !merkle_check(row_k, Kc, x, path_k, n_sib, td.const_cap.data())) return 5;
        if (!merkle_check(row_m, C, x, path_m, n_sib, cap_m)) return 5;
        if (!merkle_check(row_a, A, x, path_a, n_sib, cap_a)) return 5;
        if (!merkle_check(row_q, 4, x, path_q, n_sib, cap_q)) return 5;
        for (size_t j = 0; j < Kc; j++) row[j] = row_k[j];
        for (size_t j = 0; j < C; j++) row[Kc + j] = row_m[j];
        for (size_t j = 0; j < A; j++) row[Kc + C + j] = row_a[j];
        for (size_t j = 0; j < 4; j++) row[CT + j] = row_q[j];
        gl sx = gl_mul(GL_GEN, gl_pow(gl_root_of_unity(km), bitrev32((uint32_t)x, km)));
        gl2 s0 = gl2_from(0), s1 = gl2_from(0);
        for (size_t j = CT + 4; j-- > 0;) s0 = gl2_add(gl2_mul(s0, fa), gl2_from(row[j]));
        for (size_t j = CT; j-- > 0;) s1 = gl2_add(gl2_mul(s1, fa), gl2_from(row[j]));
        gl2 sum = gl2_mul(gl2_sub(s0, red[0]), gl2_inv(gl2_sub(gl2_from(sx), zeta)));
        sum = gl2_add(gl2_mul(sum, a_c), gl2_mul(gl2_sub(s1, red[1]), gl2_inv(gl2_sub(gl2_from(sx), zeta_next))));
        gl2 old = sum;
        size_t lrows = m;
        for (unsigned l = 0; l < n_layers; l++) {
            lrows >>= STARK_ARITY_BITS;
            const unsigned lg = ilog2(lrows);
            const unsigned lsib = lg - std::min<unsigned>(lg, STARK_CAP_HEIGHT);
            const gl* leaf = r.take(32);
            const gl* path = r.take(4 * lsib);
            if (r.err) return 1;
            gl2 ev[16];
            for (int i = 0; i < 16; i++) ev[i] = gl2_make(leaf[2 * i], leaf[2 * i + 1]);
            const unsigned within = (unsigned)(x & 15);
            const size_t coset = x >> STARK_ARITY_BITS;
            if (!gl2_eq(ev[within], old)) return 6;
            old = fold_coset(sx, within, ev, betas[l]);
            if (!merkle_check(leaf, 32, coset, path, lsib, layer_caps[l])) return 5;
            sx = gl_pow(sx, 16);
            x = coset;
        }
        if (!gl2_eq(ext_horner(fin, final_len, gl2_from(sx)), old)) return 7;
    }
    return 0;
}

void transcript_init(const gl digest[4], const uint8_t* input, size_t input_len, const uint8_t out32[32], Challenger& ch) {
    ch.observe(digest, 4);
    gl pub[128], s[12] = {0};
    size_t k = 0;
    for (size_t i = 0; i < input_len && k < 96; i++) pub[k++] = input[i];
    for (size_t i = 0; i < 32; i++) pub[k++] = out32[i];
    for (size_t off = 0; off < k; off += 8) {  // hash_no_pad
        const size_t cnt = std::min<size_t>(8, k - off);
        for (size_t i = 0; i < cnt; i++) s[i] = pub[off + i];
        poseidon_permute(s);
    }
    ch.observe(s, 4);
}

// Returns 0 when the proof verifies; otherwise 100.. for header problems, 10 * (table + 1) + code for a table, 200 for an
// unbalanced bus.
int verify_proof(const CircuitDef& def, const gl* w, size_t n_words, const uint8_t* input, size_t input_len, const uint8_t out32[32]) {
    if (input_len != (def.kind == TMX_KIND_SKIP ? 48u : 40u)) return 102;
    Reader r{w, n_words, 0, false};
    if (r.get() != STARK_PROOF_MAGIC || r.get() != def.kind || r.get() != def.n_max || r.get() != TMX_N_TABLES) return 100;
    for (int i = 0; i < 4; i++) {
        gl x = 0;
        for (int j = 0; j < 8; j++) x |= (gl)out32[8 * i + j] << (8 * j);
        if (r.get() != x) return 101;
    }
    if (r.err) return 100;
    // malleability: every word of a proof is a canonical field element
    for (size_t i = 0; i < n_words; i++)
        if (w[i] >= GL_P) return 104;
    poseidon_generate_constants();
    Challenger ch;
    transcript_init(def.digest, input, input_len, out32, ch);
    const gl* cap_m[TMX_N_TABLES] = {nullptr};
    const gl* cap_a[TMX_N_TABLES] = {nullptr};
    gl2 total[TMX_N_TABLES];
    auto cap_words = [&](int t) { return 4 * ((size_t)1 << std::min<unsigned>(def.tables[t].log_n + STARK_RATE_BITS, STARK_CAP_HEIGHT)); };
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!def.tables[t].n_main) continue;
        cap_m[t] = r.take(cap_words(t));
        if (r.err) return 100;
        ch.observe(cap_m[t], cap_words(t));
    }
    const gl2 beta = ch.get_ext(), gamma = ch.get_ext();
    gl2 balance = logic_public_terms(air_shape(def.kind, def.n_max, def.chain_id.data(), def.chain_id.size()), def.skip_max, input, out32, beta, gamma);
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!def.tables[t].n_main) continue;
        cap_a[t] = r.take(cap_words(t));
        total[t] = r.ext();
        if (r.err) return 100;
        ch.observe(cap_a[t], cap_words(t));
        ch.observe_ext(total[t]);
        balance = gl2_add(balance, total[t]);
    }
    if (balance.a0 != 0 || balance.a1 != 0) return 200;
    for (int t = 0; t < TMX_N_TABLES; t++) {
        if (!def.tables[t].n_main) continue;
        Challenger fork = ch;  // every table continues on its own fork of the transcript: common state + table index
        fork.observe((gl)t);
        const int rc = verify_table(def, t, cap_m[t], cap_a[t], beta, gamma, total[t], r, fork);
        if (rc) return 10 * (t + 1) + rc;
    }
    if (r.err || r.pos != n_words) return 103;
    return 0;
}

}  // namespace tmx
