// Host-side verifier of one table proof (the role of `circuit.verify()` [REF circuits/skip.rs:247,
// circuits/step.rs:226] for our three-table STARK): transcript replay, constraint identity at zeta with the
// shared AIR templates over the extension field, proof-of-work check, and per query the Merkle openings,
// the FRI batch combination, the arity-16 folds and the final polynomial.  CPU by nature (it is what a light
// client or the next recursion layer runs); independent of the GPU prover's code paths.
#include "stark.cuh"
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

// interpolants of the periodic / public columns depend on (table, column, length, shape) only: computed once per process
static const std::vector<gl>& cached_periodic_coeffs(int table, int pc, size_t P, AirShape shape) {
    static std::mutex m;
    static std::map<std::vector<uint64_t>, std::shared_ptr<const std::vector<gl>>> cache;
    const std::vector<uint64_t> key = {(uint64_t)table, (uint64_t)pc, (uint64_t)P, shape.kind, shape.n_max};
    std::lock_guard<std::mutex> lk(m);
    auto hit = cache.find(key);
    if (hit == cache.end())
        hit = cache.emplace(key, std::make_shared<const std::vector<gl>>(air_periodic_coeffs(table, pc, P, h_K256, h_K512, shape))).first;
    return *hit->second;  // entries are never erased
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

int verify_table(int table, size_t n, AirShape shape, const gl* proof, size_t proof_len, size_t* pos, Challenger& ch) {
    Reader r{proof, proof_len, *pos, false};
    const size_t C = (size_t)air_cols(table), m = n << STARK_RATE_BITS;
    const unsigned k = ilog2(n), km = k + STARK_RATE_BITS;
    const unsigned cap_h = std::min<unsigned>(km, STARK_CAP_HEIGHT);
    const size_t cap_n = (size_t)1 << cap_h;
    const gl* cap_t = r.take(4 * cap_n);
    if (r.err) return 1;
    ch.observe(cap_t, 4 * cap_n);
    gl alpha[2] = {0, 0};
    alpha[0] = ch.get();
    alpha[1] = ch.get();
    const gl* cap_q = r.take(4 * cap_n);
    if (r.err) return 1;
    ch.observe(cap_q, 4 * cap_n);
    const gl2 zeta = ch.get_ext();
    const gl2 zeta_next = gl2_scale(zeta, gl_root_of_unity(k));
    std::vector<FE> loc(C), nxt(C);
    gl2 quot[4];
    for (size_t c = 0; c < C; c++) loc[c] = FE::mk(r.ext());
    for (size_t c = 0; c < C; c++) nxt[c] = FE::mk(r.ext());
    for (int q = 0; q < 4; q++) quot[q] = r.ext();
    if (r.err) return 1;
    for (size_t c = 0; c < C; c++) ch.observe_ext(loc[c].v);
    for (int q = 0; q < 4; q++) ch.observe_ext(quot[q]);
    for (size_t c = 0; c < C; c++) ch.observe_ext(nxt[c].v);
    // constraint identity at zeta: (chunk0 + zeta^n chunk1) * (zeta^n - 1) == sum_i alpha^(M-1-i) C_i(zeta)
    {
        const int nper = air_n_periodic(table);
        const size_t P = air_period(table, n);
        FE per[AIR_MAX_PERIODIC];
        for (int i = 0; i < AIR_MAX_PERIODIC; i++) per[i] = FE::c(0);
        const gl2 y = gl2_pow(zeta, n / P);
        for (int pc = 0; pc < nper; pc++) {
            // interpolant of the column's one-period pattern (for the SHA-256 table's public columns: of the whole column)
            const std::vector<gl>& cb = cached_periodic_coeffs(table, pc, P, shape);
            std::vector<gl2> coef(P);
            for (size_t kk = 0; kk < P; kk++) coef[kk] = gl2_from(cb[kk]);
            per[pc] = FE::mk(ext_horner(coef.data(), P, y));
        }
        ConstraintAcc<FE> acc;
        acc.acc0 = FE::c(0); acc.acc1 = FE::c(0);
        acc.alpha0 = FE::mk(gl2_from(alpha[0])); acc.alpha1 = FE::mk(gl2_from(alpha[1]));
        ExtRow l{loc.data()}, nn{nxt.data()}, pp{per};
        air_eval<FE>(table, l, nn, pp, acc);
        const gl2 zn = gl2_pow(zeta, n), zh = gl2_sub(zn, gl2_from(1));
        for (int i = 0; i < 2; i++) {
            const gl2 q = gl2_add(quot[2 * i], gl2_mul(zn, quot[2 * i + 1]));
            if (!gl2_eq(gl2_mul(q, zh), (i == 0 ? acc.acc0 : acc.acc1).v)) return 2;
        }
    }
    const gl2 fa = ch.get_ext();
    gl2 red[2] = {gl2_from(0), gl2_from(0)};
    for (size_t j = C + 4; j-- > 0;) red[0] = gl2_add(gl2_mul(red[0], fa), j < C ? loc[j].v : quot[j - C]);
    for (size_t j = C; j-- > 0;) red[1] = gl2_add(gl2_mul(red[1], fa), nxt[j].v);
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
    const gl2 a_c = gl2_pow(fa, C), a_c4 = gl2_pow(fa, C + 4);
    const unsigned n_sib = km - cap_h;
    for (int qi = 0; qi < STARK_NUM_QUERIES; qi++) {
        size_t x = (size_t)(ch.get() % m);
        const gl* row_t = r.take(C);
        const gl* path_t = r.take(4 * n_sib);
        const gl* row_q = r.take(4);
        const gl* path_q = r.take(4 * n_sib);
        if (r.err) return 1;
        if (!merkle_check(row_t, C, x, path_t, n_sib, cap_t)) return 5;
        if (!merkle_check(row_q, 4, x, path_q, n_sib, cap_q)) return 5;
        gl sx = gl_mul(GL_GEN, gl_pow(gl_root_of_unity(km), bitrev32((uint32_t)x, km)));
        gl2 s0 = gl2_from(0), s1 = gl2_from(0);
        for (size_t j = C + 4; j-- > 0;) s0 = gl2_add(gl2_mul(s0, fa), gl2_from(j < C ? row_t[j] : row_q[j - C]));
        for (size_t j = C; j-- > 0;) s1 = gl2_add(gl2_mul(s1, fa), gl2_from(row_t[j]));
        gl2 sum = gl2_mul(gl2_sub(s0, red[0]), gl2_inv(gl2_sub(gl2_from(sx), zeta)));  // first batch: 0 * alpha^(C+4) + term
        (void)a_c4;
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
    *pos = r.pos;
    return 0;
}

}  // namespace tmx
