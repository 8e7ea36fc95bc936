// Bus (logUp) callbacks plugged into the AIR templates of air.cuh.  A table's AIR declares its interactions through
// bus.one / bus.two; what happens then depends on the pass:
//   BusCheck<F>  constraint side (quotient kernels with F = FB on the LDE coset, verifier with F = FE at zeta): every
//                declaration owns one helper column H (extension valued, two committed columns) with
//                    one:  H f = m                       two:  H fa fb = ma fb + mb fa
//                and the running sum Z of the second-round trace satisfies, on EVERY row (cyclically),
//                    Z(g x) - Z(x) - sum_k H_k(x) + S / n = 0,
//                so S is the table's total bus contribution without any first / last row selector.
//   BusGen       trace-domain pass of the prover: computes the helper values of one row.
//   BusCount     trace-domain pass of the prover: histogram of the range-checked values (multiplicity columns of the
//                range table).
// Upstream the same role is played by Curta's lookup / bus arguments behind `curta_eddsa_verify_sigs_conditional` and
// `curta_sha256_variable` [REF circuits/builder/verify.rs:202,248-259] and by plonky2's copy constraints.
#pragma once
#include "air.cuh"

namespace tmx {

struct NullEmit {
    template <class F>
    TMX_HD void operator()(F) const {}
};

template <class F, class AuxRow, class Emit>
struct BusCheck {
    Ext2<F> beta, gamma;
    const AuxRow& al;
    Emit& emit;
    int h;
    Ext2<F> sum;
    TMX_HD BusCheck(Ext2<F> b, Ext2<F> g, const AuxRow& a, Emit& e) : beta(b), gamma(g), al(a), emit(e), h(0) {
        sum = e2_mk<F>(F::c(0), F::c(0));
    }
    TMX_HD Ext2<F> helper() {
        const Ext2<F> H = e2_mk<F>(al[2 * h], al[2 * h + 1]);
        h++;
        sum = e2_add<F>(sum, H);
        return H;
    }
    template <class Tup>
    TMX_HD void one(F tag, F m, int len, const Tup& tup) {
        const Ext2<F> f = bus_fingerprint<F>(beta, gamma, tag, len, tup);
        Ext2<F> c = e2_mul<F>(helper(), f);
        c.a0 = c.a0 - m;
        emit(c.a0);
        emit(c.a1);
    }
    template <class TA, class TB>
    TMX_HD void two(F tag_a, F ma, int len_a, const TA& ta, F tag_b, F mb, int len_b, const TB& tb) {
        const Ext2<F> fa = bus_fingerprint<F>(beta, gamma, tag_a, len_a, ta);
        const Ext2<F> fb = bus_fingerprint<F>(beta, gamma, tag_b, len_b, tb);
        const Ext2<F> c = e2_sub<F>(e2_mul<F>(helper(), e2_mul<F>(fa, fb)), e2_add<F>(e2_scale<F>(fb, ma), e2_scale<F>(fa, mb)));
        emit(c.a0);
        emit(c.a1);
    }
    // two lookups (multiplicity -1 each) of single values:  H fa fb = -(fa + fb)
    TMX_HD void two_lookups(F tag_a, F va, F tag_b, F vb) {
        Ext2<F> fa = e2_add<F>(e2_scale<F>(beta, va), gamma), fb = e2_add<F>(e2_scale<F>(beta, vb), gamma);
        fa.a0 = fa.a0 + tag_a;
        fb.a0 = fb.a0 + tag_b;
        const Ext2<F> c = e2_add<F>(e2_mul<F>(helper(), e2_mul<F>(fa, fb)), e2_add<F>(fa, fb));
        emit(c.a0);
        emit(c.a1);
    }
    // an: the next row of the second-round trace; s_over_n: the table's claimed total divided by its length
    TMX_HD void finish(const AuxRow& an, Ext2<F> s_over_n) {
        const Ext2<F> z = e2_mk<F>(al[2 * h], al[2 * h + 1]), zn = e2_mk<F>(an[2 * h], an[2 * h + 1]);
        const Ext2<F> c = e2_add<F>(e2_sub<F>(e2_sub<F>(zn, z), sum), s_over_n);
        emit(c.a0);
        emit(c.a1);
    }
};

// helper values of one trace row, written to out[2 * h], out[2 * h + 1] with stride `stride`; the row's sum is kept.
// H = numerator / denominator in the extension field: the denominators of up to BATCH consecutive helpers are inverted together
// (Montgomery's trick: one field inversion, i.e. ~100 multiplications, per batch instead of per helper; same field elements).
struct BusGen {
    static constexpr int BATCH = 8;
    gl2 beta, gamma;
    gl* out;
    size_t stride;
    int h, pending;
    gl2 sum;
    gl2 num[BATCH], den[BATCH];
    int slot[BATCH];
    TMX_HD BusGen(gl2 b, gl2 g, gl* o, size_t s) : beta(b), gamma(g), out(o), stride(s), h(0), pending(0) { sum = gl2_from(0); }
    TMX_HD void store(int at, gl2 H) {
        out[(size_t)(2 * at) * stride] = H.a0;
        out[(size_t)(2 * at + 1) * stride] = H.a1;
        sum = gl2_add(sum, H);
    }
    TMX_HD void flush() {
        if (!pending) return;
        gl2 pre[BATCH];
        gl2 acc = den[0];
        pre[0] = acc;
        for (int i = 1; i < pending; i++) {
            acc = gl2_mul(acc, den[i]);
            pre[i] = acc;
        }
        gl2 inv = gl2_inv(acc);
        for (int i = pending - 1; i >= 0; i--) {
            const gl2 di = i ? gl2_mul(inv, pre[i - 1]) : inv;  // 1 / den[i]
            store(slot[i], gl2_mul(num[i], di));
            inv = gl2_mul(inv, den[i]);
        }
        pending = 0;
    }
    TMX_HD void put(gl2 n, gl2 d) {
        num[pending] = n;
        den[pending] = d;
        slot[pending] = h++;
        if (++pending == BATCH) flush();
    }
    TMX_HD void put_zero() { store(h++, gl2_from(0)); }
    template <class Tup>
    TMX_HD gl2 fp(FB tag, int len, const Tup& tup) const {
        gl2 acc = gl2_scale(beta, tup(len - 1).v);
        for (int i = len - 2; i >= 0; i--) {
            acc.a0 = gl_add(acc.a0, tup(i).v);
            acc = gl2_mul(acc, beta);
        }
        acc.a0 = gl_add(acc.a0, tag.v);
        return gl2_add(acc, gamma);
    }
    TMX_HD void two_lookups(FB tag_a, FB va, FB tag_b, FB vb) {
        gl2 fa = gl2_add(gl2_scale(beta, va.v), gamma), fb = gl2_add(gl2_scale(beta, vb.v), gamma);
        fa.a0 = gl_add(fa.a0, tag_a.v);
        fb.a0 = gl_add(fb.a0, tag_b.v);
        put(gl2_neg(gl2_add(fa, fb)), gl2_mul(fa, fb));
    }
    template <class Tup>
    TMX_HD void one(FB tag, FB m, int len, const Tup& tup) {
        if (m.v == 0) { put_zero(); return; }
        put(gl2_from(m.v), fp(tag, len, tup));
    }
    template <class TA, class TB>
    TMX_HD void two(FB tag_a, FB ma, int len_a, const TA& ta, FB tag_b, FB mb, int len_b, const TB& tb) {
        if (ma.v == 0 && mb.v == 0) { put_zero(); return; }
        const gl2 fa = fp(tag_a, len_a, ta), fb = fp(tag_b, len_b, tb);
        put(gl2_add(gl2_scale(fb, ma.v), gl2_scale(fa, mb.v)), gl2_mul(fa, fb));
    }
};

// histogram of the range lookups (multiplicity p - 1 = -1) of one trace row: hist[0 .. 2^16) 16-bit, then 2^11, 2^8, 2
struct BusCount {
    unsigned int* hist;
    int* bad;  // set when a looked-up value is outside its table (the witness cannot be proved)
    TMX_HD void add(FB tagf, FB m, gl v) {
        if (m.v != GL_P - 1) return;
        const gl tag = tagf.v;
        size_t base, lim;
        if (tag == BUS_R16) { base = 0; lim = 1u << 16; }
        else if (tag == BUS_R11) { base = 1u << 16; lim = 1u << 11; }
        else if (tag == BUS_R8) { base = (1u << 16) + (1u << 11); lim = 1u << 8; }
        else if (tag == BUS_R1) { base = (1u << 16) + (1u << 11) + (1u << 8); lim = 2; }
        else return;
        if (v >= lim) { *bad = 1; return; }
#if defined(__CUDA_ARCH__)
        atomicAdd(hist + base + v, 1u);
#else
        hist[base + v]++;
#endif
    }
    template <class Tup>
    TMX_HD void one(FB tag, FB m, int, const Tup& tup) { add(tag, m, tup(0).v); }
    template <class TA, class TB>
    TMX_HD void two(FB tag_a, FB ma, int, const TA& ta, FB tag_b, FB mb, int, const TB& tb) {
        add(tag_a, ma, ta(0).v);
        add(tag_b, mb, tb(0).v);
    }
    TMX_HD void two_lookups(FB tag_a, FB va, FB tag_b, FB vb) {
        add(tag_a, FB::mk(GL_P - 1), va.v);
        add(tag_b, FB::mk(GL_P - 1), vb.v);
    }
};
constexpr size_t BUS_HIST_SIZE = (1u << 16) + (1u << 11) + (1u << 8) + 2;

}  // namespace tmx
