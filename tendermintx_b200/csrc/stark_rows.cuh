// Per-thread bodies of the quotient kernel (K5) and of the trace-domain bus passes, host+device: the kernels in stark.cu
// map one LDE point / one trace row to a thread and call these; tools/hostsim.cpp runs the same code in a loop on the CPU so
// that tests can compare it with the oracle without a GPU.
#pragma once
#include "bus.cuh"
#include "logic.cuh"

namespace tmx {

// one row of a column-major matrix: cell c at base[c * stride]
struct ColRow {
    const gl* base;
    size_t stride;
    TMX_HD FB operator[](int c) const { return FB::mk(base[(size_t)c * stride]); }
};

// Position (bit-reversed order) of the row that follows position p on the trace domain: natural index j + 2^r.
// The low log_n bits of p hold the bit-reversed row counter, so "+1" is a reverse-carry increment (flip ones from
// the top bit down, set the first zero); the top r bits (the coset id) are unchanged.
TMX_HD size_t next_row_position(size_t p, unsigned log_n) {
    const size_t low_mask = ((size_t)1 << log_n) - 1;
    size_t q = p & low_mask;
    size_t bit = (size_t)1 << (log_n - 1);
    while (bit && (q & bit)) {
        q ^= bit;
        bit >>= 1;
    }
    q |= bit;
    return (p & ~low_mask) | q;
}


// ------------------------------------------------------------------------------------------ K5: quotient
struct QuotientArgs {
    const gl* lde_m;   // [C][m] first-round columns, bit-reversed rows
    const gl* lde_k;   // [Kc][m] constant columns
    const gl* lde_a;   // [A][m] second-round columns
    size_t m;
    unsigned log_m, rate_bits;
    const gl* pertab;  // [nper][2P]
    int P;
    AirShape shape;
    gl alpha[2];
    gl2 beta, gamma, s_over_n;
    gl zh_inv[1 << 3];  // indexed by natural index mod 2^rate_bits
    gl* out;            // [2][m] natural order
};

template <int TABLE>
TMX_HD void quotient_point(const QuotientArgs& a, size_t p) {
    const uint32_t j = bitrev32((uint32_t)p, a.log_m);
    const size_t pn = next_row_position(p, a.log_m - a.rate_bits);
    const ColRow l{a.lde_m + p, a.m}, n{a.lde_m + pn, a.m}, k{a.lde_k + p, a.m}, al{a.lde_a + p, a.m}, an{a.lde_a + pn, a.m};
    const ColRow per{a.pertab + (j & (2 * a.P - 1)), (size_t)2 * a.P};
    ConstraintAcc<FB> acc;
    acc.acc0 = FB::c(0); acc.acc1 = FB::c(0);
    acc.alpha0 = FB::mk(a.alpha[0]); acc.alpha1 = FB::mk(a.alpha[1]);
    const Ext2<FB> eb = e2_mk<FB>(FB::mk(a.beta.a0), FB::mk(a.beta.a1)), eg = e2_mk<FB>(FB::mk(a.gamma.a0), FB::mk(a.gamma.a1));
    BusCheck<FB, ColRow, ConstraintAcc<FB>> bus(eb, eg, al, acc);
    air_eval_any<FB>(TABLE, a.shape, l, n, k, per, acc, bus);
    bus.finish(an, e2_mk<FB>(FB::mk(a.s_over_n.a0), FB::mk(a.s_over_n.a1)));
    const gl zi = a.zh_inv[j & ((1u << a.rate_bits) - 1)];
    a.out[j] = gl_mul(acc.acc0.v, zi);
    a.out[a.m + j] = gl_mul(acc.acc1.v, zi);
}

// ------------------------------------------------------------------------------------------ bus: trace-domain passes
struct BusPassArgs {
    const gl* trace;   // [C][n]
    const gl* kconst;  // [Kc][n]
    const gl* per;     // [nper][P]
    size_t n;
    int P;
    AirShape shape;
    gl2 beta, gamma;
    gl* aux;           // [A][n]
    gl2* rowsum;       // [n]
    unsigned int* hist;
};
// helper columns of the table's bus interactions, one thread per trace row
template <int TABLE>
TMX_HD void bus_gen_row(const BusPassArgs& a, size_t r) {
    const size_t rn = r + 1 == a.n ? 0 : r + 1;
    const ColRow l{a.trace + r, a.n}, n{a.trace + rn, a.n}, k{a.kconst + r, a.n}, per{a.per + (r % a.P), (size_t)a.P};
    NullEmit emit;
    BusGen bus(a.beta, a.gamma, a.aux + r, a.n);
    air_eval_any<FB>(TABLE, a.shape, l, n, k, per, emit, bus);
    bus.flush();
    a.rowsum[r] = bus.sum;
}
// histogram of the table's range lookups
template <int TABLE>
TMX_HD void bus_count_row(const BusPassArgs& a, size_t r) {
    const size_t rn = r + 1 == a.n ? 0 : r + 1;
    const ColRow l{a.trace + r, a.n}, n{a.trace + rn, a.n}, k{a.kconst + r, a.n}, per{a.per + (r % a.P), (size_t)a.P};
    NullEmit emit;
    BusCount bus{a.hist, (int*)(a.hist + BUS_HIST_SIZE)};
    air_eval_any<FB>(TABLE, a.shape, l, n, k, per, emit, bus);
}
}  // namespace tmx
