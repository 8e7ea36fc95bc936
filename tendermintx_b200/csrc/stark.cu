// GPU STARK prover for one witness table: K5 (constraint quotient over the LDE), openings at zeta / g*zeta,
// K3 (FRI batch combination and arity-16 folding in EVALUATION space), K4 (query / opening gather), driven by
// a host-side duplex challenger.  Commitment uses K1 (tmx_lde) and K2 (Poseidon Merkle).
//
// Replaces, on the GPU, the plonky2 / starky proving loops behind `circuit.prove()`
// [REF circuits/skip.rs:214,244; circuits/step.rs:196,223]: PolynomialBatch::from_values, the quotient
// computation, fri/oracle.rs prove_openings, fri/prover.rs fri_committed_trees / fri_proof_of_work / query rounds.
// Same functions of the same field elements as the CPU oracle (oracle/stark.c), different algorithms: the batch
// polynomial is formed pointwise on the LDE instead of dividing coefficient vectors, and FRI layers are folded by
// a 16-point inverse NTT per coset instead of folding coefficients and re-running a coset FFT.
#include "stark.cuh"
#include "air_ed_fast.cuh"
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <algorithm>
#include <memory>
#include <mutex>

namespace tmx {

// ------------------------------------------------------------------------------------------ K5: quotient
struct LdeRow {
    const gl* base;  // lde + position
    size_t stride;   // m
    TMX_D FB operator[](int c) const { return FB::mk(base[(size_t)c * stride]); }
};
// periodic column values at one LDE position, read straight from the device table
struct LdePeriodic {
    const gl* base;  // pertab + (j mod 2P)
    size_t stride;   // 2P
    TMX_D FB operator[](int pc) const { return FB::mk(base[(size_t)pc * stride]); }
};

// Position (bit-reversed order) of the row that follows position p on the trace domain: natural index j + 2^r.
// The low log_n bits of p hold the bit-reversed row counter, so "+1" is a reverse-carry increment (flip ones from
// the top bit down, set the first zero); the top r bits (the coset id) are unchanged.
TMX_D size_t next_row_position(size_t p, unsigned log_n) {
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

struct QuotientArgs {
    const gl* lde;
    size_t m;
    unsigned log_m;
    unsigned rate_bits;
    const gl* pertab;  // [nper][2P]
    int nper, P;
    gl alpha[2];
    gl zh_inv[1 << 3];  // indexed by natural index mod 2^rate_bits
    gl* out;            // [2][m] natural order
};

template <int TABLE>
__global__ void __launch_bounds__(128) quotient_kernel(QuotientArgs a) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.m) return;
    const uint32_t j = bitrev32((uint32_t)p, a.log_m);
    const size_t pn = next_row_position(p, a.log_m - a.rate_bits);
    LdeRow l{a.lde + p, a.m}, n{a.lde + pn, a.m};
    LdePeriodic per{a.pertab + (j & (2 * a.P - 1)), (size_t)2 * a.P};
    ConstraintAcc<FB> acc;
    acc.acc0 = FB::c(0); acc.acc1 = FB::c(0);
    acc.alpha0 = FB::mk(a.alpha[0]); acc.alpha1 = FB::mk(a.alpha[1]);
    air_eval<FB>(TABLE, l, n, per, acc);
    const gl zi = a.zh_inv[j & ((1u << a.rate_bits) - 1)];
    a.out[j] = gl_mul(acc.acc0.v, zi);
    a.out[a.m + j] = gl_mul(acc.acc1.v, zi);
}

// Ed25519 table: factored evaluation (air_ed_fast.cuh), one thread per (LDE point, challenge); the two challenge
// halves of a CTA read the same cells, so the second read is an L1 hit.
struct QuotientEdArgs {
    const gl* lde;
    size_t m;
    unsigned log_m, rate_bits;
    const gl* pertab;  // [3][2P]: not_block_end, first row of [s]B, first row of [h]A
    int P;
    EdFastConsts k[2];
    gl zh_inv[1 << 3];
    gl* out;
};
__global__ void __launch_bounds__(128) quotient_ed25519_kernel(QuotientEdArgs a) {
    const int which = threadIdx.x >> 6;
    const size_t p = (size_t)blockIdx.x * 64 + (threadIdx.x & 63);
    if (p >= a.m) return;
    const uint32_t j = bitrev32((uint32_t)p, a.log_m);
    const size_t pn = next_row_position(p, a.log_m - a.rate_bits);
    LdeRow l{a.lde + p, a.m}, n{a.lde + pn, a.m};
    const uint32_t jp = j & (2 * a.P - 1);
    const gl per[3] = {a.pertab[jp], a.pertab[2 * a.P + jp], a.pertab[4 * a.P + jp]};
    const gl v = ed25519_constraints_fast(l, n, per, a.k[which]);
    a.out[(size_t)which * a.m + j] = gl_mul(v, a.zh_inv[j & ((1u << a.rate_bits) - 1)]);
}

// after the inverse NTT of size m = 2n the buffer holds q_i * 7^i; chunk k of challenge c is coefficients
// [k n, (k+1) n), whose own coset-scaled form is that slice times 7^(-k n)
__global__ void quotient_chunks_kernel(const gl* __restrict__ qcoef_m, size_t n, gl g_inv_n, gl* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * n) return;
    const size_t poly = i / n, idx = i % n;       // poly = 2 * challenge + chunk
    const size_t c = poly >> 1, k = poly & 1;
    gl v = qcoef_m[c * 2 * n + k * n + idx];
    if (k) v = gl_mul(v, g_inv_n);
    out[i] = v;
}

// ------------------------------------------------------------------------------------------ openings
__global__ void ext_powers_kernel(gl2 y, size_t n, gl2* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gl2_pow(y, i);
}

// out[col] = sum_i coeffs[col][i] * ypow[i]  (one CTA per column; exact field sums, order irrelevant)
__global__ void __launch_bounds__(256) eval_columns_kernel(const gl* __restrict__ coeffs, size_t n, const gl2* __restrict__ ypa,
                                                            const gl2* __restrict__ ypb, gl2* __restrict__ out_a,
                                                            gl2* __restrict__ out_b) {
    __shared__ gl2 red[2][256];
    const gl* c = coeffs + (size_t)blockIdx.x * n;
    // exact field sums, order irrelevant: 192-bit accumulators, one reduction per thread
    gl_acc192 a0 = gl_acc_zero(), a1 = gl_acc_zero(), b0 = gl_acc_zero(), b1 = gl_acc_zero();
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        const gl v = c[i];
        const gl2 ya = ypa[i];
        gl_acc_mac(a0, ya.a0, v);
        gl_acc_mac(a1, ya.a1, v);
        if (ypb) {
            const gl2 yb = ypb[i];
            gl_acc_mac(b0, yb.a0, v);
            gl_acc_mac(b1, yb.a1, v);
        }
    }
    const gl2 sa = gl2_make(gl_acc_reduce(a0), gl_acc_reduce(a1)), sb = gl2_make(gl_acc_reduce(b0), gl_acc_reduce(b1));
    red[0][threadIdx.x] = sa;
    red[1][threadIdx.x] = sb;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            red[0][threadIdx.x] = gl2_add(red[0][threadIdx.x], red[0][threadIdx.x + s]);
            red[1][threadIdx.x] = gl2_add(red[1][threadIdx.x], red[1][threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out_a[blockIdx.x] = red[0][0];
        if (ypb) out_b[blockIdx.x] = red[1][0];
    }
}

// ------------------------------------------------------------------------------------------ K3: FRI
struct FriBatchArgs {
    const gl* lde_t;  // [C][m]
    const gl* lde_q;  // [4][m]
    size_t C, m;
    unsigned log_m;
    const gl2* apow;  // alpha^j, j < C + 4
    gl2 red0, red1, zeta, zeta_next, alpha_c;
    gl w_m;           // primitive m-th root
    gl2* out;         // [m] bit-reversed
};

// V(x) = alpha^C (S0(x) - S0(zeta)) / (x - zeta) + (S1(x) - S1(g zeta)) / (x - g zeta), S = alpha-combinations of
// the committed columns at LDE position p
__global__ void __launch_bounds__(128) fri_batch_kernel(FriBatchArgs a) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.m) return;
    // S1 = sum_j alpha^j col_j (extension times base, component-wise): 192-bit accumulators, reduced once
    gl_acc192 u0 = gl_acc_zero(), u1 = gl_acc_zero();
    const gl* col = a.lde_t + p;
    for (size_t j = 0; j < a.C; j++) {
        const gl v = col[j * a.m];
        const gl2 w = a.apow[j];
        gl_acc_mac(u0, w.a0, v);
        gl_acc_mac(u1, w.a1, v);
    }
    const gl2 s1 = gl2_make(gl_acc_reduce(u0), gl_acc_reduce(u1));
    gl2 s0 = s1;
    for (size_t q = 0; q < 4; q++) s0 = gl2_add(s0, gl2_scale(a.apow[a.C + q], a.lde_q[q * a.m + p]));
    const gl x = gl_mul(GL_GEN, gl_pow(a.w_m, bitrev32((uint32_t)p, a.log_m)));
    const gl2 d0 = gl2_inv(gl2_sub(gl2_from(x), a.zeta));
    const gl2 d1 = gl2_inv(gl2_sub(gl2_from(x), a.zeta_next));
    const gl2 t0 = gl2_mul(gl2_mul(gl2_sub(s0, a.red0), d0), a.alpha_c);
    const gl2 t1 = gl2_mul(gl2_sub(s1, a.red1), d1);
    a.out[p] = gl2_add(t0, t1);
}

// one thread per coset of 16 consecutive (bit-reversed) evaluations: coefficients of P(x_c X) by a 16-point inverse
// DFT, then Horner at beta / x_c.  shift_inv = 1 / (domain shift of this layer), w_inv = inverse m-th root.
__global__ void __launch_bounds__(128) fri_fold_kernel(const gl2* __restrict__ in, size_t n_cosets, unsigned log_cosets,
                                                        gl shift_inv, gl w_inv, gl2 beta, gl2* __restrict__ out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cosets) return;
    gl2 u[16];
#pragma unroll
    for (int i = 0; i < 16; i++) u[i] = in[16 * c + bitrev32((uint32_t)i, 4)];  // natural order inside the coset
    const gl w16i = gl_inv(gl_root_of_unity(4));
    gl tw[16];
    tw[0] = 1;
    for (int i = 1; i < 16; i++) tw[i] = gl_mul(tw[i - 1], w16i);
    const gl inv16 = gl_inv(16);
    const gl xinv = gl_mul(shift_inv, gl_pow(w_inv, bitrev32((uint32_t)c, log_cosets)));
    const gl2 y = gl2_scale(beta, xinv);
    gl2 acc = gl2_from(0);
    for (int k = 15; k >= 0; k--) {  // a_k = (1/16) sum_i u_i w16^(-ik)
        gl2 ak = gl2_from(0);
#pragma unroll
        for (int i = 0; i < 16; i++) ak = gl2_add(ak, gl2_scale(u[i], tw[(i * k) & 15]));
        ak = gl2_scale(ak, inv16);
        acc = gl2_add(gl2_mul(acc, y), ak);
    }
    out[c] = acc;
}

// ------------------------------------------------------------------------------------------ K4: gathers
// out[q * q_stride + off + c] = src[c * m + idx[q] >> shift]
__global__ void gather_rows_kernel(const gl* __restrict__ src, size_t n_cols, size_t m, const uint32_t* __restrict__ idx,
                                   unsigned shift, gl* __restrict__ out, size_t q_stride, size_t off) {
    const size_t q = blockIdx.x;
    const size_t row = idx[q] >> shift;
    for (size_t c = threadIdx.x; c < n_cols; c += blockDim.x) out[q * q_stride + off + c] = src[c * m + row];
}
// contiguous leaves (row-major, leaf_len elements)
__global__ void gather_leaves_kernel(const gl* __restrict__ src, size_t leaf_len, const uint32_t* __restrict__ idx, unsigned shift,
                                     gl* __restrict__ out, size_t q_stride, size_t off) {
    const size_t q = blockIdx.x;
    const size_t row = idx[q] >> shift;
    for (size_t c = threadIdx.x; c < leaf_len; c += blockDim.x) out[q * q_stride + off + c] = src[row * leaf_len + c];
}
// sibling digests from the leaf level up to (excluding) the cap level
__global__ void gather_paths_kernel(const gl* __restrict__ digests, unsigned log_rows, unsigned n_sib, const uint32_t* __restrict__ idx,
                                    unsigned shift, gl* __restrict__ out, size_t q_stride, size_t off) {
    const size_t q = blockIdx.x;
    size_t i = idx[q] >> shift;
    size_t level_off = 0, rows = (size_t)1 << log_rows;
    for (unsigned l = 0; l < n_sib; l++) {
        if (threadIdx.x < 4) out[q * q_stride + off + 4 * l + threadIdx.x] = digests[4 * (level_off + (i ^ 1)) + threadIdx.x];
        level_off += rows;
        rows >>= 1;
        i >>= 1;
    }
}

static int launch_quotient(tmx_ctx* ctx, int table, const QuotientArgs& qa, cudaStream_t st) {
    const unsigned qblocks = (unsigned)((qa.m + 127) / 128);
    if (table == AIR_SHA256) quotient_kernel<AIR_SHA256><<<qblocks, 128, 0, st>>>(qa);
    else if (table == AIR_SHA512) quotient_kernel<AIR_SHA512><<<qblocks, 128, 0, st>>>(qa);
    else if (getenv("TMX_QUOTIENT_LITERAL")) quotient_kernel<AIR_ED25519><<<qblocks, 128, 0, st>>>(qa);  // debugging aid
    else {
        QuotientEdArgs ea;
        memset(&ea, 0, sizeof ea);
        ea.lde = qa.lde; ea.m = qa.m; ea.log_m = qa.log_m; ea.rate_bits = qa.rate_bits; ea.pertab = qa.pertab; ea.P = qa.P;
        ea.k[0] = ed_fast_consts(qa.alpha[0]);
        ea.k[1] = ed_fast_consts(qa.alpha[1]);
        memcpy(ea.zh_inv, qa.zh_inv, sizeof ea.zh_inv);
        ea.out = qa.out;
        quotient_ed25519_kernel<<<(unsigned)((qa.m + 63) / 64), 128, 0, st>>>(ea);
    }
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

static void fill_zh_inv(QuotientArgs& qa, size_t n) {
    const gl gn = gl_pow(GL_GEN, n);
    const gl wr = gl_root_of_unity(qa.rate_bits);  // x_j^n = 7^n * wr^j
    gl cur = gn;
    for (unsigned j = 0; j < (1u << qa.rate_bits); j++) {
        qa.zh_inv[j] = gl_inv(gl_sub(cur, 1));
        cur = gl_mul(cur, wr);
    }
}

// ------------------------------------------------------------------------------------------ host driver
static gl2 host_poly_eval_ext(const std::vector<gl2>& c, gl2 x) {
    gl2 acc = gl2_from(0);
    for (size_t i = c.size(); i-- > 0;) acc = gl2_add(gl2_mul(acc, x), c[i]);
    return acc;
}

unsigned fri_num_layers(unsigned degree_bits) {
    unsigned l = 0;
    while (degree_bits > STARK_FINAL_POLY_BITS && degree_bits + STARK_RATE_BITS - STARK_ARITY_BITS >= STARK_CAP_HEIGHT) {
        l++;
        degree_bits -= STARK_ARITY_BITS;
    }
    return l;
}

// device -> host through a pinned staging buffer owned by the prover (pageable destinations make every one of the
// ~12 transcript round trips per table a staged, driver-synchronised copy)
int TableProver::d2h(std::vector<gl>& dst, const gl* src, size_t n, cudaStream_t st) {
    if (sz_pinned < n) {
        if (h_pinned) cudaFreeHost(h_pinned);
        h_pinned = nullptr;
        sz_pinned = 0;
        const size_t want = std::max<size_t>(n + n / 4, 1 << 16);
        TMX_CUDA(cudaMallocHost((void**)&h_pinned, want * sizeof(gl)));
        sz_pinned = want;
    }
    TMX_CUDA(cudaMemcpyAsync(h_pinned, src, n * sizeof(gl), cudaMemcpyDeviceToHost, st));
    TMX_CUDA(cudaStreamSynchronize(st));
    dst.assign(h_pinned, h_pinned + n);
    return TMX_OK;
}

static void observe_ext(Challenger& ch, gl2 x) { ch.observe_ext(x); }

// TMX_TIMING=1: host wall-clock per phase (with a stream sync at each boundary) on stderr
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
    }
    PhaseTimer(cudaStream_t s) : on(getenv("TMX_TIMING") != nullptr), st(s), t0(0) {
        if (on) { cudaStreamSynchronize(st); t0 = now(); }
    }
    void tick(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const double t = now();
        fprintf(stderr, "  [tmx] %-22s %8.3f ms\n", what, t - t0);
        t0 = t;
    }
};

int TableProver::prove(tmx_ctx* ctx, int table, const gl* d_trace, unsigned log_n, Challenger& ch, std::vector<gl>& proof,
                       cudaStream_t st, const std::function<int(cudaEvent_t)>& on_trace_committed) {
    const size_t n = (size_t)1 << log_n, m = n << STARK_RATE_BITS;
    const unsigned km = log_n + STARK_RATE_BITS;
    const size_t C = (size_t)air_cols(table);
    const unsigned cap_h = std::min<unsigned>(km, STARK_CAP_HEIGHT);
    const size_t cap_n = (size_t)1 << cap_h;
    int rc;
    // ---- buffers (grow-only, owned by this prover) ----
    const size_t dig_t = tmx_merkle_digest_count(km, cap_h);
    rc = reserve(ctx, C, n, m, dig_t);
    if (rc) return rc;
    PhaseTimer pt(st);
    for (int i = 0; i < 3; i++)
        if (!ev_phase[i]) TMX_CUDA(cudaEventCreate(&ev_phase[i]));
    // ---- 1. trace commitment ----
    TMX_CUDA(cudaEventRecord(ev_phase[0], st));
    rc = tmx_lde(ctx, d_trace, d_lde, d_coeffs, C, log_n, STARK_RATE_BITS, st);
    if (rc) return rc;
    TMX_CUDA(cudaEventRecord(ev_phase[1], st));
    pt.tick("lde");
    rc = merkle_generic(ctx, d_lde, C, 1, m, km, cap_h, d_dig_t, st);
    if (rc) return rc;
    TMX_CUDA(cudaEventRecord(ev_phase[2], st));
    if (on_trace_committed) {
        rc = on_trace_committed(ev_phase[2]);
        if (rc) return rc;
    }
    pt.tick("trace merkle");
    std::vector<gl> cap;
    rc = d2h(cap, d_dig_t + 4 * (dig_t - cap_n), 4 * cap_n, st);
    if (rc) return rc;
    proof.insert(proof.end(), cap.begin(), cap.end());
    ch.observe(cap.data(), cap.size());
    pt.tick("cap d2h + observe");
    // ---- 2. constraint challenges, 3. quotient ----
    QuotientArgs qa;
    memset(&qa, 0, sizeof qa);
    qa.lde = d_lde; qa.m = m; qa.log_m = km; qa.rate_bits = STARK_RATE_BITS;
    qa.nper = air_n_periodic(table); qa.P = (int)air_period(table, n);
    qa.alpha[0] = ch.get();
    qa.alpha[1] = ch.get();
    fill_zh_inv(qa, n);
    if (qa.nper) {
        rc = periodic_tables(ctx, table, log_n, &qa.pertab);
        if (rc) return rc;
    }
    qa.out = d_qv;
    const unsigned qblocks = (unsigned)((m + 127) / 128);
    rc = launch_quotient(ctx, table, qa, st);
    if (rc) return rc;
    pt.tick("quotient kernel");
    // values on the coset (natural order) -> coefficients of Q(7 X); the 1/m factor comes with the inverse NTT
    rc = tmx_ntt(ctx, d_qv, 2, km, 1, st);
    if (rc) return rc;
    quotient_chunks_kernel<<<(unsigned)((4 * n + 255) / 256), 256, 0, st>>>(d_qv, n, gl_inv(gl_pow(GL_GEN, n)), d_qcoef);
    ctx->launches++;
    rc = lde_forward_cosets(ctx, d_qcoef, d_qlde, 4, log_n, STARK_RATE_BITS, st);
    if (rc) return rc;
    rc = merkle_generic(ctx, d_qlde, 4, 1, m, km, cap_h, d_dig_q, st);
    if (rc) return rc;
    rc = d2h(cap, d_dig_q + 4 * (dig_t - cap_n), 4 * cap_n, st);
    if (rc) return rc;
    proof.insert(proof.end(), cap.begin(), cap.end());
    ch.observe(cap.data(), cap.size());
    pt.tick("quotient commit");
    // ---- 4. openings at zeta and g * zeta (coefficients are stored coset-scaled: evaluate at zeta / 7) ----
    const gl2 zeta = ch.get_ext();
    const gl2 zeta_next = gl2_scale(zeta, gl_root_of_unity(log_n));
    const gl ginv = gl_inv(GL_GEN);
    ext_powers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gl2_scale(zeta, ginv), n, d_ypa);
    ext_powers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gl2_scale(zeta_next, ginv), n, d_ypb);
    eval_columns_kernel<<<(unsigned)C, 256, 0, st>>>(d_coeffs, n, d_ypa, d_ypb, d_open, d_open + C);
    eval_columns_kernel<<<4, 256, 0, st>>>(d_qcoef, n, d_ypa, nullptr, d_open + 2 * C, nullptr);
    ctx->launches += 4;
    TMX_CUDA(cudaGetLastError());
    pt.tick("openings kernels");
    std::vector<gl> op;
    rc = d2h(op, reinterpret_cast<const gl*>(d_open), 2 * (2 * C + 4), st);
    if (rc) return rc;
    proof.insert(proof.end(), op.begin(), op.end());  // local[C], next[C], quotient[4] as (a0, a1) pairs
    auto ext_at = [&](size_t i) { return gl2_make(op[2 * i], op[2 * i + 1]); };
    for (size_t c = 0; c < C; c++) observe_ext(ch, ext_at(c));
    for (size_t q = 0; q < 4; q++) observe_ext(ch, ext_at(2 * C + q));
    for (size_t c = 0; c < C; c++) observe_ext(ch, ext_at(C + c));
    pt.tick("openings d2h+observe");
    // ---- 5. FRI batch polynomial, pointwise ----
    const gl2 fa = ch.get_ext();
    std::vector<gl2> apow(C + 4);
    apow[0] = gl2_from(1);
    for (size_t j = 1; j < C + 4; j++) apow[j] = gl2_mul(apow[j - 1], fa);
    FriBatchArgs fb;
    memset(&fb, 0, sizeof fb);
    fb.red0 = gl2_from(0);
    fb.red1 = gl2_from(0);
    for (size_t j = 0; j < C; j++) {
        fb.red0 = gl2_add(fb.red0, gl2_mul(apow[j], ext_at(j)));
        fb.red1 = gl2_add(fb.red1, gl2_mul(apow[j], ext_at(C + j)));
    }
    for (size_t q = 0; q < 4; q++) fb.red0 = gl2_add(fb.red0, gl2_mul(apow[C + q], ext_at(2 * C + q)));
    TMX_CUDA(cudaMemcpyAsync(d_apow, apow.data(), apow.size() * sizeof(gl2), cudaMemcpyHostToDevice, st));
    fb.lde_t = d_lde; fb.lde_q = d_qlde; fb.C = C; fb.m = m; fb.log_m = km; fb.apow = d_apow;
    fb.zeta = zeta; fb.zeta_next = zeta_next; fb.alpha_c = gl2_mul(apow[C - 1], fa);
    fb.w_m = gl_root_of_unity(km);
    fb.out = d_fri[0];
    fri_batch_kernel<<<qblocks, 128, 0, st>>>(fb);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    pt.tick("fri batch");
    // ---- 6. FRI commit phase: Merkle over cosets of 16, fold with beta ----
    const unsigned n_layers = fri_num_layers(log_n);
    size_t cur = m;
    gl shift = GL_GEN;
    std::vector<size_t> layer_dig_off(n_layers), layer_rows_log(n_layers);
    size_t dig_used = 0;
    for (unsigned l = 0; l < n_layers; l++) {
        const unsigned lg_rows = ilog2(cur) - STARK_ARITY_BITS;
        const unsigned lcap_h = std::min<unsigned>(lg_rows, STARK_CAP_HEIGHT);
        const size_t ldig = tmx_merkle_digest_count(lg_rows, lcap_h);
        layer_dig_off[l] = dig_used;
        layer_rows_log[l] = lg_rows;
        gl* dg = d_dig_fri + 4 * dig_used;
        rc = merkle_generic(ctx, reinterpret_cast<const gl*>(d_fri[l]), 32, 32, 1, lg_rows, lcap_h, dg, st);
        if (rc) return rc;
        dig_used += ldig;
        rc = d2h(cap, dg + 4 * (ldig - ((size_t)1 << lcap_h)), 4 * ((size_t)1 << lcap_h), st);
        if (rc) return rc;
        proof.insert(proof.end(), cap.begin(), cap.end());
        ch.observe(cap.data(), cap.size());
        const gl2 beta = ch.get_ext();
        const size_t cosets = cur >> STARK_ARITY_BITS;
        fri_fold_kernel<<<(unsigned)((cosets + 127) / 128), 128, 0, st>>>(d_fri[l], cosets, lg_rows, gl_inv(shift),
                                                                         gl_inv(gl_root_of_unity(ilog2(cur))), beta, d_fri[l + 1]);
        ctx->launches++;
        TMX_CUDA(cudaGetLastError());
        cur = cosets;
        shift = gl_pow(shift, 16);
    }
    pt.tick("fri layers");
    // final polynomial: the remaining `cur` evaluations (bit-reversed) on shift * <w_cur> -> coefficients
    std::vector<gl> fin;
    rc = d2h(fin, reinterpret_cast<const gl*>(d_fri[n_layers]), 2 * cur, st);
    if (rc) return rc;
    {
        const unsigned lg = ilog2(cur);
        const gl winv = gl_inv(gl_root_of_unity(lg)), sinv = gl_inv(shift), ninv = gl_inv((gl)cur);
        std::vector<gl2> coef(cur);
        for (size_t k = 0; k < cur; k++) {
            gl2 acc = gl2_from(0);
            for (size_t i = 0; i < cur; i++) {
                const size_t pos = bitrev32((uint32_t)i, lg);
                acc = gl2_add(acc, gl2_scale(gl2_make(fin[2 * pos], fin[2 * pos + 1]), gl_pow(winv, (i * k) % cur)));
            }
            coef[k] = gl2_scale(acc, gl_mul(ninv, gl_pow(sinv, k)));
        }
        const size_t final_len = cur >> STARK_RATE_BITS;
        for (size_t k = final_len; k < cur; k++)
            if (coef[k].a0 || coef[k].a1) return fail(TMX_E_CUDA, "FRI: final polynomial exceeds its degree bound (internal error)");
        proof.push_back((gl)final_len);
        for (size_t k = 0; k < final_len; k++) {
            proof.push_back(coef[k].a0);
            proof.push_back(coef[k].a1);
            observe_ext(ch, coef[k]);
        }
    }
    pt.tick("final poly");
    // ---- 7. proof of work (K9) ----
    {
        gl state[12];
        for (int i = 0; i < 12; i++) state[i] = ch.state[i];
        for (int i = 0; i < ch.n_in; i++) state[i] = ch.in[i];
        uint64_t wit = 0;
        rc = pow_grind(ctx, state, ch.n_in, STARK_POW_BITS, &wit, st);
        if (rc) return rc;
        ch.observe((gl)wit);
        (void)ch.get();
        proof.push_back((gl)wit);
    }
    pt.tick("pow");
    // ---- 8. queries (K4) ----
    std::vector<uint32_t> idx(STARK_NUM_QUERIES);
    for (int q = 0; q < STARK_NUM_QUERIES; q++) idx[q] = (uint32_t)(ch.get() % m);
    TMX_CUDA(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    const unsigned n_sib = km - cap_h;
    size_t qlen = C + 4 * n_sib + 4 + 4 * n_sib;
    for (unsigned l = 0; l < n_layers; l++)
        qlen += 32 + 4 * (layer_rows_log[l] - std::min<unsigned>((unsigned)layer_rows_log[l], STARK_CAP_HEIGHT));
    rc = reserve_queries(ctx, qlen * STARK_NUM_QUERIES);
    if (rc) return rc;
    size_t off = 0;
    const unsigned NQ = STARK_NUM_QUERIES;
    gather_rows_kernel<<<NQ, 256, 0, st>>>(d_lde, C, m, d_idx, 0, d_query, qlen, off);
    off += C;
    gather_paths_kernel<<<NQ, 32, 0, st>>>(d_dig_t, km, n_sib, d_idx, 0, d_query, qlen, off);
    off += 4 * n_sib;
    gather_rows_kernel<<<NQ, 32, 0, st>>>(d_qlde, 4, m, d_idx, 0, d_query, qlen, off);
    off += 4;
    gather_paths_kernel<<<NQ, 32, 0, st>>>(d_dig_q, km, n_sib, d_idx, 0, d_query, qlen, off);
    off += 4 * n_sib;
    ctx->launches += 4;
    for (unsigned l = 0; l < n_layers; l++) {
        const unsigned lg_rows = (unsigned)layer_rows_log[l];
        const unsigned ns = lg_rows - std::min<unsigned>(lg_rows, STARK_CAP_HEIGHT);
        const unsigned sh = STARK_ARITY_BITS * (l + 1);
        gather_leaves_kernel<<<NQ, 32, 0, st>>>(reinterpret_cast<const gl*>(d_fri[l]), 32, d_idx, sh, d_query, qlen, off);
        off += 32;
        gather_paths_kernel<<<NQ, 32, 0, st>>>(d_dig_fri + 4 * layer_dig_off[l], lg_rows, ns, d_idx, sh, d_query, qlen, off);
        off += 4 * ns;
        ctx->launches += 2;
    }
    TMX_CUDA(cudaGetLastError());
    std::vector<gl> qd;
    rc = d2h(qd, d_query, qlen * NQ, st);
    if (rc) return rc;
    proof.insert(proof.end(), qd.begin(), qd.end());
    pt.tick("queries");
    // the stream has been synchronised by the copy above: the stamps of this table's commitment are complete
    TMX_CUDA(cudaEventElapsedTime(&last_lde_ms, ev_phase[0], ev_phase[1]));
    TMX_CUDA(cudaEventElapsedTime(&last_merkle_ms, ev_phase[1], ev_phase[2]));
    (void)host_poly_eval_ext;
    return TMX_OK;
}

static int grow(void** p, size_t* have, size_t want) {
    if (*have >= want) return TMX_OK;
    if (*p) TMX_CUDA(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    TMX_CUDA(cudaMalloc(p, want));
    *have = want;
    return TMX_OK;
}

int TableProver::reserve(tmx_ctx* ctx, size_t C, size_t n, size_t m, size_t dig_t) {
    (void)ctx;
    int rc;
    if ((rc = grow((void**)&d_lde, &sz_lde, C * m * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_coeffs, &sz_coeffs, C * n * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_dig_t, &sz_dig_t, 4 * dig_t * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_dig_q, &sz_dig_q, 4 * dig_t * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_dig_fri, &sz_dig_fri, 4 * (m / 4) * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_qv, &sz_qv, 2 * m * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_qcoef, &sz_qcoef, 4 * n * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_qlde, &sz_qlde, 4 * m * sizeof(gl)))) return rc;
    if ((rc = grow((void**)&d_ypa, &sz_yp, 2 * n * sizeof(gl2)))) return rc;
    d_ypb = d_ypa + n;
    if ((rc = grow((void**)&d_open, &sz_open, (2 * C + 4) * sizeof(gl2)))) return rc;
    if ((rc = grow((void**)&d_apow, &sz_apow, (C + 4) * sizeof(gl2)))) return rc;
    if ((rc = grow((void**)&d_idx, &sz_idx, STARK_NUM_QUERIES * sizeof(uint32_t)))) return rc;
    // FRI layer arrays: m, m/16, m/256, ... ext values, carved from one allocation
    size_t tot = 0, cur = m;
    for (int l = 0; l < 9; l++) {
        tot += cur;
        cur = cur >> 4 ? cur >> 4 : 1;
    }
    if ((rc = grow((void**)&d_fri_base, &sz_fri, tot * sizeof(gl2)))) return rc;
    cur = m;
    size_t o = 0;
    for (int l = 0; l < 9; l++) {
        d_fri[l] = d_fri_base + o;
        o += cur;
        cur = cur >> 4 ? cur >> 4 : 1;
    }
    return TMX_OK;
}

int TableProver::reserve_queries(tmx_ctx* ctx, size_t n) {
    (void)ctx;
    return grow((void**)&d_query, &sz_query, n * sizeof(gl));
}

int TableProver::periodic_tables(tmx_ctx* ctx, int table, unsigned log_n, const gl** out) {
    const uint64_t key = ((uint64_t)shape.kind << 48) | ((uint64_t)shape.n_max << 16) | ((uint64_t)table << 8) | log_n;
    auto it = pertabs.find(key);
    if (it != pertabs.end()) {
        *out = it->second;
        return TMX_OK;
    }
    // the host-side values depend on (shape, table, log_n) only: computed once per process (the SHA-256 table's public
    // columns cost six NTTs of the table length), shared by every prover / context
    static std::mutex cache_mutex;
    static std::map<uint64_t, std::shared_ptr<const std::vector<gl>>> cache;
    std::shared_ptr<const std::vector<gl>> host;
    {
        std::lock_guard<std::mutex> lk(cache_mutex);
        auto hit = cache.find(key);
        if (hit == cache.end())
            hit = cache.emplace(key, std::make_shared<const std::vector<gl>>(air_periodic_lde_table(table, log_n, h_K256, h_K512, shape))).first;
        host = hit->second;
    }
    const std::vector<gl>& tab = *host;
    void* d = nullptr;
    TMX_CUDA(cudaMalloc(&d, tab.size() * sizeof(gl)));
    TMX_CUDA(cudaMemcpy(d, tab.data(), tab.size() * sizeof(gl), cudaMemcpyHostToDevice));
    TMX_CUDA(cudaStreamSynchronize(cudaStreamLegacy));  // staged copy: the DMA must land before kernels on other streams read it
    pertabs[key] = (gl*)d;
    *out = (gl*)d;
    (void)ctx;
    return TMX_OK;
}

void TableProver::release() {
    for (int i = 0; i < 3; i++)
        if (ev_phase[i]) cudaEventDestroy(ev_phase[i]);
    if (h_pinned) cudaFreeHost(h_pinned);
    void* ps[] = {d_lde, d_coeffs, d_dig_t, d_dig_q, d_dig_fri, d_qv, d_qcoef, d_qlde, d_ypa, d_open, d_apow, d_idx, d_fri_base, d_query};
    for (void* p : ps)
        if (p) cudaFree(p);
    for (auto& kv : pertabs) cudaFree(kv.second);
    pertabs.clear();
    *this = TableProver();
}

}  // namespace tmx

using namespace tmx;

// Host-side self check (no GPU): the literal constraint fold of air_ed25519() and the factored evaluation used by the
// quotient kernel on one (local row, next row) pair.  out = {literal(alpha0), literal(alpha1), fast(alpha0), fast(alpha1)}.
struct HostRow {
    const gl* p;
    FB operator[](int c) const { return FB::mk(p[c]); }
};
extern "C" int tmx_host_air_ed25519(const uint64_t* row_l, const uint64_t* row_n, const uint64_t periodic[3], const uint64_t alpha[2],
                                    uint64_t out[4]) {
    if (!row_l || !row_n || !periodic || !alpha || !out) return fail(TMX_E_INPUT, "tmx_host_air_ed25519: NULL argument");
    HostRow l{row_l}, n{row_n};
    const FB per[3] = {FB::mk(periodic[0]), FB::mk(periodic[1]), FB::mk(periodic[2])};
    ConstraintAcc<FB> acc;
    acc.acc0 = FB::c(0); acc.acc1 = FB::c(0);
    acc.alpha0 = FB::mk(alpha[0]); acc.alpha1 = FB::mk(alpha[1]);
    air_ed25519<FB>(l, n, per, acc);
    out[0] = acc.acc0.v;
    out[1] = acc.acc1.v;
    out[2] = ed25519_constraints_fast(l, n, periodic, ed_fast_consts(alpha[0]));
    out[3] = ed25519_constraints_fast(l, n, periodic, ed_fast_consts(alpha[1]));
    return TMX_OK;
}

// K5 as a kernel-level entry point (parity tests, ncu): constraint quotient of one table on its LDE coset.
extern "C" int tmx_quotient(tmx_ctx* ctx, uint32_t kind, uint32_t n_max, int table, const uint64_t* d_lde, unsigned log_n,
                            const uint64_t alpha[2], uint64_t* d_out, void* stream) {
    if (!ctx || !d_lde || !alpha || !d_out || table < 0 || table > 2 || log_n < 8 || log_n > 28)
        return fail(TMX_E_INPUT, "tmx_quotient: bad arguments");
    static thread_local TableProver tp;  // only its periodic-table cache is used
    tp.shape = AirShape{kind, n_max};
    QuotientArgs qa;
    memset(&qa, 0, sizeof qa);
    const size_t n = (size_t)1 << log_n;
    qa.lde = d_lde; qa.m = n << STARK_RATE_BITS; qa.log_m = log_n + STARK_RATE_BITS; qa.rate_bits = STARK_RATE_BITS;
    qa.nper = air_n_periodic(table); qa.P = (int)air_period(table, n);
    qa.alpha[0] = alpha[0]; qa.alpha[1] = alpha[1];
    fill_zh_inv(qa, n);
    if (qa.nper) {
        int rc = tp.periodic_tables(ctx, table, log_n, &qa.pertab);
        if (rc) return rc;
    }
    qa.out = d_out;
    return launch_quotient(ctx, table, qa, pick_stream(ctx, stream));
}
