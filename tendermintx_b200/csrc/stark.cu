// GPU STARK prover for the witness tables on a shared bus: round 1 (first-round traces, range-lookup histogram), round 2
// (helper columns and running sums of the bus interactions), then per table K5 (constraint quotient over the LDE),
// openings at zeta / g*zeta, K3 (FRI batch combination and arity-16 folding in EVALUATION space), K4 (query / opening
// gather), driven by a host-side duplex challenger.  Commitments use K1 (tmx_lde) and K2 (Poseidon Merkle).
//
// Replaces, on the GPU, the plonky2 / starky / Curta proving loops behind `circuit.prove()`
// [REF circuits/skip.rs:214,244; circuits/step.rs:196,223]: PolynomialBatch::from_values, the lookup / bus accumulators,
// the quotient computation, fri/oracle.rs prove_openings, fri/prover.rs fri_committed_trees / fri_proof_of_work / query
// rounds.  Same functions of the same field elements as the CPU oracle (oracle/stark.c), different algorithms: the
// constraints are compiled templates here and interpreted data there, the batch polynomial is formed pointwise on the LDE
// instead of dividing coefficient vectors, and FRI layers are folded by a 16-point inverse NTT per coset instead of
// folding coefficients and re-running a coset FFT.
#include "stark.cuh"
#include "stark_rows.cuh"
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <algorithm>
#include <memory>
#include <mutex>

namespace tmx {

// The Ed25519 row wants 218 registers (2 CTAs per SM, 11 % occupancy); capped at 128 it spills about 0.5 KB per thread to L1
// and runs 1 ms faster (four CTAs per SM, and the co-running tables' kernels still fit beside it).  The same cap on the two
// SHA tables (~250 registers uncapped) makes their kernels faster when they run alone (1.2 -> 0.8 / 0.9 ms: straight-line code
// waiting on column loads and instruction fetch) but the pool slower (46.0 against 45.6 ms per proof, A/B on one box): uncapped.
template <int TABLE>
__global__ void __launch_bounds__(128, TABLE == AIR_ED25519 ? 4 : 1) quotient_kernel(QuotientArgs a) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.m) quotient_point<TABLE>(a, p);
}
// helper columns of the table's bus interactions, one thread per trace row
template <int TABLE>
__global__ void __launch_bounds__(128) bus_gen_kernel(BusPassArgs a) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < a.n) bus_gen_row<TABLE>(a, r);
}
// histogram of the table's range lookups
template <int TABLE>
__global__ void __launch_bounds__(128) bus_count_kernel(BusPassArgs a) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < a.n) bus_count_row<TABLE>(a, r);
}
// running sum: Z(0) = 0, Z(r + 1) = Z(r) + rowsum(r) - total / n (closes up cyclically); one CTA
__global__ void __launch_bounds__(1024) bus_scan_kernel(const gl2* __restrict__ rowsum, size_t n, gl* __restrict__ z0, gl* __restrict__ z1,
                                                         gl* __restrict__ total_out) {
    __shared__ gl2 s[1024];
    __shared__ gl2 step;
    const size_t chunk = (n + 1023) / 1024;
    const size_t lo = std::min(n, (size_t)threadIdx.x * chunk), hi = std::min(n, lo + chunk);
    gl2 acc = gl2_from(0);
    for (size_t r = lo; r < hi; r++) acc = gl2_add(acc, rowsum[r]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {  // inclusive scan
        gl2 v = gl2_from(0);
        if ((int)threadIdx.x >= d) v = s[threadIdx.x - d];
        __syncthreads();
        s[threadIdx.x] = gl2_add(s[threadIdx.x], v);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const gl2 total = s[1023];
        step = gl2_scale(total, gl_inv((gl)n));
        total_out[0] = total.a0;
        total_out[1] = total.a1;
    }
    __syncthreads();
    gl2 z = threadIdx.x ? s[threadIdx.x - 1] : gl2_from(0);
    z = gl2_sub(z, gl2_scale(step, (gl)lo));
    for (size_t r = lo; r < hi; r++) {
        z0[r] = z.a0;
        z1[r] = z.a1;
        z = gl2_sub(gl2_add(z, rowsum[r]), step);
    }
}
// first-round trace of the range table from the histogram
__global__ void range_fill_kernel(const unsigned int* __restrict__ hist, gl* __restrict__ trace, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    trace[(size_t)RG_M16 * n + i] = hist[i];
    trace[(size_t)RG_M11 * n + i] = i < (1u << 11) ? hist[(1u << 16) + i] : 0;
    trace[(size_t)RG_M8 * n + i] = i < (1u << 8) ? hist[(1u << 16) + (1u << 11) + i] : 0;
    trace[(size_t)RG_M1 * n + i] = i < 2 ? hist[(1u << 16) + (1u << 11) + (1u << 8) + i] : 0;
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
    const gl* seg[3];  // constant, first-round, second-round LDE columns, [count][m] each
    size_t segc[3];
    const gl* lde_q;   // [4][m]
    size_t C, m;       // C = total number of columns opened at both points
    unsigned log_m;
    const gl2* apow;   // alpha^j, j < C + 4
    gl2 red0, red1, zeta, zeta_next, alpha_c;
    gl w_m;            // primitive m-th root
    gl2* out;          // [m] bit-reversed
};

// V(x) = alpha^C (S0(x) - S0(zeta)) / (x - zeta) + (S1(x) - S1(g zeta)) / (x - g zeta), S = alpha-combinations of
// the committed columns at LDE position p
__global__ void __launch_bounds__(128) fri_batch_kernel(FriBatchArgs a) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.m) return;
    // S1 = sum_j alpha^j col_j (extension times base, component-wise): 192-bit accumulators, reduced once
    gl_acc192 u0 = gl_acc_zero(), u1 = gl_acc_zero();
    size_t j0 = 0;
    for (int s = 0; s < 3; s++) {
        const gl* col = a.seg[s] + p;
        for (size_t j = 0; j < a.segc[s]; j++) {
            const gl v = col[j * a.m];
            const gl2 w = a.apow[j0 + j];
            gl_acc_mac(u0, w.a0, v);
            gl_acc_mac(u1, w.a1, v);
        }
        j0 += a.segc[s];
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

// ------------------------------------------------------------------------------------------ host driver
static int launch_quotient(tmx_ctx* ctx, int table, const QuotientArgs& qa, cudaStream_t st) {
    const unsigned qblocks = (unsigned)((qa.m + 127) / 128);
    switch (table) {
        case AIR_SHA256: quotient_kernel<AIR_SHA256><<<qblocks, 128, 0, st>>>(qa); break;
        case AIR_SHA512: quotient_kernel<AIR_SHA512><<<qblocks, 128, 0, st>>>(qa); break;
        case AIR_ED25519: quotient_kernel<AIR_ED25519><<<qblocks, 128, 0, st>>>(qa); break;
        case AIR_LOGIC: quotient_kernel<AIR_LOGIC><<<qblocks, 128, 0, st>>>(qa); break;
        default: quotient_kernel<AIR_RANGE><<<qblocks, 128, 0, st>>>(qa); break;
    }
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}
static int launch_bus_gen(tmx_ctx* ctx, int table, const BusPassArgs& a, cudaStream_t st) {
    const unsigned T = 128;
    const unsigned blocks = (unsigned)((a.n + T - 1) / T);
    switch (table) {
        case AIR_SHA256: bus_gen_kernel<AIR_SHA256><<<blocks, T, 0, st>>>(a); break;
        case AIR_SHA512: bus_gen_kernel<AIR_SHA512><<<blocks, T, 0, st>>>(a); break;
        case AIR_ED25519: bus_gen_kernel<AIR_ED25519><<<blocks, T, 0, st>>>(a); break;
        case AIR_LOGIC: bus_gen_kernel<AIR_LOGIC><<<blocks, T, 0, st>>>(a); break;
        default: bus_gen_kernel<AIR_RANGE><<<blocks, T, 0, st>>>(a); break;
    }
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}
static int launch_bus_count(tmx_ctx* ctx, int table, const BusPassArgs& a, cudaStream_t st) {
    const unsigned T = 128;
    const unsigned blocks = (unsigned)((a.n + T - 1) / T);
    switch (table) {
        case AIR_SHA256: bus_count_kernel<AIR_SHA256><<<blocks, T, 0, st>>>(a); break;
        case AIR_SHA512: bus_count_kernel<AIR_SHA512><<<blocks, T, 0, st>>>(a); break;
        case AIR_ED25519: bus_count_kernel<AIR_ED25519><<<blocks, T, 0, st>>>(a); break;
        case AIR_LOGIC: bus_count_kernel<AIR_LOGIC><<<blocks, T, 0, st>>>(a); break;
        default: return TMX_OK;  // the range table provides, it does not look up
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

unsigned fri_num_layers(unsigned degree_bits) {
    unsigned l = 0;
    while (degree_bits > STARK_FINAL_POLY_BITS && degree_bits + STARK_RATE_BITS - STARK_ARITY_BITS >= STARK_CAP_HEIGHT) {
        l++;
        degree_bits -= STARK_ARITY_BITS;
    }
    return l;
}

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

// device -> host through a pinned staging buffer owned by the prover (pageable destinations make every one of the
// transcript round trips a staged, driver-synchronised copy)
int Prover::d2h(std::vector<gl>& dst, const gl* src, size_t n, cudaStream_t st, int slot) {
    Tail& tl = tail[slot];
    if (tl.sz_pinned < n) {
        if (tl.h_pinned) cudaFreeHost(tl.h_pinned);
        tl.h_pinned = nullptr;
        tl.sz_pinned = 0;
        const size_t want = std::max<size_t>(n + n / 4, 1 << 16);
        TMX_CUDA(cudaMallocHost((void**)&tl.h_pinned, want * sizeof(gl)));
        tl.sz_pinned = want;
    }
    TMX_CUDA(cudaMemcpyAsync(tl.h_pinned, src, n * sizeof(gl), cudaMemcpyDeviceToHost, st));
    // Wait for the copy: poll for up to 200 us (one proof at a time: the kernels in front of the copy are short and
    // a sleeping thread wakes up late), then sleep on a blocking-sync event.  With several proofs in flight the wait is
    // milliseconds of other proofs' kernels, and twenty spinning threads per GPU starve a host with four cores per GPU
    // (8 GPUs, 32 cores: 48.3 ms per proof spinning, 46.4 ms sleeping).
    if (!tl.ev_wait) TMX_CUDA(cudaEventCreateWithFlags(&tl.ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
    TMX_CUDA(cudaEventRecord(tl.ev_wait, st));
    static const int spin_us = getenv("TMX_SYNC_SPIN_US") ? atoi(getenv("TMX_SYNC_SPIN_US")) : 200;
    const double t_end = PhaseTimer::now() + spin_us * 1e-3;
    cudaError_t q = cudaEventQuery(tl.ev_wait);
    while (q == cudaErrorNotReady && PhaseTimer::now() < t_end) q = cudaEventQuery(tl.ev_wait);
    if (q == cudaErrorNotReady) q = cudaEventSynchronize(tl.ev_wait);
    if (q != cudaSuccess) return fail(TMX_E_CUDA, std::string("device-to-host copy: ") + cudaGetErrorString(q));
    dst.assign(tl.h_pinned, tl.h_pinned + n);
    return TMX_OK;
}

int Prover::alloc(void** p, size_t bytes) {
    *p = nullptr;
    if (!bytes) bytes = 8;
    TMX_CUDA(cudaMalloc(p, bytes));
    owned.push_back(*p);
    return TMX_OK;
}

static unsigned cap_height_of(unsigned km) { return std::min<unsigned>(km, STARK_CAP_HEIGHT); }

int Prover::setup(tmx_ctx* ctx, std::shared_ptr<const CircuitDef> d) {
    def = d;
    shape = air_shape(d->kind, d->n_max, d->chain_id.data(), d->chain_id.size());
    cudaStream_t st = ctx->stream;
    int rc;
    size_t max_m = 0, max_ct = 0;
    for (int t = 0; t < STARK_N_TABLES; t++) {
        const TableDef& td = def->tables[t];
        if (!td.n_main) continue;
        TableDevice& tb = tab[t];
        const size_t n = td.rows(), m = n << STARK_RATE_BITS, Kc = td.n_const, C = td.n_main, A = (size_t)td.n_aux();
        const unsigned km = td.log_n + STARK_RATE_BITS;
        const size_t dig = tmx_merkle_digest_count(km, cap_height_of(km));
        max_m = std::max(max_m, m);
        max_ct = std::max(max_ct, Kc + C + A);
        if ((rc = alloc((void**)&tb.d_const, Kc * n * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_const_coef, Kc * n * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_const_lde, Kc * m * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_const_dig, 4 * dig * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_per_trace, (size_t)td.n_per * td.period * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_per_lde, (size_t)td.n_per * 2 * td.period * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_aux, A * n * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_lde_m, C * m * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_coef_m, C * n * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_dig_m, 4 * dig * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_lde_a, A * m * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_coef_a, A * n * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tb.d_dig_a, 4 * dig * sizeof(gl)))) return rc;
        // constant columns: upload, commit, compare the cap with the definition's (computed on the host)
        if (Kc) {
            TMX_CUDA(cudaMemcpyAsync(tb.d_const, td.constants.data(), Kc * n * sizeof(gl), cudaMemcpyHostToDevice, st));
            if ((rc = tmx_lde(ctx, tb.d_const, tb.d_const_lde, tb.d_const_coef, Kc, td.log_n, STARK_RATE_BITS, st))) return rc;
            if ((rc = merkle_generic(ctx, tb.d_const_lde, Kc, 1, m, km, cap_height_of(km), tb.d_const_dig, st))) return rc;
            const size_t cap_w = 4 * ((size_t)1 << cap_height_of(km));
            std::vector<gl> cap;
            if ((rc = d2h(cap, tb.d_const_dig + 4 * dig - cap_w, cap_w, st))) return rc;
            if (cap != td.const_cap) return fail(TMX_E_CUDA, "constant-column commitment of table " + std::to_string(t) + " differs between the GPU and the host");
        }
        if (td.n_per) {
            TMX_CUDA(cudaMemcpyAsync(tb.d_per_trace, td.periodic.data(), td.periodic.size() * sizeof(gl), cudaMemcpyHostToDevice, st));
            // values of the periodic columns on the LDE coset: the interpolant of one period composed with x -> x^(n/P)
            const size_t P = td.period;
            std::vector<gl> lde((size_t)td.n_per * 2 * P);
            const gl sh = gl_pow(GL_GEN, n / P);
            for (uint32_t pc = 0; pc < td.n_per; pc++) {
                std::vector<gl> coef(td.periodic.begin() + (size_t)pc * P, td.periodic.begin() + (size_t)(pc + 1) * P);
                air_host_ntt(coef, true);
                coef.resize(2 * P, 0);
                gl s = 1;
                for (size_t k = 0; k < P; k++) {
                    coef[k] = gl_mul(coef[k], s);
                    s = gl_mul(s, sh);
                }
                air_host_ntt(coef, false);
                std::copy(coef.begin(), coef.end(), lde.begin() + (size_t)pc * 2 * P);
            }
            TMX_CUDA(cudaMemcpyAsync(tb.d_per_lde, lde.data(), lde.size() * sizeof(gl), cudaMemcpyHostToDevice, st));
            TMX_CUDA(cudaStreamSynchronize(st));
        }
    }
    const size_t max_n = max_m >> STARK_RATE_BITS;
    for (int t = 0; t < STARK_N_TABLES; t++) {
        const TableDef& td = def->tables[t];
        if (!td.n_main) continue;
        Tail& tl = tail[t];
        const size_t n = td.rows(), m = n << STARK_RATE_BITS, ct = td.n_const + td.n_main + (size_t)td.n_aux();
        const unsigned km = td.log_n + STARK_RATE_BITS;
        if ((rc = alloc((void**)&tl.d_dig_q, 4 * tmx_merkle_digest_count(km, cap_height_of(km)) * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tl.d_dig_fri, 4 * (m / 4) * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tl.d_qv, 2 * m * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tl.d_ntt_tmp, 2 * m * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tl.d_pow, 16 * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tl.d_qcoef, 4 * n * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tl.d_qlde, 4 * m * sizeof(gl)))) return rc;
        if ((rc = alloc((void**)&tl.d_ypa, 2 * n * sizeof(gl2)))) return rc;
        tl.d_ypb = tl.d_ypa + n;
        if ((rc = alloc((void**)&tl.d_open, (2 * ct + 4) * sizeof(gl2)))) return rc;
        if ((rc = alloc((void**)&tl.d_apow, (ct + 4) * sizeof(gl2)))) return rc;
        if ((rc = alloc((void**)&tl.d_idx, STARK_NUM_QUERIES * sizeof(uint32_t)))) return rc;
        // FRI layer arrays: m, m/16, m/256, ... ext values, carved from one allocation
        size_t tot = 0, cur = m;
        for (int l = 0; l < 9; l++) {
            tot += cur;
            cur = cur >> 4 ? cur >> 4 : 1;
        }
        gl2* base = nullptr;
        if ((rc = alloc((void**)&base, tot * sizeof(gl2)))) return rc;
        cur = m;
        for (int l = 0; l < 9; l++) {
            tl.d_fri[l] = base;
            base += cur;
            cur = cur >> 4 ? cur >> 4 : 1;
        }
    }
    rowsum_stride = max_n;
    if ((rc = alloc((void**)&d_rowsum, (size_t)STARK_N_TABLES * max_n * sizeof(gl2)))) return rc;
    if ((rc = alloc((void**)&d_small, (size_t)STARK_N_TABLES * 66 * sizeof(gl)))) return rc;
    TMX_CUDA(cudaMemsetAsync(d_small, 0, (size_t)STARK_N_TABLES * 66 * sizeof(gl), st));  // the totals' slots are copied (unused) in round 1
    if ((rc = alloc((void**)&d_hist, (BUS_HIST_SIZE + 4) * sizeof(unsigned int)))) return rc;
    if (def->tables[AIR_RANGE].n_main)
        if ((rc = alloc((void**)&d_range_trace, def->tables[AIR_RANGE].rows() * RG_COLS * sizeof(gl)))) return rc;
    for (auto& e : ev_phase) TMX_CUDA(cudaEventCreate(&e));
    TMX_CUDA(cudaStreamSynchronize(st));
    return TMX_OK;
}

void Prover::release() {
    for (auto& e : ev_phase)
        if (e) cudaEventDestroy(e);
    for (auto& tl : tail) {
        if (tl.ev_wait) cudaEventDestroy(tl.ev_wait);
        if (tl.h_pinned) cudaFreeHost(tl.h_pinned);
        if (tl.d_query) cudaFree(tl.d_query);
    }
    for (void* p : owned) cudaFree(p);
    *this = Prover();
}

static BusPassArgs bus_pass_args(const Prover& pr, int t, const gl* d_trace) {
    const TableDef& td = pr.def->tables[t];
    BusPassArgs a;
    memset(&a, 0, sizeof a);
    a.trace = d_trace;
    a.kconst = pr.tab[t].d_const;
    a.per = pr.tab[t].d_per_trace;
    a.n = td.rows();
    a.P = (int)td.period;
    a.shape = pr.shape;
    return a;
}

int Prover::count_lookups(tmx_ctx* ctx, int t, const gl* d_trace, cudaStream_t st) {
    if (!def->tables[t].n_main || t == AIR_RANGE) return TMX_OK;
    BusPassArgs a = bus_pass_args(*this, t, d_trace);
    a.hist = d_hist;
    return launch_bus_count(ctx, t, a, st);
}

int Prover::fill_range_trace(tmx_ctx* ctx, cudaStream_t st) {
    const size_t n = def->tables[AIR_RANGE].rows();
    range_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_hist, d_range_trace, n);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

int Prover::commit_main(tmx_ctx* ctx, int t, const gl* d_trace, cudaStream_t st) {
    const TableDef& td = def->tables[t];
    if (!td.n_main) return TMX_OK;
    TableDevice& tb = tab[t];
    const size_t m = td.rows() << STARK_RATE_BITS;
    const unsigned km = td.log_n + STARK_RATE_BITS;
    const size_t dig = tmx_merkle_digest_count(km, cap_height_of(km)), cap_w = 4 * ((size_t)1 << cap_height_of(km));
    int rc;
    TMX_CUDA(cudaEventRecord(ev_phase[3 * t], st));
    if ((rc = tmx_lde(ctx, d_trace, tb.d_lde_m, tb.d_coef_m, td.n_main, td.log_n, STARK_RATE_BITS, st))) return rc;
    TMX_CUDA(cudaEventRecord(ev_phase[3 * t + 1], st));
    if ((rc = merkle_generic(ctx, tb.d_lde_m, td.n_main, 1, m, km, cap_height_of(km), tb.d_dig_m, st))) return rc;
    TMX_CUDA(cudaEventRecord(ev_phase[3 * t + 2], st));
    TMX_CUDA(cudaMemcpyAsync(d_small + 66 * t, tb.d_dig_m + 4 * dig - cap_w, cap_w * sizeof(gl), cudaMemcpyDeviceToDevice, st));
    return TMX_OK;
}

// one device -> host copy for all first-round caps and the range-check verdict; the caps enter the proof and the transcript
int Prover::finish_round1(tmx_ctx* ctx, Challenger& ch, std::vector<gl>& proof, cudaStream_t st, bool* range_ok) {
    (void)ctx;
    TMX_CUDA(cudaMemcpyAsync(d_small + 66 * (STARK_N_TABLES - 1) + 65, d_hist + BUS_HIST_SIZE, sizeof(unsigned int), cudaMemcpyDeviceToDevice, st));
    std::vector<gl> all;
    int rc = d2h(all, d_small, (size_t)STARK_N_TABLES * 66, st);
    if (rc) return rc;
    *range_ok = (uint32_t)all[66 * (STARK_N_TABLES - 1) + 65] == 0;
    for (int t = 0; t < STARK_N_TABLES; t++) {
        const TableDef& td = def->tables[t];
        if (!td.n_main) continue;
        const size_t cap_w = 4 * ((size_t)1 << cap_height_of(td.log_n + STARK_RATE_BITS));
        proof.insert(proof.end(), all.begin() + 66 * t, all.begin() + 66 * t + cap_w);
        ch.observe(all.data() + 66 * t, cap_w);
    }
    // the commitment stamps are complete: LDE / Merkle device time per table
    for (int t = 0; t < STARK_N_TABLES; t++) {
        if (!def->tables[t].n_main) continue;
        cudaEventElapsedTime(&lde_ms[t], ev_phase[3 * t], ev_phase[3 * t + 1]);
        cudaEventElapsedTime(&merkle_ms[t], ev_phase[3 * t + 1], ev_phase[3 * t + 2]);
    }
    return TMX_OK;
}

int Prover::commit_aux(tmx_ctx* ctx, int t, const gl* d_trace, gl2 beta, gl2 gamma, cudaStream_t st) {
    const TableDef& td = def->tables[t];
    if (!td.n_main) return TMX_OK;
    TableDevice& tb = tab[t];
    const size_t n = td.rows(), m = n << STARK_RATE_BITS, A = (size_t)td.n_aux(), H = td.n_helpers;
    const unsigned km = td.log_n + STARK_RATE_BITS;
    const size_t dig = tmx_merkle_digest_count(km, cap_height_of(km)), cap_w = 4 * ((size_t)1 << cap_height_of(km));
    BusPassArgs a = bus_pass_args(*this, t, d_trace);
    a.beta = beta;
    a.gamma = gamma;
    a.aux = tb.d_aux;
    a.rowsum = d_rowsum + (size_t)t * rowsum_stride;
    int rc = launch_bus_gen(ctx, t, a, st);
    if (rc) return rc;
    bus_scan_kernel<<<1, 1024, 0, st>>>(a.rowsum, n, tb.d_aux + (2 * H) * n, tb.d_aux + (2 * H + 1) * n, d_small + 66 * t + 64);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    if ((rc = tmx_lde(ctx, tb.d_aux, tb.d_lde_a, tb.d_coef_a, A, td.log_n, STARK_RATE_BITS, st))) return rc;
    if ((rc = merkle_generic(ctx, tb.d_lde_a, A, 1, m, km, cap_height_of(km), tb.d_dig_a, st))) return rc;
    TMX_CUDA(cudaMemcpyAsync(d_small + 66 * t, tb.d_dig_a + 4 * dig - cap_w, cap_w * sizeof(gl), cudaMemcpyDeviceToDevice, st));
    return TMX_OK;
}

int Prover::finish_round2(tmx_ctx* ctx, Challenger& ch, std::vector<gl>& proof, cudaStream_t st) {
    (void)ctx;
    std::vector<gl> all;
    int rc = d2h(all, d_small, (size_t)STARK_N_TABLES * 66, st);
    if (rc) return rc;
    for (int t = 0; t < STARK_N_TABLES; t++) {
        const TableDef& td = def->tables[t];
        if (!td.n_main) continue;
        const size_t cap_w = 4 * ((size_t)1 << cap_height_of(td.log_n + STARK_RATE_BITS));
        proof.insert(proof.end(), all.begin() + 66 * t, all.begin() + 66 * t + cap_w);
        tab[t].total = gl2_make(all[66 * t + 64], all[66 * t + 65]);
        proof.push_back(tab[t].total.a0);
        proof.push_back(tab[t].total.a1);
        ch.observe(all.data() + 66 * t, cap_w);
        ch.observe_ext(tab[t].total);
    }
    return TMX_OK;
}

int Prover::prove_tail(tmx_ctx* ctx, int table, gl2 beta, gl2 gamma, Challenger ch, std::vector<gl>& proof, cudaStream_t st) {
    const TableDef& td = def->tables[table];
    if (!td.n_main) return TMX_OK;
    TableDevice& tb = tab[table];
    const unsigned log_n = td.log_n, km = log_n + STARK_RATE_BITS;
    const size_t n = td.rows(), m = n << STARK_RATE_BITS;
    const size_t Kc = td.n_const, C = td.n_main, A = (size_t)td.n_aux(), CT = Kc + C + A;
    const unsigned cap_h = cap_height_of(km);
    const size_t cap_n = (size_t)1 << cap_h;
    const size_t dig_t = tmx_merkle_digest_count(km, cap_h);
    int rc;
    PhaseTimer pt(st);
    Tail& tl = tail[table];
    gl* const d_dig_q = tl.d_dig_q; gl* const d_dig_fri = tl.d_dig_fri; gl* const d_qv = tl.d_qv; gl* const d_qcoef = tl.d_qcoef;
    gl* const d_qlde = tl.d_qlde; gl2* const d_ypa = tl.d_ypa; gl2* const d_ypb = tl.d_ypb; gl2* const d_open = tl.d_open;
    gl2* const d_apow = tl.d_apow; uint32_t* const d_idx = tl.d_idx; gl2* const* d_fri = tl.d_fri;
    // ---- constraint challenges, quotient ----
    QuotientArgs qa;
    memset(&qa, 0, sizeof qa);
    qa.lde_m = tb.d_lde_m; qa.lde_k = tb.d_const_lde; qa.lde_a = tb.d_lde_a;
    qa.m = m; qa.log_m = km; qa.rate_bits = STARK_RATE_BITS;
    qa.pertab = tb.d_per_lde; qa.P = (int)td.period;
    qa.shape = shape;
    qa.alpha[0] = ch.get();
    qa.alpha[1] = ch.get();
    qa.beta = beta; qa.gamma = gamma;
    qa.s_over_n = gl2_scale(tb.total, gl_inv((gl)n));
    fill_zh_inv(qa, n);
    qa.out = d_qv;
    const unsigned qblocks = (unsigned)((m + 127) / 128);
    rc = launch_quotient(ctx, table, qa, st);
    if (rc) return rc;
    pt.tick("quotient kernel");
    // values on the coset (natural order) -> coefficients of Q(7 X); the 1/m factor comes with the inverse NTT
    rc = ntt_with_scratch(ctx, d_qv, 2, km, true, tl.d_ntt_tmp, st);
    if (rc) return rc;
    quotient_chunks_kernel<<<(unsigned)((4 * n + 255) / 256), 256, 0, st>>>(d_qv, n, gl_inv(gl_pow(GL_GEN, n)), d_qcoef);
    ctx->launches++;
    rc = lde_forward_cosets(ctx, d_qcoef, d_qlde, 4, log_n, STARK_RATE_BITS, st);
    if (rc) return rc;
    rc = merkle_generic(ctx, d_qlde, 4, 1, m, km, cap_h, d_dig_q, st);
    if (rc) return rc;
    std::vector<gl> cap;
    rc = d2h(cap, d_dig_q + 4 * (dig_t - cap_n), 4 * cap_n, st, table);
    if (rc) return rc;
    proof.insert(proof.end(), cap.begin(), cap.end());
    ch.observe(cap.data(), cap.size());
    pt.tick("quotient commit");
    // ---- openings at zeta and g * zeta (coefficients are stored coset-scaled: evaluate at zeta / 7) ----
    const gl2 zeta = ch.get_ext();
    const gl2 zeta_next = gl2_scale(zeta, gl_root_of_unity(log_n));
    const gl ginv = gl_inv(GL_GEN);
    ext_powers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gl2_scale(zeta, ginv), n, d_ypa);
    ext_powers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gl2_scale(zeta_next, ginv), n, d_ypb);
    ctx->launches += 2;
    {   // d_open: local[CT], next[CT], quotient[4]
        const gl* seg[3] = {tb.d_const_coef, tb.d_coef_m, tb.d_coef_a};
        const size_t segc[3] = {Kc, C, A};
        size_t off = 0;
        for (int s = 0; s < 3; s++) {
            if (segc[s]) {
                eval_columns_kernel<<<(unsigned)segc[s], 256, 0, st>>>(seg[s], n, d_ypa, d_ypb, d_open + off, d_open + CT + off);
                ctx->launches++;
            }
            off += segc[s];
        }
        eval_columns_kernel<<<4, 256, 0, st>>>(d_qcoef, n, d_ypa, nullptr, d_open + 2 * CT, nullptr);
        ctx->launches++;
    }
    TMX_CUDA(cudaGetLastError());
    pt.tick("openings kernels");
    std::vector<gl> op;
    rc = d2h(op, reinterpret_cast<const gl*>(d_open), 2 * (2 * CT + 4), st, table);
    if (rc) return rc;
    proof.insert(proof.end(), op.begin(), op.end());  // local[CT], next[CT], quotient[4] as (a0, a1) pairs
    auto ext_at = [&](size_t i) { return gl2_make(op[2 * i], op[2 * i + 1]); };
    for (size_t c = 0; c < CT; c++) ch.observe_ext(ext_at(c));
    for (size_t q = 0; q < 4; q++) ch.observe_ext(ext_at(2 * CT + q));
    for (size_t c = 0; c < CT; c++) ch.observe_ext(ext_at(CT + c));
    pt.tick("openings d2h+observe");
    // ---- FRI batch polynomial, pointwise ----
    const gl2 fa = ch.get_ext();
    std::vector<gl2> apow(CT + 4);
    apow[0] = gl2_from(1);
    for (size_t j = 1; j < CT + 4; j++) apow[j] = gl2_mul(apow[j - 1], fa);
    FriBatchArgs fb;
    memset(&fb, 0, sizeof fb);
    fb.red0 = gl2_from(0);
    fb.red1 = gl2_from(0);
    for (size_t j = 0; j < CT; j++) {
        fb.red0 = gl2_add(fb.red0, gl2_mul(apow[j], ext_at(j)));
        fb.red1 = gl2_add(fb.red1, gl2_mul(apow[j], ext_at(CT + j)));
    }
    for (size_t q = 0; q < 4; q++) fb.red0 = gl2_add(fb.red0, gl2_mul(apow[CT + q], ext_at(2 * CT + q)));
    TMX_CUDA(cudaMemcpyAsync(d_apow, apow.data(), apow.size() * sizeof(gl2), cudaMemcpyHostToDevice, st));
    fb.seg[0] = tb.d_const_lde; fb.seg[1] = tb.d_lde_m; fb.seg[2] = tb.d_lde_a;
    fb.segc[0] = Kc; fb.segc[1] = C; fb.segc[2] = A;
    fb.lde_q = d_qlde; fb.C = CT; fb.m = m; fb.log_m = km; fb.apow = d_apow;
    fb.zeta = zeta; fb.zeta_next = zeta_next; fb.alpha_c = gl2_mul(apow[CT - 1], fa);
    fb.w_m = gl_root_of_unity(km);
    fb.out = d_fri[0];
    fri_batch_kernel<<<qblocks, 128, 0, st>>>(fb);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    pt.tick("fri batch");
    // ---- FRI commit phase: Merkle over cosets of 16, fold with beta ----
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
        rc = d2h(cap, dg + 4 * (ldig - ((size_t)1 << lcap_h)), 4 * ((size_t)1 << lcap_h), st, table);
        if (rc) return rc;
        proof.insert(proof.end(), cap.begin(), cap.end());
        ch.observe(cap.data(), cap.size());
        const gl2 fbeta = ch.get_ext();
        const size_t cosets = cur >> STARK_ARITY_BITS;
        fri_fold_kernel<<<(unsigned)((cosets + 127) / 128), 128, 0, st>>>(d_fri[l], cosets, lg_rows, gl_inv(shift),
                                                                         gl_inv(gl_root_of_unity(ilog2(cur))), fbeta, d_fri[l + 1]);
        ctx->launches++;
        TMX_CUDA(cudaGetLastError());
        cur = cosets;
        shift = gl_pow(shift, 16);
    }
    pt.tick("fri layers");
    // final polynomial: the remaining `cur` evaluations (bit-reversed) on shift * <w_cur> -> coefficients
    std::vector<gl> fin;
    rc = d2h(fin, reinterpret_cast<const gl*>(d_fri[n_layers]), 2 * cur, st, table);
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
            if (coef[k].a0 || coef[k].a1) return fail(TMX_E_CUDA, "FRI: final polynomial exceeds its degree bound (the witness does not satisfy the constraints of table " + std::to_string(table) + ")");
        proof.push_back((gl)final_len);
        for (size_t k = 0; k < final_len; k++) {
            proof.push_back(coef[k].a0);
            proof.push_back(coef[k].a1);
            ch.observe_ext(coef[k]);
        }
    }
    pt.tick("final poly");
    // ---- proof of work (K9) ----
    {
        gl state[12];
        for (int i = 0; i < 12; i++) state[i] = ch.state[i];
        for (int i = 0; i < ch.n_in; i++) state[i] = ch.in[i];
        uint64_t wit = 0;
        rc = pow_grind(ctx, state, ch.n_in, STARK_POW_BITS, &wit, tl.d_pow, st);
        if (rc) return rc;
        ch.observe((gl)wit);
        (void)ch.get();
        proof.push_back((gl)wit);
    }
    pt.tick("pow");
    // ---- queries (K4) ----
    std::vector<uint32_t> idx(STARK_NUM_QUERIES);
    for (int q = 0; q < STARK_NUM_QUERIES; q++) idx[q] = (uint32_t)(ch.get() % m);
    TMX_CUDA(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    const unsigned n_sib = km - cap_h;
    size_t qlen = (Kc ? Kc + 4 * n_sib : 0) + C + 4 * n_sib + A + 4 * n_sib + 4 + 4 * n_sib;
    for (unsigned l = 0; l < n_layers; l++)
        qlen += 32 + 4 * (layer_rows_log[l] - std::min<unsigned>((unsigned)layer_rows_log[l], STARK_CAP_HEIGHT));
    if (tl.sz_query < qlen * STARK_NUM_QUERIES) {
        if (tl.d_query) TMX_CUDA(cudaFree(tl.d_query));
        tl.d_query = nullptr;
        tl.sz_query = 0;
        TMX_CUDA(cudaMalloc((void**)&tl.d_query, qlen * STARK_NUM_QUERIES * sizeof(gl)));
        tl.sz_query = qlen * STARK_NUM_QUERIES;
    }
    gl* const d_query = tl.d_query;
    size_t off = 0;
    const unsigned NQ = STARK_NUM_QUERIES;
    auto open_tree = [&](const gl* lde, size_t cols, const gl* dig) {
        gather_rows_kernel<<<NQ, 256, 0, st>>>(lde, cols, m, d_idx, 0, d_query, qlen, off);
        off += cols;
        gather_paths_kernel<<<NQ, 32, 0, st>>>(dig, km, n_sib, d_idx, 0, d_query, qlen, off);
        off += 4 * n_sib;
        ctx->launches += 2;
    };
    if (Kc) open_tree(tb.d_const_lde, Kc, tb.d_const_dig);
    open_tree(tb.d_lde_m, C, tb.d_dig_m);
    open_tree(tb.d_lde_a, A, tb.d_dig_a);
    open_tree(d_qlde, 4, d_dig_q);
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
    rc = d2h(qd, d_query, qlen * NQ, st, table);
    if (rc) return rc;
    proof.insert(proof.end(), qd.begin(), qd.end());
    pt.tick("queries");
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;

// K5 as a kernel-level entry point (parity tests, ncu): constraint quotient of one table on its LDE coset.  d_lde_main /
// d_lde_aux: LDEs ([cols][2n], bit-reversed rows, as tmx_lde produces them) of the first- and second-round traces; total:
// the table's bus total; d_out: [2][2n] in NATURAL order, one row per constraint challenge alpha[i]:
// sum_k alpha^(M-1-k) C_k(x) / (x^n - 1).  The constant and periodic columns are the circuit's.
int tmx::stark_quotient(tmx_ctx* ctx, Prover& pr, int table, const gl* d_lde_main, const gl* d_lde_aux, gl2 total, gl2 beta, gl2 gamma,
                        const gl alpha[2], gl* d_out, cudaStream_t st) {
    const TableDef& td = pr.def->tables[table];
    const size_t n = td.rows();
    QuotientArgs qa;
    memset(&qa, 0, sizeof qa);
    qa.lde_m = d_lde_main; qa.lde_k = pr.tab[table].d_const_lde; qa.lde_a = d_lde_aux;
    qa.m = n << STARK_RATE_BITS; qa.log_m = td.log_n + STARK_RATE_BITS; qa.rate_bits = STARK_RATE_BITS;
    qa.pertab = pr.tab[table].d_per_lde; qa.P = (int)td.period;
    qa.shape = pr.shape;
    qa.alpha[0] = alpha[0]; qa.alpha[1] = alpha[1];
    qa.beta = beta; qa.gamma = gamma;
    qa.s_over_n = gl2_scale(total, gl_inv((gl)n));
    fill_zh_inv(qa, n);
    qa.out = d_out;
    return launch_quotient(ctx, table, qa, st);
}

// K3 fold as a kernel-level entry point: one arity-16 FRI folding step in evaluation space.  d_in: 16 * n_cosets extension
// elements (interleaved (a0, a1), bit-reversed order) on the coset shift * <w>, d_out: n_cosets folded values (bit-reversed)
// on shift^16 * <w^16>.  [plonky2 fri/prover.rs fri_committed_trees, one layer]
extern "C" int tmx_fri_fold(tmx_ctx* ctx, const uint64_t* d_in, unsigned log_cosets, uint64_t shift, const uint64_t beta[2],
                            uint64_t* d_out, void* stream) {
    if (!ctx || !d_in || !d_out || !beta || log_cosets > 26 || shift == 0) return fail(TMX_E_INPUT, "tmx_fri_fold: bad arguments");
    const size_t cosets = (size_t)1 << log_cosets;
    cudaStream_t st = pick_stream(ctx, stream);
    fri_fold_kernel<<<(unsigned)((cosets + 127) / 128), 128, 0, st>>>(reinterpret_cast<const gl2*>(d_in), cosets, log_cosets, gl_inv(shift),
                                                                     gl_inv(gl_root_of_unity(log_cosets + STARK_ARITY_BITS)),
                                                                     gl2_make(beta[0], beta[1]), reinterpret_cast<gl2*>(d_out));
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}
