// K1: batched radix-2 Goldilocks NTT / iNTT / coset low-degree extension for sm_100a.
//
// Replaces plonky2_field fft.rs (fft / ifft / coset_fft) and plonky2 fri/oracle.rs
// PolynomialBatch::{from_values, lde_values}, reached from the reference through `circuit.prove()`
// [REF circuits/skip.rs:214, circuits/step.rs:196].
//
// Design (not a port of the CPU algorithm): an N = 2^k transform is split into 1..3 passes; pass i covers
// l_i <= 10 butterfly stages of an in-place decimation-in-frequency network.  Each CTA stages a tile of
// 8 vectors x 2^l points in shared memory (point-major, stride 9 so both the strided and the transposing
// access patterns are bank-conflict free), runs the stages as radix-16 register blocks with twiddles from a
// shared-memory table, applies the inter-pass twiddle w_N^(lo * bitrev(p)) from a flat L2-resident table and
// writes the tile back in 64-byte (strided passes) or full-row (contiguous pass) segments.  In-place DIF leaves the result bit-reversed,
// which is exactly the Merkle-leaf order plonky2 wants for LDE values, so the forward transform never
// reorders; the inverse transform folds the bit-reversal, the 1/n factor and the coset shift 7^i into the
// scatter of its last pass.  HBM-shaped integer work that is in fact bound by 64-bit modular-arithmetic issue (see
// profiles/r1_ncu_summary.md): no tensor cores.
#include "ctx.cuh"
#include <algorithm>
#include <cstring>

namespace tmx {

// Vectors per tile.  Measured on the Ed25519 table's LDE (then 1217 x 2^16): 16 vectors 4.42 ms, 8 vectors 4.13 ms, 4 vectors
// 4.35 ms.  Narrower tiles mean smaller CTAs (128 threads, 19 KB of shared memory, < 60 registers): more CTAs per SM in
// different phases, so tile loads overlap better with butterflies; below 64-byte rows the global accesses get too short.
#ifndef TMX_NTT_TILE_LOG
#define TMX_NTT_TILE_LOG 3
#endif
constexpr int TILE_LOG = TMX_NTT_TILE_LOG;
constexpr int TILE_T = 1 << TILE_LOG;  // vectors per tile
constexpr int TILE_TS = TILE_T + 1;    // padded stride

struct PassArgs {
    const gl* in;
    gl* out;
    size_t in_col_stride, out_col_stride;
    size_t n_cols;
    unsigned log_n;      // column length 2^log_n
    unsigned log_block;  // this pass runs inside blocks of 2^log_block consecutive elements
    const gl* tw_small;  // w_L^e, e < L/2
    // inter-pass twiddle w_N^(e << tw_shift), flat table of N = 2^log_n entries
    const gl* tw;
    unsigned tw_shift;
    // optional pre-scale of element i by ps[(ps_mult * i) & ps_mask] (coset selection of the LDE)
    const gl* ps;
    uint64_t ps_mult, ps_mask;
    // contiguous pass only: scatter to natural order (bit reversal of the in-place position) and optional
    // post-scale of natural index j by os[j]
    int scatter_natural;
    const gl* os;
};

// r consecutive DIF stages starting at `stage0` of an L-point vector held point-major in shared memory.
// Twiddles equal to one (the whole last stage of a transform, and the first butterfly of every group in the last
// round) are not multiplied.
template <int LOG_L, int R_LOG, int STAGE0>
TMX_D void dif_round(gl* tile, const gl* tws, int tid, int nthreads) {
    constexpr int L = 1 << LOG_L;
    constexpr int R = 1 << R_LOG;
    constexpr int rs = L >> (STAGE0 + R_LOG);  // stride between the R points of one register block
    constexpr int total = (L / R) * TILE_T;
    for (int g = tid; g < total; g += nthreads) {
        const int v = g % TILE_T;
        const int gi = g / TILE_T;
        const int lo = gi % rs;
        const int hi = gi / rs;
        const int base = hi * R * rs + lo;
        gl x[R];
#pragma unroll
        for (int m = 0; m < R; m++) x[m] = tile[(base + m * rs) * TILE_TS + v];
#pragma unroll
        for (int ss = 0; ss < R_LOG; ss++) {
            const int hm = R >> (ss + 1);
            const int tw_step = (L / 2) / (hm * rs);  // w_{2h} = w_L^tw_step, h = hm*rs
#pragma unroll
            for (int m = 0; m < R; m++) {
                if ((m & hm) == 0) {
                    const gl a = x[m], b = x[m + hm];
                    x[m] = gl_add(a, b);
                    const gl d = gl_sub(a, b);
                    if (rs == 1 && (m % hm) == 0) x[m + hm] = d;  // exponent 0 at compile time
                    else x[m + hm] = gl_mul(d, tws[((m % hm) * rs + lo) * tw_step]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < R; m++) tile[(base + m * rs) * TILE_TS + v] = x[m];
    }
}

template <int LOG_L>
TMX_D void dif_vector(gl* tile, const gl* tws, int tid, int nthreads) {
    constexpr int full = LOG_L / 4, rem = LOG_L % 4;
    // the short round goes first so that the last round (stride 1, compile-time unit twiddles) is a full radix-16
    if constexpr (rem == 3) dif_round<LOG_L, 3, 0>(tile, tws, tid, nthreads);
    if constexpr (rem == 2) dif_round<LOG_L, 2, 0>(tile, tws, tid, nthreads);
    if constexpr (rem == 1) dif_round<LOG_L, 1, 0>(tile, tws, tid, nthreads);
    if constexpr (rem != 0) __syncthreads();
    if constexpr (full >= 1) {
        dif_round<LOG_L, 4, rem>(tile, tws, tid, nthreads);
        __syncthreads();
    }
    if constexpr (full >= 2) {
        dif_round<LOG_L, 4, rem + 4>(tile, tws, tid, nthreads);
        __syncthreads();
    }
}

template <int LOG_L>
constexpr int pass_threads() {
    // one thread per radix-16 register block of the tile
    constexpr int t = ((1 << LOG_L) * TILE_T) / 16;
    return t < 32 ? 32 : (t > 512 ? 512 : t);
}

// All sizes are powers of two: every index split below is a shift / mask (the first version divided 64-bit indices
// by runtime values, which cost more instructions than the butterflies).
template <int LOG_L, bool STRIDED>
__global__ void __launch_bounds__(pass_threads<LOG_L>()) ntt_pass_kernel(PassArgs a) {
    constexpr int L = 1 << LOG_L;
    extern __shared__ gl sm[];
    gl* tile = sm;
    gl* tws = sm + L * TILE_TS;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    for (int i = tid; i < L / 2; i += nthreads) tws[i] = a.tw_small[i];

    if constexpr (STRIDED) {
        const unsigned log_s = a.log_block - LOG_L;  // >= TILE_LOG: S = 2^log_s elements between the points of a vector
        const size_t S = (size_t)1 << log_s;
        const unsigned log_tiles = log_s - TILE_LOG;                  // tiles per block
        const unsigned log_blocks = a.log_n - a.log_block;     // blocks per column
        const size_t t = blockIdx.x;
        const size_t lo0 = (t & (((size_t)1 << log_tiles) - 1)) << TILE_LOG;
        const size_t hi = (t >> log_tiles) & (((size_t)1 << log_blocks) - 1);
        const size_t col = t >> (log_tiles + log_blocks);
        const size_t off = (hi << a.log_block) + lo0;
        const gl* src = a.in + col * a.in_col_stride + off;
        gl* dst = a.out + col * a.out_col_stride + off;
        for (int idx = tid; idx < L * TILE_T; idx += nthreads) {
            const int v = idx & (TILE_T - 1), m = idx >> TILE_LOG;
            gl x = src[((size_t)m << log_s) + v];
            if (a.ps) {
                const uint64_t i = off + ((uint64_t)m << log_s) + v;
                x = gl_mul(x, a.ps[(i * a.ps_mult) & a.ps_mask]);
            }
            tile[m * TILE_TS + v] = x;
        }
        __syncthreads();
        dif_vector<LOG_L>(tile, tws, tid, nthreads);
        for (int idx = tid; idx < L * TILE_T; idx += nthreads) {
            const int v = idx & (TILE_T - 1), p = idx >> TILE_LOG;
            gl x = tile[p * TILE_TS + v];
            const uint32_t j1 = bitrev32((uint32_t)p, LOG_L);
            if (j1) x = gl_mul(x, a.tw[((lo0 + v) * j1) << a.tw_shift]);
            dst[((size_t)p << log_s) + v] = x;
        }
    } else {
        const unsigned log_vecs = a.log_n - LOG_L;  // vectors per column
        const size_t vec_mask = ((size_t)1 << log_vecs) - 1;
        const size_t total_vecs = a.n_cols << log_vecs;
        const size_t gid0 = (size_t)blockIdx.x * TILE_T;
        for (int idx = tid; idx < L * TILE_T; idx += nthreads) {
            const int m = idx & (L - 1), v = idx >> LOG_L;
            const size_t gid = gid0 + v;
            gl x = 0;
            if (gid < total_vecs) {
                const size_t col = gid >> log_vecs, c = gid & vec_mask;
                const size_t hi = a.scatter_natural ? bitrev32((uint32_t)c, log_vecs) : c;
                const uint64_t i = (hi << LOG_L) + m;
                x = a.in[col * a.in_col_stride + i];
                if (a.ps) x = gl_mul(x, a.ps[(i * a.ps_mult) & a.ps_mask]);
            }
            tile[m * TILE_TS + v] = x;
        }
        __syncthreads();
        dif_vector<LOG_L>(tile, tws, tid, nthreads);
        if (a.scatter_natural) {
            for (int idx = tid; idx < L * TILE_T; idx += nthreads) {
                const int v = idx & (TILE_T - 1), p = idx >> TILE_LOG;
                const size_t gid = gid0 + v;
                if (gid < total_vecs) {
                    const size_t col = gid >> log_vecs, c = gid & vec_mask;
                    const uint64_t j = ((uint64_t)bitrev32((uint32_t)p, LOG_L) << log_vecs) + c;
                    gl x = tile[p * TILE_TS + v];
                    if (a.os) x = gl_mul(x, a.os[j]);
                    a.out[col * a.out_col_stride + j] = x;
                }
            }
        } else {
            for (int idx = tid; idx < L * TILE_T; idx += nthreads) {
                const int p = idx & (L - 1), v = idx >> LOG_L;
                const size_t gid = gid0 + v;
                if (gid < total_vecs) {
                    const size_t col = gid >> log_vecs, c = gid & vec_mask;
                    a.out[col * a.out_col_stride + (c << LOG_L) + p] = tile[p * TILE_TS + v];
                }
            }
        }
    }
}

template <int LOG_L, bool STRIDED>
static int launch_pass_t(tmx_ctx* ctx, const PassArgs& a, cudaStream_t st) {
    constexpr int L = 1 << LOG_L;
    const size_t smem = (size_t)L * TILE_TS * sizeof(gl) + (L / 2 + 1) * sizeof(gl);
    auto kern = ntt_pass_kernel<LOG_L, STRIDED>;
    if (smem > 48 * 1024)
        TMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t n = (size_t)1 << a.log_n;
    size_t grid;
    if (STRIDED)
        grid = a.n_cols * (n >> LOG_L) / TILE_T;
    else
        grid = ((n >> LOG_L) * a.n_cols + TILE_T - 1) / TILE_T;
    kern<<<(unsigned)grid, pass_threads<LOG_L>(), smem, st>>>(a);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

template <bool STRIDED>
static int launch_pass(tmx_ctx* ctx, unsigned log_l, const PassArgs& a, cudaStream_t st) {
    switch (log_l) {
#define CASE(k) \
    case k:     \
        return launch_pass_t<k, STRIDED>(ctx, a, st);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
    }
    return fail(TMX_E_INPUT, "ntt: unsupported pass size");
}

// split k stages into passes; later passes get the larger share and the last one (contiguous) is never smaller
// than 16 points unless the whole transform is.  Passes of 10 stages (1024-point tiles, 139 KB of shared memory: one
// CTA per SM) measured slower than one more sweep over the table: 2^20 ran at 580 GB/s as {10, 10} against 720 GB/s
// for 2^21 as {7, 7, 7}; so a pass is at most 9 stages once there is more than one.
static std::vector<unsigned> plan_passes(unsigned k) {
    std::vector<unsigned> p;
    if (k == 0) return p;
    unsigned np = k <= 10 ? 1 : (k + 8) / 9;
    unsigned left = k;
    for (unsigned i = 0; i < np; i++) {
        unsigned l = left / (np - i);
        p.push_back(l);
        left -= l;
    }
    return p;  // ascending: e.g. 17 -> {8, 9}
}

// One full transform of n_cols columns.  in -> out, bit-reversed (in-place order) or natural (scatter).
// tmp: buffer with the layout of `out` used when the last pass scatters (may alias out otherwise).
struct XformDesc {
    const gl* in;
    size_t in_col_stride;
    gl* out;
    size_t out_col_stride;
    gl* tmp;  // same strides as out; required when natural && passes > 1 ... see below
    size_t tmp_col_stride;
    size_t n_cols;
    unsigned log_n;
    bool inverse;
    bool natural_out;
    const gl* ps;  // pre-scale table (first pass): element i times ps[(ps_mult * i) & ps_mask]
    uint64_t ps_mult, ps_mask;
    const gl* os;  // post-scale (last pass, natural_out only), indexed by the natural output index
};

static int run_xform(tmx_ctx* ctx, const XformDesc& d, cudaStream_t st) {
    const NttTables* T = nullptr;
    std::vector<unsigned> plan = plan_passes(d.log_n);
    if (plan.empty()) return fail(TMX_E_INPUT, "ntt: log_n must be >= 1");
    const NttTables* Tn = nullptr;
    int rc = ctx_ntt_tables(ctx, d.log_n, d.inverse, &Tn);
    if (rc) return rc;
    unsigned log_block = d.log_n;
    const gl* cur_in = d.in;
    size_t cur_in_stride = d.in_col_stride;
    for (size_t i = 0; i < plan.size(); i++) {
        const bool last = (i + 1 == plan.size());
        const unsigned l = plan[i];
        rc = ctx_ntt_tables(ctx, l, d.inverse, &T);
        if (rc) return rc;
        PassArgs a;
        memset(&a, 0, sizeof a);
        a.in = cur_in;
        a.in_col_stride = cur_in_stride;
        a.n_cols = d.n_cols;
        a.log_n = d.log_n;
        a.log_block = log_block;
        a.tw_small = T->small;
        a.tw = Tn->full;
        a.tw_shift = d.log_n - log_block;
        if (i == 0 && d.ps) {
            a.ps = d.ps;
            a.ps_mult = d.ps_mult;
            a.ps_mask = d.ps_mask;
        }
        if (last) {
            a.scatter_natural = d.natural_out ? 1 : 0;
            if (d.natural_out && d.os) a.os = d.os;
            a.out = d.out;
            a.out_col_stride = d.out_col_stride;
            if (d.natural_out && (const gl*)a.out == a.in) return fail(TMX_E_INPUT, "ntt: scatter pass cannot run in place");
            rc = launch_pass<false>(ctx, l, a, st);
        } else {
            // intermediate results: natural_out needs a separate buffer for the final scatter
            gl* mid = d.natural_out ? d.tmp : d.out;
            size_t mid_stride = d.natural_out ? d.tmp_col_stride : d.out_col_stride;
            a.out = mid;
            a.out_col_stride = mid_stride;
            rc = launch_pass<true>(ctx, l, a, st);
            cur_in = mid;
            cur_in_stride = mid_stride;
        }
        if (rc) return rc;
        log_block -= l;
    }
    return TMX_OK;
}

__global__ void scale_kernel(gl* p, size_t n, gl s) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = gl_mul(p[i], s);
}
static int launch_scale(tmx_ctx* ctx, gl* p, size_t n, gl s, cudaStream_t st) {
    size_t blocks = std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n, s);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;

int tmx::ntt_with_scratch(tmx_ctx* ctx, gl* d_data, size_t n_cols, unsigned log_n, bool inverse, gl* d_tmp, cudaStream_t st) {
    const size_t n = (size_t)1 << log_n;
    const bool multi = plan_passes(log_n).size() > 1;
    XformDesc d;
    memset(&d, 0, sizeof d);
    d.n_cols = n_cols;
    d.log_n = log_n;
    d.inverse = inverse;
    d.natural_out = true;
    d.in_col_stride = n;
    d.out = d_data;
    d.out_col_stride = n;
    if (multi) {
        // data -> tmp (strided passes) -> data (scatter): the scatter pass reads tmp and writes data, never aliased
        d.in = d_data;
        d.tmp = d_tmp;
        d.tmp_col_stride = n;
    } else {
        TMX_CUDA(cudaMemcpyAsync(d_tmp, d_data, n * n_cols * sizeof(gl), cudaMemcpyDeviceToDevice, st));
        d.in = d_tmp;
    }
    int rc = run_xform(ctx, d, st);
    if (rc) return rc;
    if (inverse) rc = launch_scale(ctx, d_data, n * n_cols, gl_inv((gl)n), st);
    return rc;
}

extern "C" int tmx_ntt(tmx_ctx* ctx, uint64_t* d_data, size_t n_cols, unsigned log_n, int inverse, void* stream) {
    if (!ctx || !d_data || log_n < 1 || log_n > 30) return fail(TMX_E_INPUT, "tmx_ntt: bad arguments");
    if (n_cols == 0) return TMX_OK;
    void* tmp = nullptr;
    int rc = ctx_scratch(ctx, 0, ((size_t)n_cols << log_n) * sizeof(gl), &tmp);
    if (rc) return rc;
    return ntt_with_scratch(ctx, d_data, n_cols, log_n, inverse != 0, (gl*)tmp, pick_stream(ctx, stream));
}

namespace tmx {
// coset-scaled coefficients c_i * 7^i ([n_cols][n], natural order) -> evaluations on the 2^rate_bits cosets,
// out[n_cols][n << rate_bits] bit-reversed.  Region rho of a column holds coset s = bitrev_r(rho), whose input is
// pre-scaled by w_{n*2^r}^(s*i); the DIF passes leave it bit-reversed in place.
int lde_forward_cosets(tmx_ctx* ctx, const gl* coeffs, gl* d_out, size_t n_cols, unsigned log_n, unsigned rate_bits,
                       cudaStream_t st) {
    const size_t n = (size_t)1 << log_n;
    const size_t m = n << rate_bits;
    const NttTables* Tm = nullptr;
    int rc = ctx_ntt_tables(ctx, log_n + rate_bits, false, &Tm);
    if (rc) return rc;
    for (unsigned rho = 0; rho < (1u << rate_bits); rho++) {
        const unsigned s = bitrev32(rho, rate_bits);
        XformDesc f;
        memset(&f, 0, sizeof f);
        f.in = coeffs;
        f.in_col_stride = n;
        f.out = d_out + (size_t)rho * n;
        f.out_col_stride = m;
        f.n_cols = n_cols;
        f.log_n = log_n;
        f.inverse = false;
        f.natural_out = false;
        if (s != 0) {
            f.ps = Tm->full;
            f.ps_mult = s;
            f.ps_mask = m - 1;
        }
        rc = run_xform(ctx, f, st);
        if (rc) return rc;
    }
    return TMX_OK;
}
}  // namespace tmx

extern "C" int tmx_lde(tmx_ctx* ctx, const uint64_t* d_values, uint64_t* d_out, uint64_t* d_coeffs, size_t n_cols,
                       unsigned log_n, unsigned rate_bits, void* stream) {
    if (!ctx || !d_values || !d_out || log_n < 1 || log_n + rate_bits > 30 || rate_bits > 4)
        return fail(TMX_E_INPUT, "tmx_lde: bad arguments");
    if (n_cols == 0) return TMX_OK;
    cudaStream_t st = pick_stream(ctx, stream);
    const size_t n = (size_t)1 << log_n;
    const size_t m = n << rate_bits;
    gl* coeffs = d_coeffs;
    if (!coeffs) {
        void* p = nullptr;
        int rc = ctx_scratch(ctx, 0, n * n_cols * sizeof(gl), &p);
        if (rc) return rc;
        coeffs = (gl*)p;
    }
    const gl* cs = nullptr;
    int rc = ctx_coset_scale(ctx, log_n, &cs);
    if (rc) return rc;
    // 1. inverse transform: values -> (region 0 of out as intermediate) -> coeffs, natural order,
    //    scaled by 7^i / n
    XformDesc d;
    memset(&d, 0, sizeof d);
    d.in = d_values;
    d.in_col_stride = n;
    d.out = coeffs;
    d.out_col_stride = n;
    d.tmp = d_out;
    d.tmp_col_stride = m;
    d.n_cols = n_cols;
    d.log_n = log_n;
    d.inverse = true;
    d.natural_out = true;
    d.os = cs;
    rc = run_xform(ctx, d, st);
    if (rc) return rc;
    return lde_forward_cosets(ctx, coeffs, d_out, n_cols, log_n, rate_bits, st);
}
