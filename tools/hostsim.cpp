// CPU simulation of the witness kernels' per-thread logic (witness_jobs.cuh / witness.cuh compiled for the
// host).  TEST TOOL ONLY: lets tests/test_hostsim.py compare the kernel logic with the CPU oracle in the
// authoring container, which has no GPU.  It is not part of libtmx.so and nothing in the product calls it.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../tendermintx_b200/csrc/witness_jobs.cuh"
#include "../tendermintx_b200/csrc/stark_rows.cuh"

using namespace tmx;

static char g_chain[64] = "mocha-4";  // chain id of the circuits the tests simulate
extern "C" void hostsim_set_chain(const char* c) { strncpy(g_chain, c, 50); g_chain[50] = 0; }
static const uint64_t* g_plan = nullptr;
static size_t g_plan_rows = 0;
extern "C" void hostsim_set_plan(const uint64_t* k, size_t rows) { g_plan = k; g_plan_rows = rows; }
namespace tmx {
uint64_t logic_const_value(int kc, size_t row, AirShape) { return g_plan ? g_plan[(size_t)kc * g_plan_rows + row] : 0; }
gl2 logic_public_terms(AirShape, uint64_t, const uint8_t*, const uint8_t*, gl2, gl2) { return gl2_from(0); }
}

static unsigned log2u(uint32_t x) {
    unsigned k = 0;
    while ((1u << k) < x) k++;
    return k;
}

extern "C" int hostsim_build_traces(const uint8_t* blob, uint64_t* t256, size_t n256, uint64_t* t512, size_t n512,
                                    uint64_t* ted, size_t ned, uint8_t* aux) {
    const tmx_offchain_head* h = blob_head(blob);
    WitnessArgs a;
    a.blob = blob;
    a.kind = h->kind;
    a.n_max = h->n_max;
    a.np = 1;
    while (a.np < a.n_max) a.np *= 2;
    a.log_np = log2u(a.np);
    a.t256 = t256; a.n256 = n256; a.t512 = t512; a.n512 = n512; a.ted = ted; a.ned = ned;
    std::vector<uint8_t> nodes((size_t)2 * 2 * a.np * 32), en((size_t)2 * 2 * a.np);
    a.nodes = nodes.data();
    a.node_en = en.data();
    a.aux = aux;
    Sha256Hist hs[2];
    // leaves
    for (uint32_t s = 0; s < n_sets(a.kind); s++)
        for (uint32_t i = 0; i < a.np; i++)
            if (sha256_leaf_prepare(a, s, i, &hs[0]))
                for (int t = 0; t < 64; t++) sha256_row_cells(a.t256, a.n256, sha256_leaf_row0(a, s, i) + t, t, &hs[0]);
    // inner levels
    for (uint32_t l = 1; l <= a.log_np; l++)
        for (uint32_t s = 0; s < n_sets(a.kind); s++)
            for (uint32_t i = 0; i < (a.np >> l); i++) {
                sha256_inner_prepare(a, s, l, i, hs);
                for (int r = 0; r < 128; r++)
                    sha256_row_cells(a.t256, a.n256, sha256_inner_row0(a, s, l, i) + r, r & 63, &hs[r >> 6]);
            }
    // header proofs
    for (uint32_t k = 0; k < n_header_proofs(a.kind); k++) {
        HeaderProofDesc d;
        header_proof_desc(a, k, &d);
        uint8_t cur[32];
        size_t chunk = d.chunk0;
        for (int j = 0; j < 5; j++) {
            int nb = header_proof_prepare(d, j, cur, hs);
            for (int r = 0; r < nb * 64; r++) sha256_row_cells(a.t256, a.n256, chunk * 64 + r, r & 63, &hs[r >> 6]);
            chunk += nb;
        }
        memcpy(aux + AUX_PROOF_ROOT + 32 * k, cur, 32);
    }
    sha256_padding_prepare(&hs[0]);
    for (size_t c = sha256_used_chunks(a.kind, a.n_max, a.np); (c + 1) * 64 <= n256; c++)
        for (int t = 0; t < 64; t++) sha256_row_cells(a.t256, a.n256, c * 64 + t, t, &hs[0]);
    if (!t512 || !ted) return 0;
    // validator slots: SHA-512 + Ed25519 (padding slots included)
    const size_t slots = ned / ED_ROWS_PER_VALIDATOR;
    std::vector<ge_acc_packed> acc(ED_ROWS_PER_VALIDATOR);
    Sha512Hist h5[2];
    for (size_t i = 0; i < slots; i++) {
        EdSlotInfo e;
        memset(&e, 0, sizeof e);
        ge_cached51 tab[4];
        ge51 R = ge_identity51();
        if (i < a.n_max) {
            EdTriple t;
            effective_triple(blob_validators(blob) + i, &t);
            uint8_t digest[64];
            sha512_validator_prepare(t, h5, digest);
            ed_slot_prepare(t, digest, &e, tab, &R);
        } else {
            sha512_padding_prepare(&h5[0]);
            sha512_padding_prepare(&h5[1]);
            ed_padding_table(tab, &e.tab);
            e.ok = 1;
        }
        const ge_acc51 q = ed_straus_ladder(e.s, e.h, tab, acc.data());
        if (i < a.n_max) aux[AUX_SIG_OK + i] = (e.ok && ed_result_equals(q, R)) ? 1 : 0;
        for (int r = 0; r < ED_ROWS_PER_VALIDATOR; r++) {
            const size_t row = i * ED_ROWS_PER_VALIDATOR + r;
            if (row < n512) sha512_row_cells(a.t512, a.n512, row, r % S512_ROWS_PER_CHUNK, &h5[r / S512_ROWS_PER_CHUNK]);
            ed_row_cells(a.ted, a.ned, row, r, e.s, e.h, acc[r], e.tab);
        }
    }
    return 0;
}

// ---- the kernels' per-thread bodies (stark_rows.cuh) in a loop ----
static std::vector<gl> periodic_lde(const uint64_t* periodic, int n_per, size_t P, size_t n) {
    std::vector<gl> lde((size_t)n_per * 2 * P);
    const gl sh = gl_pow(GL_GEN, n / P);
    for (int pc = 0; pc < n_per; pc++) {
        std::vector<gl> coef(periodic + (size_t)pc * P, periodic + (size_t)(pc + 1) * P);
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
    return lde;
}

template <template <int> class Fn, class... A>
static void dispatch(int table, A&&... a) {
    switch (table) {
        case AIR_SHA256: Fn<AIR_SHA256>::run(a...); break;
        case AIR_SHA512: Fn<AIR_SHA512>::run(a...); break;
        case AIR_ED25519: Fn<AIR_ED25519>::run(a...); break;
        case AIR_LOGIC: Fn<AIR_LOGIC>::run(a...); break;
        default: Fn<AIR_RANGE>::run(a...); break;
    }
}
template <int T> struct GenFn { static void run(const BusPassArgs& a, size_t r) { bus_gen_row<T>(a, r); } };
template <int T> struct CountFn { static void run(const BusPassArgs& a, size_t r) { bus_count_row<T>(a, r); } };
template <int T> struct QuotFn { static void run(const QuotientArgs& a, size_t p) { quotient_point<T>(a, p); } };

static BusPassArgs pass_args(uint32_t kind, uint32_t n_max, const uint64_t* trace, const uint64_t* consts, const uint64_t* periodic,
                             size_t n, size_t P) {
    BusPassArgs a;
    memset(&a, 0, sizeof a);
    a.trace = trace; a.kconst = consts; a.per = periodic; a.n = n; a.P = (int)P;
    a.shape = air_shape(kind, n_max, g_chain, strlen(g_chain));
    return a;
}

// bus_count_kernel: hist [BUS_HIST_SIZE + 1] (last entry: out-of-range flag)
extern "C" void hostsim_bus_count(uint32_t kind, uint32_t n_max, int table, const uint64_t* trace, const uint64_t* consts,
                                  const uint64_t* periodic, size_t n, size_t P, uint32_t* hist) {
    BusPassArgs a = pass_args(kind, n_max, trace, consts, periodic, n, P);
    a.hist = hist;
    for (size_t r = 0; r < n; r++) dispatch<CountFn>(table, a, r);
}

// bus_gen_kernel + bus_scan_kernel: aux [2 (H + 1)][n], total[2]
extern "C" void hostsim_bus_aux(uint32_t kind, uint32_t n_max, int table, int n_helpers, const uint64_t* trace, const uint64_t* consts,
                                const uint64_t* periodic, size_t n, size_t P, const uint64_t beta[2], const uint64_t gamma[2],
                                uint64_t* aux, uint64_t total[2]) {
    BusPassArgs a = pass_args(kind, n_max, trace, consts, periodic, n, P);
    a.beta = gl2_make(beta[0], beta[1]);
    a.gamma = gl2_make(gamma[0], gamma[1]);
    a.aux = aux;
    std::vector<gl2> rowsum(n);
    a.rowsum = rowsum.data();
    for (size_t r = 0; r < n; r++) dispatch<GenFn>(table, a, r);
    gl2 tot = gl2_from(0);
    for (size_t r = 0; r < n; r++) tot = gl2_add(tot, rowsum[r]);
    const gl2 step = gl2_scale(tot, gl_inv((gl)n));
    gl2 z = gl2_from(0);
    for (size_t r = 0; r < n; r++) {
        aux[(size_t)(2 * n_helpers) * n + r] = z.a0;
        aux[(size_t)(2 * n_helpers + 1) * n + r] = z.a1;
        z = gl2_sub(gl2_add(z, rowsum[r]), step);
    }
    total[0] = tot.a0;
    total[1] = tot.a1;
}

// quotient_kernel: LDEs [cols][m] bit-reversed -> out [2][m] natural
extern "C" void hostsim_quotient(uint32_t kind, uint32_t n_max, int table, const uint64_t* lde_m, const uint64_t* lde_k,
                                 const uint64_t* lde_a, const uint64_t* periodic, int n_per, size_t P, size_t n, const uint64_t total[2],
                                 const uint64_t beta[2], const uint64_t gamma[2], const uint64_t alpha[2], uint64_t* out) {
    const unsigned log_n = log2u((uint32_t)n);
    const std::vector<gl> tab = periodic_lde(periodic, n_per, P, n);
    QuotientArgs qa;
    memset(&qa, 0, sizeof qa);
    qa.lde_m = lde_m; qa.lde_k = lde_k; qa.lde_a = lde_a;
    qa.m = n << 1; qa.log_m = log_n + 1; qa.rate_bits = 1;
    qa.pertab = tab.data(); qa.P = (int)P;
    qa.shape = air_shape(kind, n_max, g_chain, strlen(g_chain));
    qa.alpha[0] = alpha[0]; qa.alpha[1] = alpha[1];
    qa.beta = gl2_make(beta[0], beta[1]); qa.gamma = gl2_make(gamma[0], gamma[1]);
    qa.s_over_n = gl2_scale(gl2_make(total[0], total[1]), gl_inv((gl)n));
    const gl gn = gl_pow(GL_GEN, n);
    qa.zh_inv[0] = gl_inv(gl_sub(gn, 1));
    qa.zh_inv[1] = gl_inv(gl_sub(gl_neg(gn), 1));
    qa.out = out;
    for (size_t p = 0; p < qa.m; p++) dispatch<QuotFn>(table, qa, p);
}
