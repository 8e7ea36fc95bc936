// CPU simulation of the witness kernels' per-thread logic (witness_jobs.cuh / witness.cuh compiled for the
// host).  TEST TOOL ONLY: lets tests/test_hostsim.py compare the kernel logic with the CPU oracle in the
// authoring container, which has no GPU.  It is not part of libtmx.so and nothing in the product calls it.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../tendermintx_b200/csrc/witness_jobs.cuh"
#include "../tendermintx_b200/csrc/air.cuh"

using namespace tmx;

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
    // validators: SHA-512 + Ed25519
    std::vector<ge_packed> res(512), tmp(512);
    Sha512Hist h5[2];
    for (uint32_t i = 0; i < a.n_max; i++) {
        EdTriple t;
        effective_triple(blob_validators(blob) + i, &t);
        uint8_t digest[64];
        sha512_validator_prepare(t, h5, digest);
        for (int r = 0; r < S512_ROWS_PER_VALIDATOR; r++)
            sha512_row_cells(a.t512, a.n512, (size_t)i * S512_ROWS_PER_VALIDATOR + r, r % S512_ROWS_PER_CHUNK, &h5[r / S512_ROWS_PER_CHUNK]);
        EdSlot e;
        ed_slot_prepare(t, digest, &e);
        ge51 Ps = ed_ladder(e.s, ge_base51(), res.data(), tmp.data());
        ge51 Ph = ed_ladder(e.h, e.A, res.data() + 256, tmp.data() + 256);
        aux[AUX_SIG_OK + i] = ed_slot_verdict(e, Ps, Ph) ? 1 : 0;
        for (int r = 0; r < 512; r++) {
            const uint64_t* sc = r < 256 ? e.s : e.h;
            int bit = (sc[(r & 255) >> 6] >> (r & 63)) & 1;
            ed_row_cells(a.ted, a.ned, (size_t)i * 512 + r, bit, res[r], tmp[r]);
        }
    }
    sha512_padding_prepare(&h5[0]);
    for (size_t row = (size_t)a.n_max * S512_ROWS_PER_VALIDATOR; row < n512; row++)
        sha512_row_cells(a.t512, a.n512, row, (int)((row - (size_t)a.n_max * S512_ROWS_PER_VALIDATOR) % S512_ROWS_PER_CHUNK), &h5[0]);
    if ((size_t)a.n_max * 512 < ned) {
        uint64_t zero[4] = {0, 0, 0, 0};
        ed_ladder(zero, ge_base51(), res.data(), tmp.data());
        for (size_t row = (size_t)a.n_max * 512; row < ned; row++) ed_row_cells(a.ted, a.ned, row, 0, res[row & 255], tmp[row & 255]);
    }
    return 0;
}

// quotient kernel logic (stark.cu quotient_kernel) on the host: lde [C][m] bit-reversed -> out [2][m] natural
struct HostRow {
    const gl* base;
    size_t stride;
    FB operator[](int c) const { return FB::mk(base[(size_t)c * stride]); }
};
extern "C" void hostsim_quotient(uint32_t kind, uint32_t n_max, int table, const uint64_t* lde, size_t n, const uint64_t alpha[2],
                                 uint64_t* out) {
    const unsigned log_n = log2u((uint32_t)n), log_m = log_n + 1;
    const size_t m = n << 1;
    std::vector<gl> tab = air_periodic_lde_table(table, log_n, h_K256, h_K512, AirShape{kind, n_max});
    const size_t P = air_period(table, n);
    const gl gn = gl_pow(GL_GEN, n);
    const gl zh_inv[2] = {gl_inv(gl_sub(gn, 1)), gl_inv(gl_sub(gl_neg(gn), 1))};
    for (size_t p = 0; p < m; p++) {
        const uint32_t j = bitrev32((uint32_t)p, log_m);
        const uint32_t jn = (j + 2) & (uint32_t)(m - 1);
        const size_t pn = bitrev32(jn, log_m);
        HostRow l{lde + p, m}, nx{lde + pn, m}, per{tab.data() + (j & (2 * P - 1)), (size_t)2 * P};
        ConstraintAcc<FB> acc;
        acc.acc0 = FB::c(0); acc.acc1 = FB::c(0);
        acc.alpha0 = FB::mk(alpha[0]); acc.alpha1 = FB::mk(alpha[1]);
        air_eval<FB>(table, l, nx, per, acc);
        out[j] = gl_mul(acc.acc0.v, zh_inv[j & 1]);
        out[m + j] = gl_mul(acc.acc1.v, zh_inv[j & 1]);
    }
}
