// Shared declarations of the STARK layer: protocol parameters (Curta's STARK configuration as recalled in
// rate_bits 1, cap height 4, 16-bit PoW, 84 queries, arity-16 FRI, final polynomial <= 2^5)
// and the per-table GPU prover state.
#pragma once
#include "ctx.cuh"
#include "poseidon.cuh"
#include "air.cuh"
#include "witness.cuh"
#include <vector>
#include <map>
#include <functional>

namespace tmx {

constexpr unsigned STARK_RATE_BITS = 1;
constexpr unsigned STARK_CAP_HEIGHT = 4;
constexpr unsigned STARK_POW_BITS = 16;
constexpr int STARK_NUM_QUERIES = 84;
constexpr unsigned STARK_ARITY_BITS = 4;
constexpr unsigned STARK_FINAL_POLY_BITS = 5;
constexpr uint64_t STARK_PROOF_MAGIC = 0x50584D54ULL;  // "TMXP"
constexpr int STARK_N_TABLES = 3;

unsigned fri_num_layers(unsigned degree_bits);

// Device buffers reused from proof to proof (grow-only) and the cached periodic-column tables.
struct TableProver {
    gl* d_lde = nullptr; size_t sz_lde = 0;
    gl* d_coeffs = nullptr; size_t sz_coeffs = 0;
    gl* d_dig_t = nullptr; size_t sz_dig_t = 0;
    gl* d_dig_q = nullptr; size_t sz_dig_q = 0;
    gl* d_dig_fri = nullptr; size_t sz_dig_fri = 0;
    gl* d_qv = nullptr; size_t sz_qv = 0;
    gl* d_qcoef = nullptr; size_t sz_qcoef = 0;
    gl* d_qlde = nullptr; size_t sz_qlde = 0;
    gl2* d_ypa = nullptr; gl2* d_ypb = nullptr; size_t sz_yp = 0;
    gl2* d_open = nullptr; size_t sz_open = 0;
    gl2* d_apow = nullptr; size_t sz_apow = 0;
    uint32_t* d_idx = nullptr; size_t sz_idx = 0;
    gl2* d_fri_base = nullptr; size_t sz_fri = 0;
    gl2* d_fri[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    gl* d_query = nullptr; size_t sz_query = 0;
    std::map<uint64_t, gl*> pertabs;
    AirShape shape = {0, 0};  // circuit shape (kind, n_max): the SHA-256 table's public columns depend on it
    // device-time stamps of the trace commitment of the last prove(): LDE start, LDE end = Merkle start, Merkle end
    cudaEvent_t ev_phase[3] = {nullptr, nullptr, nullptr};
    float last_lde_ms = 0.f, last_merkle_ms = 0.f;
    gl* h_pinned = nullptr; size_t sz_pinned = 0;  // pinned staging buffer for the transcript's device -> host copies

    // Appends this table's proof to `proof` and advances the transcript.  `on_trace_committed` (optional) is called once
    // the trace commitment has been enqueued and its completion event recorded: the caller uses it to start
    // work on another stream that should overlap with the latency-bound rest of this table's proof.
    int prove(tmx_ctx* ctx, int table, const gl* d_trace, unsigned log_n, Challenger& ch, std::vector<gl>& proof, cudaStream_t st,
              const std::function<int(cudaEvent_t)>& on_trace_committed = nullptr);
    int d2h(std::vector<gl>& dst, const gl* src, size_t n, cudaStream_t st);
    int reserve(tmx_ctx* ctx, size_t C, size_t n, size_t m, size_t dig_t);
    int reserve_queries(tmx_ctx* ctx, size_t n);
    int periodic_tables(tmx_ctx* ctx, int table, unsigned log_n, const gl** out);
    void release();
};

// host verifier of one table proof; returns 0 or a diagnostic code
int verify_table(int table, size_t n, AirShape shape, const gl* proof, size_t proof_len, size_t* pos, Challenger& ch);

}  // namespace tmx
