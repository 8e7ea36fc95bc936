// Shared declarations of the STARK layer: the per-circuit GPU prover state (device buffers of every table, the constant
// columns' commitment) and the phases of one proof.  Protocol parameters: params.cuh.
#pragma once
#include "ctx.cuh"
#include "params.cuh"
#include "poseidon.cuh"
#include "air.cuh"
#include "logic.cuh"
#include "circuit_def.cuh"
#include "witness.cuh"
#include <vector>
#include <map>
#include <functional>
#include <memory>

namespace tmx {

constexpr int STARK_N_TABLES = TMX_N_TABLES;

unsigned fri_num_layers(unsigned degree_bits);

// device buffers of one table (grow-only, reused from proof to proof)
struct TableDevice {
    // per circuit: constant columns (trace values, coset-scaled coefficients, LDE, Merkle digests), periodic columns
    gl* d_const = nullptr; gl* d_const_coef = nullptr; gl* d_const_lde = nullptr; gl* d_const_dig = nullptr;
    gl* d_per_trace = nullptr;  // [n_per][P] one period of every periodic column
    gl* d_per_lde = nullptr;    // [n_per][2P] values on the LDE coset
    // per proof
    gl* d_aux = nullptr;                      // second-round trace [A][n]
    gl* d_lde_m = nullptr; gl* d_coef_m = nullptr; gl* d_dig_m = nullptr;
    gl* d_lde_a = nullptr; gl* d_coef_a = nullptr; gl* d_dig_a = nullptr;
    gl2 total = {0, 0};
};

struct Prover {
    std::shared_ptr<const CircuitDef> def;
    AirShape shape = {0, 0};
    TableDevice tab[STARK_N_TABLES];
    // Scratch of one table's tail (quotient, openings, FRI, queries).  The tails run side by side, each on its own stream
    // and host thread, so every table owns a full set sized for itself, including a pinned staging buffer.
    struct Tail {
        gl* d_dig_q = nullptr; gl* d_dig_fri = nullptr; gl* d_qv = nullptr; gl* d_qcoef = nullptr; gl* d_qlde = nullptr;
        gl* d_ntt_tmp = nullptr; gl* d_pow = nullptr;
        gl2* d_ypa = nullptr; gl2* d_ypb = nullptr; gl2* d_open = nullptr; gl2* d_apow = nullptr;
        uint32_t* d_idx = nullptr;
        gl2* d_fri[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        gl* d_query = nullptr; size_t sz_query = 0;
        gl* h_pinned = nullptr; size_t sz_pinned = 0;
        cudaEvent_t ev_wait = nullptr;  // blocking-sync event: the host thread sleeps instead of spinning while it waits
    };
    Tail tail[STARK_N_TABLES + 1];  // the last one: staging buffer of the commitment rounds
    gl2* d_rowsum = nullptr;      // per-row bus sums, one region of rowsum_stride elements per table (tables run on their own streams)
    size_t rowsum_stride = 0;
    gl* d_small = nullptr;        // caps / totals staging: [STARK_N_TABLES][64 + 2]
    unsigned int* d_hist = nullptr;  // range-lookup histogram, then one int "bad" flag
    gl* d_range_trace = nullptr;  // first-round trace of the range table
    cudaEvent_t ev_phase[3 * STARK_N_TABLES] = {};  // per table: LDE start, LDE end = Merkle start, Merkle end
    float lde_ms[STARK_N_TABLES] = {}, merkle_ms[STARK_N_TABLES] = {};
    std::vector<void*> owned;

    int setup(tmx_ctx* ctx, std::shared_ptr<const CircuitDef> d);  // allocations + commitment of the constant columns
    void release();
    // round 1: histogram of table t's range lookups (t != range), range-table trace from the histogram, commitment
    int count_lookups(tmx_ctx* ctx, int t, const gl* d_trace, cudaStream_t st);
    int fill_range_trace(tmx_ctx* ctx, cudaStream_t st);
    int commit_main(tmx_ctx* ctx, int t, const gl* d_trace, cudaStream_t st);
    int finish_round1(tmx_ctx* ctx, Challenger& ch, std::vector<gl>& proof, cudaStream_t st, bool* range_ok);
    // round 2: helper columns + running sum of every table, commitment
    int commit_aux(tmx_ctx* ctx, int t, const gl* d_trace, gl2 beta, gl2 gamma, cudaStream_t st);
    int finish_round2(tmx_ctx* ctx, Challenger& ch, std::vector<gl>& proof, cudaStream_t st);
    // quotient, openings, FRI, queries of one table
    // (the table's transcript is a fork of the common one: the caller passes a copy that has absorbed the table index)
    int prove_tail(tmx_ctx* ctx, int t, gl2 beta, gl2 gamma, Challenger ch, std::vector<gl>& proof, cudaStream_t st);
    int d2h(std::vector<gl>& dst, const gl* src, size_t n, cudaStream_t st, int slot = STARK_N_TABLES);
    int alloc(void** p, size_t bytes);
};

int stark_quotient(tmx_ctx* ctx, Prover& pr, int table, const gl* d_lde_main, const gl* d_lde_aux, gl2 total, gl2 beta, gl2 gamma,
                   const gl alpha[2], gl* d_out, cudaStream_t st);

// host verifier; returns 0 or a diagnostic code (verify.cu)
int verify_proof(const CircuitDef& def, const gl* proof, size_t n_words, const uint8_t* input, size_t input_len, const uint8_t out32[32]);
void transcript_init(const gl digest[4], const uint8_t* input, size_t input_len, const uint8_t out32[32], Challenger& ch);

}  // namespace tmx
