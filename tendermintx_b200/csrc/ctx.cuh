// Internal definition of tmx_ctx (one per GPU) and launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <map>
#include <atomic>
#include <mutex>
#include "gl.cuh"
#include "../../include/tmx.h"

namespace tmx {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define TMX_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::tmx::fail(TMX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// Twiddle tables are flat (at most 2^18 entries of 8 bytes for the tables proved here: L2-resident).  A two-level
// table saves memory but costs a second load and an extra field multiplication per use, and the transforms are
// bound by integer issue, not by memory.
struct NttTables {
    gl* small = nullptr;  // w_L^e, e < L/2 (L = 2^k, k <= 10)
    gl* full = nullptr;   // w_{2^k}^e, e < 2^k
};

}  // namespace tmx

struct tmx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::atomic<uint64_t> launches{0};  // the tails of one proof launch from several host threads
    std::mutex tables_mu;               // guards the lazily built twiddle / coset tables below
    int sm_count = 148;
    std::map<unsigned, tmx::NttTables> fwd, inv;      // keyed by log size
    std::map<unsigned, tmx::gl*> coset_scale;         // keyed by log n: 7^i / n, i < n
    // grow-only scratch buffers
    void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    std::vector<void*> owned;  // device allocations freed at destroy
};

namespace tmx {

int ctx_scratch(tmx_ctx* ctx, int slot, size_t bytes, void** out);
int ctx_ntt_tables(tmx_ctx* ctx, unsigned log_n, bool inverse, const NttTables** out);
int ctx_coset_scale(tmx_ctx* ctx, unsigned log_n, const gl** out);
inline cudaStream_t pick_stream(tmx_ctx* ctx, void* stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

// internal cross-file entry points
int lde_forward_cosets(tmx_ctx* ctx, const gl* coeffs, gl* d_out, size_t n_cols, unsigned log_n, unsigned rate_bits,
                       cudaStream_t st);
// Merkle tree over n_rows = 2^log_rows leaves; leaf j = hash_or_noop of leaf_len elements at
// base[j * row_stride + i * elem_stride]
int merkle_generic(tmx_ctx* ctx, const gl* base, size_t leaf_len, size_t row_stride, size_t elem_stride, unsigned log_rows,
                   unsigned cap_height, gl* d_digests, cudaStream_t st);
// smallest w < 2^40 such that Poseidon(state with state[pos] = w)[7] has `bits` leading zero bits
int pow_grind(tmx_ctx* ctx, const gl state[12], int pos, unsigned bits, uint64_t* witness, gl* d_scratch, cudaStream_t st);
// in-place transform of n_cols columns with a caller-owned temporary of the same size (tmx_ntt uses the context's scratch)
int ntt_with_scratch(tmx_ctx* ctx, gl* d_data, size_t n_cols, unsigned log_n, bool inverse, gl* d_tmp, cudaStream_t st);

// per translation unit Poseidon constant upload hooks
int merkle_tu_init();
int witness_tu_init();

}  // namespace tmx
