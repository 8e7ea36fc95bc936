// K2: Poseidon-12 leaf hashing and Merkle tree (with cap) over the rows of a column-major matrix.
//
// Replaces plonky2 0.2.0 hash/merkle_tree.rs MerkleTree::new + hash/poseidon.rs as used by
// PolynomialBatch::from_values, reached from the reference through `circuit.prove()`
// [REF circuits/skip.rs:214, circuits/step.rs:196].
//
// One thread per leaf: the LDE matrix is column-major, so the 32 threads of a warp read 32 consecutive rows of
// one column per load (256 coalesced bytes) and no transpose pass is ever materialised (plonky2 transposes
// on the CPU).  The sponge state lives in registers; round constants sit in __constant__ memory (every
// thread of a warp reads the same constant in the same cycle).  This kernel is bound by 32-bit integer
// multiply issue (about 1.1 k Goldilocks multiplications per permutation), not by HBM.
#include "ctx.cuh"
#include "poseidon.cuh"
#include <cstdlib>

namespace tmx {

#ifndef TMX_LEAF_T
#define TMX_LEAF_T 128
#endif
// Five CTAs per SM (92 registers): measured inside the pool bench against 9 / 8 / 7 / 6 / 4 / 3 CTAs per SM (56 ... 138 registers):
// 47.8 / 46.7 / 46.6 / 46.3 / 46.4 / 46.5 ms per proof, 46.0 ms here -- instruction-level parallelism inside a permutation pays more
// than extra warps.
__global__ void __launch_bounds__(TMX_LEAF_T, 640 / TMX_LEAF_T) leaf_hash_kernel(const gl* __restrict__ base, size_t leaf_len, size_t row_stride,
                                                         size_t elem_stride, size_t n_rows, gl* __restrict__ digests) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_rows) return;
    gl out[4];
    poseidon_hash_row(base + j * row_stride, elem_stride, leaf_len, out);
    reinterpret_cast<ulonglong2*>(digests + 4 * j)[0] = make_ulonglong2(out[0], out[1]);
    reinterpret_cast<ulonglong2*>(digests + 4 * j)[1] = make_ulonglong2(out[2], out[3]);
}

__global__ void __launch_bounds__(128) merkle_level_kernel(const gl* __restrict__ children, gl* __restrict__ parents,
                                                            size_t n_parents) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parents) return;
    const ulonglong2* c = reinterpret_cast<const ulonglong2*>(children + 8 * i);
    ulonglong2 a = c[0], b = c[1], d = c[2], e = c[3];
    gl l[4] = {a.x, a.y, b.x, b.y}, r[4] = {d.x, d.y, e.x, e.y}, out[4];
    poseidon_two_to_one(l, r, out);
    reinterpret_cast<ulonglong2*>(parents + 4 * i)[0] = make_ulonglong2(out[0], out[1]);
    reinterpret_cast<ulonglong2*>(parents + 4 * i)[1] = make_ulonglong2(out[2], out[3]);
}

__global__ void poseidon_permute_kernel(gl* states, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gl s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
    poseidon_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) states[12 * i + k] = s[k];
}

int merkle_tu_init() {
    TMX_CUDA(poseidon_upload_constants_tu());
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;

extern "C" size_t tmx_merkle_digest_count(unsigned log_rows, unsigned cap_height) {
    if (cap_height > log_rows) cap_height = log_rows;
    size_t total = 0;
    for (unsigned l = 0; l + cap_height <= log_rows; l++) total += (size_t)1 << (log_rows - l);
    return total;
}

namespace tmx {
int merkle_generic(tmx_ctx* ctx, const gl* base, size_t leaf_len, size_t row_stride, size_t elem_stride, unsigned log_rows,
                   unsigned cap_height, gl* d_digests, cudaStream_t st) {
    if (cap_height > log_rows) cap_height = log_rows;
    const size_t n = (size_t)1 << log_rows;
    leaf_hash_kernel<<<(unsigned)((n + TMX_LEAF_T - 1) / TMX_LEAF_T), TMX_LEAF_T, 0, st>>>(base, leaf_len, row_stride, elem_stride, n, d_digests);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    gl* lvl = d_digests;
    size_t m = n;
    // One launch per level.  Large levels are throughput work (128-thread CTAs); small ones are latency work: a
    // permutation is ~23 k instructions, so a warp that shares its SM sub-partition with another one takes twice as
    // long.  Below 16 Ki parents the level is spread one warp per CTA over as many SMs as possible (the first version
    // reduced the last nine levels inside one 256-thread CTA per cap subtree: 16 busy SMs, 0.3 - 0.45 ms per tree,
    // fifteen trees per proof).
    for (unsigned l = 0; l + cap_height < log_rows; l++) {
        gl* nxt = lvl + 4 * m;
        m >>= 1;
        const unsigned threads = m >= 16384 ? 128 : 32;
        merkle_level_kernel<<<(unsigned)((m + threads - 1) / threads), threads, 0, st>>>(lvl, nxt, m);
        ctx->launches++;
        TMX_CUDA(cudaGetLastError());
        lvl = nxt;
    }
    return TMX_OK;
}

// K9: proof-of-work grind.  Every thread tries one candidate; the smallest hit wins (deterministic, unlike
// plonky2's rayon find_any).
__global__ void __launch_bounds__(128) pow_grind_kernel(const gl* __restrict__ state, int pos, unsigned bits, uint64_t first,
                                                         unsigned long long* best) {
    const uint64_t cand = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    gl s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = state[i];
#pragma unroll
    for (int i = 0; i < 12; i++)
        if (i == pos) s[i] = cand;
    poseidon_permute(s);
    if ((s[7] >> (64 - bits)) == 0) atomicMin(best, (unsigned long long)cand);
}

int pow_grind(tmx_ctx* ctx, const gl state[12], int pos, unsigned bits, uint64_t* witness, gl* d_scratch, cudaStream_t st) {
    if (bits == 0) {
        *witness = 0;
        return TMX_OK;
    }
    if (!d_scratch) {  // 13 words; callers that run side by side bring their own
        void* p = nullptr;
        int rc = ctx_scratch(ctx, 3, 13 * sizeof(gl), &p);
        if (rc) return rc;
        d_scratch = (gl*)p;
    }
    gl* d_state = d_scratch;
    unsigned long long* d_best = (unsigned long long*)(d_state + 12);
    unsigned long long best = ~0ULL;
    TMX_CUDA(cudaMemcpyAsync(d_state, state, 12 * sizeof(gl), cudaMemcpyHostToDevice, st));
    TMX_CUDA(cudaMemcpyAsync(d_best, &best, sizeof best, cudaMemcpyHostToDevice, st));
    const uint64_t batch = (uint64_t)1 << (bits + 1 > 30 ? 30 : bits + 1);
    for (uint64_t first = 0; first < ((uint64_t)1 << 40); first += batch) {
        pow_grind_kernel<<<(unsigned)(batch / 128), 128, 0, st>>>(d_state, pos, bits, first, d_best);
        ctx->launches++;
        TMX_CUDA(cudaGetLastError());
        TMX_CUDA(cudaMemcpyAsync(&best, d_best, sizeof best, cudaMemcpyDeviceToHost, st));
        TMX_CUDA(cudaStreamSynchronize(st));
        if (best != ~0ULL) {
            *witness = best;
            return TMX_OK;
        }
    }
    return fail(TMX_E_CUDA, "pow_grind: no witness below 2^40");
}
}  // namespace tmx

extern "C" int tmx_poseidon_merkle(tmx_ctx* ctx, const uint64_t* d_cols, size_t n_cols, unsigned log_rows,
                                   unsigned cap_height, uint64_t* d_digests, void* stream) {
    if (!ctx || !d_cols || !d_digests || n_cols == 0 || log_rows > 30)
        return fail(TMX_E_INPUT, "tmx_poseidon_merkle: bad arguments");
    return merkle_generic(ctx, d_cols, n_cols, 1, (size_t)1 << log_rows, log_rows, cap_height, d_digests,
                          pick_stream(ctx, stream));
}

extern "C" int tmx_pow_grind(tmx_ctx* ctx, const uint64_t state[12], int pos, unsigned bits, uint64_t* witness, void* stream) {
    if (!ctx || !state || !witness || pos < 0 || pos >= 8 || bits > 40) return fail(TMX_E_INPUT, "tmx_pow_grind: bad arguments");
    return pow_grind(ctx, state, pos, bits, witness, nullptr, pick_stream(ctx, stream));
}

// The arithmetic of the device fast path (multiplier-free linear layer, unreduced lanes) compiled for the HOST, so
// the CPU test suite can pin it against the plain formulation without a GPU.  variant 0 = plain, 1 = device fast path, 2 = the host transcript's formulation.
extern "C" int tmx_host_poseidon_permute(uint64_t* states, size_t n, int variant) {
    if (!states || variant < 0 || variant > 2) return fail(TMX_E_INPUT, "tmx_host_poseidon_permute: bad arguments");
    poseidon_generate_constants();
    for (size_t i = 0; i < n; i++) {
        if (variant == 0) poseidon_permute_plain(states + 12 * i);
        else if (variant == 1) poseidon_permute_fast(states + 12 * i);
        else poseidon_permute_host(states + 12 * i);
    }
    return TMX_OK;
}

extern "C" int tmx_poseidon_permute(tmx_ctx* ctx, uint64_t* d_states, size_t n, void* stream) {
    if (!ctx || !d_states) return fail(TMX_E_INPUT, "tmx_poseidon_permute: bad arguments");
    if (n == 0) return TMX_OK;
    cudaStream_t st = pick_stream(ctx, stream);
    poseidon_permute_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_states, n);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}
