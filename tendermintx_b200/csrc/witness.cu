// K6 (SHA-256 table), K7 (SHA-512 table) and K8 (Ed25519 table): witness columns filled on the GPU.
//
// Replaces the witness generation of plonky2x's `curta_sha256_variable`, `get_root_from_merkle_proof*`,
// `get_root_from_hashed_leaves` and `curta_eddsa_verify_sigs_conditional`
// [REF circuits/builder/verify.rs:147,165,202,205,248-259,285,376; validator.rs:228,248; shared.rs:194,197].
//
// Mapping to the machine:
//   * SHA-256: one CTA per message (64 threads per 64-byte chunk).  Thread 0 runs the sequential compression into
//     a 832-byte shared-memory history; then every thread owns one round = one trace row and writes its 370
//     cells, so each column receives 64 (or 128) consecutive rows = 512-byte coalesced runs.  The validator-set
//     tree is one launch per level (the only true dependency), header proofs one CTA per proof.
//   * SHA-512 + Ed25519: one validator per CTA, in two phases.  Phase 1 (64 threads, small footprint so it can share
//     the SMs with another stream's kernels): h = SHA-512(R || A || M) is reduced mod l on the SM and feeds the [h]A
//     ladder; the two 256-step ladders ([s]B and [h]A) run on two warps with 5 x 51-bit-limb field arithmetic in
//     registers and park the canonical (res, temp) of every step in a 128 KB-per-validator scratch.  Phase 2 (512
//     threads): each thread expands one ladder row into its 945 cells (17 multiplication gadgets: product limbs,
//     quotient, carries) and the 2 x 128 SHA-512 rows are written.
//   Integer / bit work, HBM-write bound at best: no tensor cores.
#include "ctx.cuh"
#include "witness_jobs.cuh"

namespace tmx {

__global__ void __launch_bounds__(64) sha256_leaves_kernel(WitnessArgs a) {
    __shared__ Sha256Hist hs;
    __shared__ int active;
    const uint32_t s = blockIdx.x / a.np, i = blockIdx.x % a.np;
    if (threadIdx.x == 0) active = sha256_leaf_prepare(a, s, i, &hs) ? 1 : 0;
    __syncthreads();
    if (active) sha256_row_cells(a.t256, a.n256, sha256_leaf_row0(a, s, i) + threadIdx.x, threadIdx.x, &hs);
}

__global__ void __launch_bounds__(128) sha256_inner_kernel(WitnessArgs a, uint32_t level) {
    __shared__ Sha256Hist hs[2];
    const uint32_t per_set = a.np >> level;
    const uint32_t s = blockIdx.x / per_set, i = blockIdx.x % per_set;
    if (threadIdx.x == 0) sha256_inner_prepare(a, s, level, i, hs);
    __syncthreads();
    sha256_row_cells(a.t256, a.n256, sha256_inner_row0(a, s, level, i) + threadIdx.x, threadIdx.x & 63, &hs[threadIdx.x >> 6]);
}

__global__ void __launch_bounds__(128) sha256_header_kernel(WitnessArgs a) {
    __shared__ Sha256Hist hs[2];
    __shared__ HeaderProofDesc d;
    __shared__ uint8_t cur[32];
    __shared__ int nb;
    if (threadIdx.x == 0) header_proof_desc(a, blockIdx.x, &d);
    __syncthreads();
    size_t chunk = d.chunk0;
    for (int j = 0; j < 5; j++) {
        if (threadIdx.x == 0) nb = header_proof_prepare(d, j, cur, hs);
        __syncthreads();
        if ((int)threadIdx.x < nb * 64)
            sha256_row_cells(a.t256, a.n256, chunk * 64 + threadIdx.x, threadIdx.x & 63, &hs[threadIdx.x >> 6]);
        chunk += nb;
        __syncthreads();
    }
    if (threadIdx.x < 32) a.aux[AUX_PROOF_ROOT + 32 * blockIdx.x + threadIdx.x] = cur[threadIdx.x];
}

__global__ void __launch_bounds__(64) sha256_padding_kernel(WitnessArgs a, size_t first_chunk) {
    __shared__ Sha256Hist hs;
    if (threadIdx.x == 0) sha256_padding_prepare(&hs);
    __syncthreads();
    sha256_row_cells(a.t256, a.n256, (first_chunk + blockIdx.x) * 64 + threadIdx.x, threadIdx.x, &hs);
}

__global__ void __launch_bounds__(128) sha512_padding_kernel(WitnessArgs a, size_t first_row) {
    __shared__ Sha512Hist hs;
    if (threadIdx.x == 0) sha512_padding_prepare(&hs);
    __syncthreads();
    const size_t row = first_row + (size_t)blockIdx.x * S512_ROWS_PER_CHUNK + threadIdx.x;
    if (row < a.n512) sha512_row_cells(a.t512, a.n512, row, threadIdx.x, &hs);
}

struct LadderShared {
    Sha512Hist h5[2];
    EdTriple triple;
    EdSlot slot;
    uint8_t digest[64];
    ge51 Ps, Ph;
    bool ok_r;
};

// Phase 1 (latency-bound, small footprint: 64 threads, so other kernels share the SMs while it runs): per validator
// SHA-512(R || A || M) -> h mod l, decompress A and R on two warps, run the [s]B and [h]A ladders on two warps and park
// the canonical (res, temp) of every step in global scratch (128 B per point); verdict of the signature equation.
__global__ void __launch_bounds__(64) ed25519_ladder_kernel(WitnessArgs a, ge_packed* __restrict__ points) {
    __shared__ LadderShared sh;
    const uint32_t i = blockIdx.x, tid = threadIdx.x;
    ge_packed* res = points + (size_t)i * 1024;
    ge_packed* tmp = res + 512;
    if (tid == 0) {
        effective_triple(blob_validators(a.blob) + i, &sh.triple);
        sha512_validator_prepare(sh.triple, sh.h5, sh.digest);
        fe256 sb = fe256_from_bytes(sh.triple.sig + 32);
        for (int k = 0; k < 4; k++) sh.slot.s[k] = sb.w[k];
        sc_reduce512(sh.digest, sh.slot.h);
        sh.slot.ok = sc_lt_l(sh.slot.s);
    }
    __syncthreads();
    if (tid == 0 && !ge_decompress51(sh.triple.pk, &sh.slot.A)) {
        sh.slot.ok = false;
        sh.slot.A = ge_identity51();
    }
    if (tid == 32) {
        sh.ok_r = ge_decompress51(sh.triple.sig, &sh.slot.R);
        if (!sh.ok_r) sh.slot.R = ge_identity51();
    }
    __syncthreads();
    if (tid == 0) sh.Ps = ed_ladder(sh.slot.s, ge_base51(), res, tmp);
    if (tid == 32) sh.Ph = ed_ladder(sh.slot.h, sh.slot.A, res + 256, tmp + 256);
    __syncthreads();
    if (tid == 0) a.aux[AUX_SIG_OK + i] = (sh.ok_r && ed_slot_verdict(sh.slot, sh.Ps, sh.Ph)) ? 1 : 0;
}

// Phase 2 (throughput): one thread per trace row -- 2 x 128 SHA-512 rows and 512 ladder steps per validator.
__global__ void __launch_bounds__(512) ed25519_expand_kernel(WitnessArgs a, const ge_packed* __restrict__ points) {
    __shared__ Sha512Hist h5[2];
    __shared__ uint64_t sc[2][4];
    const uint32_t i = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        EdTriple t;
        uint8_t digest[64];
        effective_triple(blob_validators(a.blob) + i, &t);
        sha512_validator_prepare(t, h5, digest);
        fe256 sb = fe256_from_bytes(t.sig + 32);
        for (int k = 0; k < 4; k++) sc[0][k] = sb.w[k];
        sc_reduce512(digest, sc[1]);
    }
    __syncthreads();
    if (tid < S512_ROWS_PER_VALIDATOR)
        sha512_row_cells(a.t512, a.n512, (size_t)i * S512_ROWS_PER_VALIDATOR + tid, tid % S512_ROWS_PER_CHUNK, &h5[tid / S512_ROWS_PER_CHUNK]);
    const uint64_t* s = sc[tid >> 8];
    const int bit = (int)((s[(tid & 255) >> 6] >> (tid & 63)) & 1);
    const ge_packed* res = points + (size_t)i * 1024;
    ed_row_cells(a.ted, a.ned, (size_t)i * ED_ROWS_PER_VALIDATOR + tid, bit, res[tid], res[512 + tid]);
}

// padding blocks of the Ed25519 table: [0]B ladders (valid rows, bit = 0)
__global__ void __launch_bounds__(256) ed25519_padding_kernel(WitnessArgs a, size_t first_row) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ge_packed* res = reinterpret_cast<ge_packed*>(smem_raw);
    ge_packed* tmp = res + 256;
    if (threadIdx.x == 0) {
        const uint64_t zero[4] = {0, 0, 0, 0};
        ed_ladder(zero, ge_base51(), res, tmp);
    }
    __syncthreads();
    const size_t row = first_row + (size_t)blockIdx.x * 256 + threadIdx.x;
    if (row < a.ned) ed_row_cells(a.ted, a.ned, row, 0, res[threadIdx.x], tmp[threadIdx.x]);
}

int witness_tu_init() {
    TMX_CUDA(cudaMemcpyToSymbol(d_DUMMY_SIGNATURE, DUMMY_SIGNATURE, 64));
    TMX_CUDA(cudaStreamSynchronize(cudaStreamLegacy));  // staged copy, see context.cu upload()
    return TMX_OK;
}

static size_t pow2_at_least(size_t x) {
    size_t p = 1;
    while (p < x) p *= 2;
    return p;
}

static int make_args(tmx_ctx* ctx, const uint8_t* d_blob, uint32_t kind, uint32_t n_max, uint64_t* t256, uint64_t* t512,
                     uint64_t* ted, uint8_t* d_aux, WitnessArgs* a) {
    if (kind > 1 || n_max == 0 || n_max > 4096) return fail(TMX_E_INPUT, "witness: bad kind / n_max");
    a->blob = d_blob;
    a->kind = kind;
    a->n_max = n_max;
    a->np = (uint32_t)pow2_at_least(n_max);
    a->log_np = ilog2(a->np);
    size_t dims[6];
    tmx_trace_dims(kind, n_max, dims);
    a->t256 = t256; a->n256 = dims[0];
    a->t512 = t512; a->n512 = dims[2];
    a->ted = ted; a->ned = dims[4];
    void* p = nullptr;
    int rc = ctx_scratch(ctx, 2, (size_t)2 * 2 * a->np * 33, &p);
    if (rc) return rc;
    a->nodes = (uint8_t*)p;
    a->node_en = a->nodes + (size_t)2 * 2 * a->np * 32;
    a->aux = d_aux;
    return TMX_OK;
}

static int run_sha256(tmx_ctx* ctx, const WitnessArgs& a, cudaStream_t st) {
    const uint32_t ns = n_sets(a.kind);
    sha256_leaves_kernel<<<ns * a.np, 64, 0, st>>>(a);
    ctx->launches++;
    for (uint32_t l = 1; l <= a.log_np; l++) {
        sha256_inner_kernel<<<ns * (a.np >> l), 128, 0, st>>>(a, l);
        ctx->launches++;
    }
    sha256_header_kernel<<<n_header_proofs(a.kind), 128, 0, st>>>(a);
    ctx->launches++;
    const size_t used = sha256_used_chunks(a.kind, a.n_max, a.np), total = a.n256 / 64;
    if (total > used) {
        sha256_padding_kernel<<<(unsigned)(total - used), 64, 0, st>>>(a, used);
        ctx->launches++;
    }
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

size_t witness_points_bytes(uint32_t n_max) { return (size_t)n_max * 1024 * sizeof(ge_packed); }

// phase 1 only (stream-ordered); points must hold witness_points_bytes(n_max)
int run_ed25519_ladder(tmx_ctx* ctx, const WitnessArgs& a, void* points, cudaStream_t st) {
    // Measured: while the ladders run (one CTA on 128 of the 148 SMs, ~6 ms) the NTT passes of the SHA-256 table are kept
    // off those SMs (different L1 / shared-memory split), so that LDE takes 6.8 ms instead of 1.6 ms.  Forcing the
    // largest shared-memory split on the ladder kernel fixes the LDE (3.1 ms) but then the ladder's two warps share
    // issue slots with the leaf hashing and the proof gets 2.4 ms slower overall; left as is.
    ed25519_ladder_kernel<<<a.n_max, 64, 0, st>>>(a, (ge_packed*)points);
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

// phase 2 + padding rows of the SHA-512 and Ed25519 tables
int run_ed25519_expand(tmx_ctx* ctx, const WitnessArgs& a, const void* points, cudaStream_t st) {
    ed25519_expand_kernel<<<a.n_max, 512, 0, st>>>(a, (const ge_packed*)points);
    ctx->launches++;
    const size_t used512 = (size_t)a.n_max * S512_ROWS_PER_VALIDATOR;
    if (a.n512 > used512) {
        sha512_padding_kernel<<<(unsigned)((a.n512 - used512 + S512_ROWS_PER_CHUNK - 1) / S512_ROWS_PER_CHUNK), S512_ROWS_PER_CHUNK, 0, st>>>(a, used512);
        ctx->launches++;
    }
    const size_t used_ed = (size_t)a.n_max * ED_ROWS_PER_VALIDATOR;
    if (a.ned > used_ed) {
        const size_t psmem = 512 * sizeof(ge_packed);
        TMX_CUDA(cudaFuncSetAttribute(ed25519_padding_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        ed25519_padding_kernel<<<(unsigned)((a.ned - used_ed + 255) / 256), 256, psmem, st>>>(a, used_ed);
        ctx->launches++;
    }
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

static int run_ed25519(tmx_ctx* ctx, const WitnessArgs& a, cudaStream_t st) {
    void* points = nullptr;
    int rc = ctx_scratch(ctx, 1, witness_points_bytes(a.n_max), &points);
    if (rc) return rc;
    rc = run_ed25519_ladder(ctx, a, points, st);
    if (rc) return rc;
    return run_ed25519_expand(ctx, a, points, st);
}

int witness_make_args(tmx_ctx* ctx, const uint8_t* d_blob, uint32_t kind, uint32_t n_max, uint64_t* t256, uint64_t* t512,
                      uint64_t* ted, uint8_t* d_aux, WitnessArgs* a) {
    return make_args(ctx, d_blob, kind, n_max, t256, t512, ted, d_aux, a);
}
int witness_run_sha256(tmx_ctx* ctx, const WitnessArgs& a, cudaStream_t st) { return run_sha256(ctx, a, st); }

}  // namespace tmx

using namespace tmx;

extern "C" int tmx_trace_dims(uint32_t kind, uint32_t n_max, size_t dims[6]) {
    if (kind > 1 || n_max == 0 || !dims) return fail(TMX_E_INPUT, "tmx_trace_dims: bad arguments");
    const size_t np = pow2_at_least(n_max);
    dims[0] = pow2_at_least(sha256_used_chunks(kind, n_max, (uint32_t)np) * S256_ROUNDS);
    dims[1] = S256_COLS;
    dims[2] = pow2_at_least((size_t)n_max * S512_ROWS_PER_VALIDATOR);
    dims[3] = S512_COLS;
    dims[4] = pow2_at_least((size_t)n_max * ED_ROWS_PER_VALIDATOR);
    dims[5] = ED_COLS;
    return TMX_OK;
}

extern "C" size_t tmx_witness_aux_bytes(uint32_t n_max) { return aux_bytes(n_max); }

extern "C" int tmx_sha256_trace(tmx_ctx* ctx, const uint8_t* d_blob, uint32_t kind, uint32_t n_max, uint64_t* d_t256,
                                uint8_t* d_aux, void* stream) {
    if (!ctx || !d_blob || !d_t256 || !d_aux) return fail(TMX_E_INPUT, "tmx_sha256_trace: NULL argument");
    WitnessArgs a;
    int rc = make_args(ctx, d_blob, kind, n_max, d_t256, nullptr, nullptr, d_aux, &a);
    if (rc) return rc;
    return run_sha256(ctx, a, pick_stream(ctx, stream));
}

extern "C" int tmx_ed25519_trace(tmx_ctx* ctx, const uint8_t* d_blob, uint32_t kind, uint32_t n_max, uint64_t* d_t512,
                                 uint64_t* d_ted, uint8_t* d_aux, void* stream) {
    if (!ctx || !d_blob || !d_t512 || !d_ted || !d_aux) return fail(TMX_E_INPUT, "tmx_ed25519_trace: NULL argument");
    WitnessArgs a;
    int rc = make_args(ctx, d_blob, kind, n_max, nullptr, d_t512, d_ted, d_aux, &a);
    if (rc) return rc;
    return run_ed25519(ctx, a, pick_stream(ctx, stream));
}

extern "C" int tmx_witness_generate(tmx_ctx* ctx, const uint8_t* d_blob, uint32_t kind, uint32_t n_max, uint64_t* d_t256,
                                    uint64_t* d_t512, uint64_t* d_ted, uint8_t* d_aux, void* stream) {
    if (!ctx || !d_blob || !d_t256 || !d_t512 || !d_ted || !d_aux)
        return fail(TMX_E_INPUT, "tmx_witness_generate: NULL argument");
    WitnessArgs a;
    int rc = make_args(ctx, d_blob, kind, n_max, d_t256, d_t512, d_ted, d_aux, &a);
    if (rc) return rc;
    cudaStream_t st = pick_stream(ctx, stream);
    rc = run_ed25519(ctx, a, st);
    if (rc) return rc;
    return run_sha256(ctx, a, st);
}
