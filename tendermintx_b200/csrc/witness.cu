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

// Phase 1 (latency-bound, small footprint so other kernels share the SMs while it runs): per validator slot
// SHA-512(R || A || M) -> h mod l, decompression of A and R, the addend table {O, B, -A, B - A}, then the 256 sequential
// rows acc' = 2 acc + T[bs + 2 bh] in 5 x 51-bit-limb field arithmetic; the canonical accumulator of every row is parked
// in global scratch (96 B per row) for phase 2.  Slots >= n_max are padding: zero scalars, addend O.
__global__ void __launch_bounds__(32) ed25519_ladder_kernel(WitnessArgs a, EdSlotInfo* __restrict__ info, ge_acc_packed* __restrict__ points) {
    __shared__ Sha512Hist h5[2];
    __shared__ EdTriple triple;
    __shared__ uint8_t digest[64];
    const uint32_t i = blockIdx.x;
    if (threadIdx.x != 0) return;
    EdSlotInfo* e = info + i;
    ge_cached51 tab[4];
    ge51 R = ge_identity51();
    if (i < a.n_max) {
        effective_triple(blob_validators(a.blob) + i, &triple);
        sha512_validator_prepare(triple, h5, digest);
        ed_slot_prepare(triple, digest, e, tab, &R);
    } else {
        for (int k = 0; k < 4; k++) e->s[k] = e->h[k] = 0;
        ed_padding_table(tab, &e->tab);
        e->ok = 1;
    }
    const ge_acc51 q = ed_straus_ladder(e->s, e->h, tab, points + (size_t)i * ED_ROWS_PER_VALIDATOR);
    e->QX = fe_freeze(q.X); e->QY = fe_freeze(q.Y); e->QZ = fe_freeze(q.Z);
    if (i < a.n_max) {
        const bool ok = e->ok && ed_result_equals(q, R);
        e->ok = ok ? 1u : 0u;
        a.aux[AUX_SIG_OK + i] = ok ? 1 : 0;
    }
}

// Phase 2 (throughput): one thread per trace row -- the 2 x 128 SHA-512 rows and the 256 Ed25519 rows of a slot.
__global__ void __launch_bounds__(256) ed25519_expand_kernel(WitnessArgs a, const EdSlotInfo* __restrict__ info,
                                                            const ge_acc_packed* __restrict__ points) {
    __shared__ Sha512Hist h5[2];
    __shared__ EdAddendTable tab;
    const uint32_t i = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        if (i < a.n_max) {
            EdTriple t;
            uint8_t digest[64];
            effective_triple(blob_validators(a.blob) + i, &t);
            sha512_validator_prepare(t, h5, digest);
        } else {
            sha512_padding_prepare(&h5[0]);
            sha512_padding_prepare(&h5[1]);
        }
    }
    for (int k = tid; k < 4 * 48; k += 256) tab.limb[k / 48][k % 48] = info[i].tab.limb[k / 48][k % 48];
    __syncthreads();
    const size_t row = (size_t)i * ED_ROWS_PER_VALIDATOR + tid;
    if (a.t512 && row < a.n512) sha512_row_cells(a.t512, a.n512, row, tid % S512_ROWS_PER_CHUNK, &h5[tid / S512_ROWS_PER_CHUNK]);
    if (a.ted && row < a.ned) ed_row_cells(a.ted, a.ned, row, (int)tid, info[i].s, info[i].h, points[row], tab);
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

size_t witness_slot_count(uint32_t n_max) { return pow2_at_least((size_t)n_max * ED_ROWS_PER_VALIDATOR) / ED_ROWS_PER_VALIDATOR; }
size_t witness_points_bytes(uint32_t n_max) {
    const size_t slots = witness_slot_count(n_max);
    return slots * sizeof(EdSlotInfo) + slots * ED_ROWS_PER_VALIDATOR * sizeof(ge_acc_packed);
}

// phase 1 only (stream-ordered); points must hold witness_points_bytes(n_max)
int run_ed25519_ladder(tmx_ctx* ctx, const WitnessArgs& a, void* points, cudaStream_t st) {
    const size_t slots = witness_slot_count(a.n_max);
    EdSlotInfo* info = (EdSlotInfo*)points;
    ed25519_ladder_kernel<<<(unsigned)slots, 32, 0, st>>>(a, info, (ge_acc_packed*)(info + slots));
    ctx->launches++;
    TMX_CUDA(cudaGetLastError());
    return TMX_OK;
}

// phase 2: all rows of the SHA-512 and Ed25519 tables (padding slots included)
int run_ed25519_expand(tmx_ctx* ctx, const WitnessArgs& a, const void* points, cudaStream_t st) {
    const size_t slots = witness_slot_count(a.n_max);
    const EdSlotInfo* info = (const EdSlotInfo*)points;
    ed25519_expand_kernel<<<(unsigned)slots, 256, 0, st>>>(a, info, (const ge_acc_packed*)(info + slots));
    ctx->launches++;
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

// K7 alone: the SHA-512 table (h = SHA-512(R || A || M) of every validator slot), without the Ed25519 rows
extern "C" int tmx_sha512_trace(tmx_ctx* ctx, const uint8_t* d_blob, uint32_t kind, uint32_t n_max, uint64_t* d_t512, void* stream) {
    if (!ctx || !d_blob || !d_t512) return fail(TMX_E_INPUT, "tmx_sha512_trace: NULL argument");
    WitnessArgs a;
    int rc = make_args(ctx, d_blob, kind, n_max, nullptr, d_t512, nullptr, nullptr, &a);
    if (rc) return rc;
    void* points = nullptr;
    rc = ctx_scratch(ctx, 1, witness_points_bytes(a.n_max), &points);
    if (rc) return rc;
    return run_ed25519_expand(ctx, a, points, pick_stream(ctx, stream));
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
