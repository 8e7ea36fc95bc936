// tmx_ctx lifetime, error reporting, twiddle / power tables, Poseidon round-constant generation.
#include "ctx.cuh"
#include "poseidon.cuh"
#include <mutex>
#include <cstring>
#include <cstdio>
#include <cstdlib>

namespace tmx {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// ---- Poseidon round constants: ChaCha8 keystream keyed by rand-0.8's seed_from_u64(0) (PCG32 expansion),
// u64 = lo | hi << 32, mapped into [0, p) with rand's widening-multiply rejection sampler (zone = p - 1).
gl h_poseidon_rc[POSEIDON_ROUNDS * POSEIDON_WIDTH];
gl h_poseidon_rc_fast[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH];
static bool g_rc_ready = false;

static inline uint32_t rol(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void chacha8(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = (uint32_t)counter;
    in[13] = (uint32_t)(counter >> 32);
    in[14] = in[15] = 0;
    uint32_t x[16];
    memcpy(x, in, sizeof x);
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rol(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rol(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rol(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rol(x[b] ^ x[c], 7);
    };
    for (int dr = 0; dr < 4; dr++) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

static void poseidon_generate_constants_once() {
    uint64_t st = 0;
    uint32_t key[8];
    for (int i = 0; i < 8; i++) {
        st = st * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
        uint32_t rot = (uint32_t)(st >> 59);
        key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    uint32_t block[16];
    uint64_t ctr = 0;
    int pos = 16, n = 0;
    auto next32 = [&]() {
        if (pos == 16) {
            chacha8(key, ctr++, block);
            pos = 0;
        }
        return block[pos++];
    };
    while (n < POSEIDON_ROUNDS * POSEIDON_WIDTH) {
        uint64_t lo = next32();
        uint64_t hi = next32();
        unsigned __int128 m = (unsigned __int128)(lo | (hi << 32)) * GL_P;
        if ((uint64_t)m <= GL_P - 1) h_poseidon_rc[n++] = (gl)(m >> 64);
    }
    // fast-path table (poseidon.cuh): the constants of lanes 1..11 of the partial rounds travel through the linear layers
    for (int i = 0; i < (POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH; i++)
        h_poseidon_rc_fast[i] = i < POSEIDON_ROUNDS * POSEIDON_WIDTH ? h_poseidon_rc[i] : 0;
    const int first = POSEIDON_HALF_FULL, last = POSEIDON_HALF_FULL + POSEIDON_PARTIAL - 1;  // partial rounds first..last
    gl h[12];
    for (int i = 0; i < 12; i++) h[i] = h_poseidon_rc[12 * first + i];
    for (int r = first; r <= last; r++) {
        gl d[12];
        for (int i = 0; i < 12; i++) d[i] = i == 0 ? 0 : h[i];  // deferred part of this round's constants
        poseidon_mds(d);
        for (int i = 0; i < 12; i++) h[i] = gl_add(h_poseidon_rc[12 * (r + 1) + i], d[i]);
        if (r > first)
            for (int i = 1; i < 12; i++) h_poseidon_rc_fast[12 * r + i] = 0;
        for (int i = 0; i < 12; i++) h_poseidon_rc_fast[12 * (r + 1) + i] = h[i];
    }
    // poseidon_recombine() wants a non-negative integer: the lowest piece of a lane in piece form can be as low as -2^21, and
    // the constant added on top must cover that wherever such lanes are recombined (a fixed table: this never fires)
    for (int r = first + 1; r <= last + 1; r++)
        for (int i = 0; i < (r == last + 1 ? 12 : 1); i++)
            if (h_poseidon_rc_fast[12 * r + i] < ((gl)1 << 22)) {
                fprintf(stderr, "tmx: Poseidon fast-path constant %d/%d is too small for the piece arithmetic\n", r, i);
                abort();
            }
    g_rc_ready = true;
}
// tmx_verify_params may be the first caller on several threads at once: the table is written exactly once
void poseidon_generate_constants() {
    static std::once_flag once;
    std::call_once(once, poseidon_generate_constants_once);
}

// ---- device tables ----
static int upload(tmx_ctx* ctx, const std::vector<gl>& h, gl** out) {
    void* d = nullptr;
    TMX_CUDA(cudaMalloc(&d, h.size() * sizeof(gl)));
    // cudaMemcpy from pageable memory returns once the data is STAGED; the DMA runs in the legacy stream, which the
    // provers' non-blocking streams do not wait for.  With several provers in flight a kernel could read a table before
    // it had landed (seen as a wrong quotient cap in the first proof of a fresh context).  Wait for the DMA.
    TMX_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(gl), cudaMemcpyHostToDevice));
    TMX_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    ctx->owned.push_back(d);
    *out = (gl*)d;
    return TMX_OK;
}

// scale * base^e for e < 2^log_size
static int build_pow_table(tmx_ctx* ctx, gl base, unsigned log_size, gl scale, gl** out) {
    std::vector<gl> t((size_t)1 << log_size);
    gl cur = scale;
    for (size_t i = 0; i < t.size(); i++) {
        t[i] = cur;
        cur = gl_mul(cur, base);
    }
    return upload(ctx, t, out);
}

int ctx_ntt_tables(tmx_ctx* ctx, unsigned log_n, bool inverse, const NttTables** out) {
    std::lock_guard<std::mutex> lock(ctx->tables_mu);
    auto& cache = inverse ? ctx->inv : ctx->fwd;
    auto it = cache.find(log_n);
    if (it != cache.end()) {
        *out = &it->second;
        return TMX_OK;
    }
    NttTables t;
    gl w = gl_root_of_unity(log_n);
    if (inverse) w = gl_inv(w);
    int rc = build_pow_table(ctx, w, log_n, 1, &t.full);
    if (rc) return rc;
    t.small = t.full;  // the first half of the full table is exactly w^e, e < 2^(k-1)
    auto ins = cache.emplace(log_n, t);
    *out = &ins.first->second;
    return TMX_OK;
}

int ctx_coset_scale(tmx_ctx* ctx, unsigned log_n, const gl** out) {
    std::lock_guard<std::mutex> lock(ctx->tables_mu);
    auto it = ctx->coset_scale.find(log_n);
    if (it != ctx->coset_scale.end()) {
        *out = it->second;
        return TMX_OK;
    }
    gl* t = nullptr;
    int rc = build_pow_table(ctx, GL_GEN, log_n, gl_inv((gl)((uint64_t)1 << log_n)), &t);
    if (rc) return rc;
    ctx->coset_scale.emplace(log_n, t);
    *out = t;
    return TMX_OK;
}

int ctx_scratch(tmx_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            TMX_CUDA(cudaStreamSynchronize(ctx->stream));
            TMX_CUDA(cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 8;
        TMX_CUDA(cudaMalloc(&ctx->scratch[slot], want));
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;

extern "C" const char* tmx_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* tmx_version(void) { return "tmx-b200 0.1 (sm_100a)"; }

extern "C" int tmx_ctx_create(int device, tmx_ctx** out) {
    if (!out) return fail(TMX_E_INPUT, "tmx_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(TMX_E_CUDA, std::string("tmx_ctx_create: no CUDA device (") + cudaGetErrorString(e) +
                                    "); this library has no CPU fallback");
    if (device < 0 || device >= count) return fail(TMX_E_INPUT, "tmx_ctx_create: device index out of range");
    TMX_CUDA(cudaSetDevice(device));
    tmx_ctx* ctx = new tmx_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    TMX_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        delete ctx;
        return fail(TMX_E_CUDA, "tmx_ctx_create: device is not sm_100 class (kernels are built for sm_100a only)");
    }
    TMX_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    poseidon_generate_constants();
    int rc = merkle_tu_init();
    if (!rc) rc = witness_tu_init();
    if (rc) {
        delete ctx;
        return rc;
    }
    *out = ctx;
    return TMX_OK;
}

extern "C" void tmx_ctx_destroy(tmx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (void* p : ctx->owned) cudaFree(p);
    for (int i = 0; i < 4; i++)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int tmx_ctx_sync(tmx_ctx* ctx) {
    if (!ctx) return fail(TMX_E_INPUT, "tmx_ctx_sync: ctx is NULL");
    TMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return TMX_OK;
}

extern "C" void* tmx_ctx_stream(const tmx_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" uint64_t tmx_ctx_launch_count(const tmx_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
