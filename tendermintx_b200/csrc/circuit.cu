// Circuit-level surface: build / prove / verify for the skip and step circuits.
//
// Mirrors, behind the C ABI, plonky2x's `Circuit::define` + `builder.build()` [REF circuits/skip.rs:113-143,173;
// circuits/step.rs:100-127], `circuit.prove()` [REF skip.rs:214,244] and `circuit.verify()` [REF skip.rs:247].
// The statement proved is `verify_skip` / `verify_step` [REF circuits/builder/verify.rs:469-563]: the hashing and
// signature work is witnessed by the GPU tables and proved by the STARK; the remaining gadget checks (skip
// distance, sign-bytes fields, voting thresholds, chain-id bytes, hash linkages) are evaluated here on the host
// from the input blob and the digests the witness kernels produced, and abort the proof with TMX_E_UNSAT exactly
// where the reference's witness generation would panic.
#include "stark.cuh"
#include "bus.cuh"
#include "witness_jobs.cuh"
#include "logic_plan.cuh"
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <memory>
#include <string>
#include <thread>

using namespace tmx;

struct tmx_circuit {
    tmx_ctx* ctx = nullptr;
    uint32_t kind = 0, n_max = 0;
    std::string chain_id;
    uint64_t skip_max = 0;
    size_t dims[6] = {0, 0, 0, 0, 0, 0};
    std::shared_ptr<const CircuitDef> def;  // table shapes, constant columns, constraint DAG, digest
    // device state reused across proofs
    gl* d_trace[3] = {nullptr, nullptr, nullptr};
    uint8_t* d_blob = nullptr;          // resident inputs (tmx_circuit_set_inputs)
    uint8_t* d_blob_job = nullptr;      // inputs of a proof that brings its own blob: the resident copy stays valid
    uint8_t* d_aux = nullptr;
    void* d_points = nullptr;           // Ed25519 slot info + accumulators (phase 1 -> phase 2)
    // logic table: filled on the host from the inputs and the slot infos of the sequential Ed25519 phase
    std::shared_ptr<const LogicPlan> plan;
    EdSlotInfo* h_slots = nullptr;      // pinned
    gl* h_logic = nullptr;              // pinned, [LG_COLS][rows]
    gl* d_logic = nullptr;
    cudaStream_t side = nullptr;        // second stream: the latency-bound sequential Ed25519 rows
    cudaEvent_t ev_inputs = nullptr, ev_ladder = nullptr;
    // The tables of one commitment round are independent until their caps meet in the transcript: each is committed on its
    // own stream (tstream[0] is the context's stream) so that the latency-bound pieces of one table (upper Merkle levels, the
    // 4096-row logic table, one-CTA scans) run under the throughput-bound kernels of another.
    cudaStream_t tstream[STARK_N_TABLES] = {};
    cudaEvent_t ev_counted[STARK_N_TABLES] = {}, ev_done[STARK_N_TABLES] = {}, ev_expand = nullptr, ev_fork = nullptr;
    std::vector<uint8_t> h_blob;  // host copy of the resident inputs (tmx_circuit_set_inputs)
    bool resident = false;
    Prover prover;
    float phase_ms[6] = {0, 0, 0, 0, 0, 0};  // per witness table of the last proof: LDE (K1), trace Merkle tree (K2), device time
};

struct tmx_proof {
    std::vector<gl> words;
};

namespace tmx {

static thread_local int g_last_check = 0;

static uint64_t be64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}

// check ids (same numbering as the reference-assertion map in INTEGRATION.md)
enum {
    CHECK_OK = 0, CHECK_SKIP_DISTANCE, CHECK_TRUSTED_HEADER_PROOF, CHECK_TRUSTED_VALHASH, CHECK_TRUSTED_THRESHOLD,
    CHECK_SIGNATURE, CHECK_VALHASH, CHECK_VALHASH_PROOF, CHECK_THRESHOLD, CHECK_SIGN_BYTES, CHECK_CHAIN_ID, CHECK_HEIGHT,
    CHECK_LAST_BLOCK_ID, CHECK_NEXT_VALHASH, CHECK_VOTING_OVERFLOW, CHECK_VARINT_SIGN, CHECK_ROUND_SIGN, CHECK_INPUT
};

// REF circuits/builder/voting.rs:31-109 and verify.rs:439-467
static int voting_threshold(const std::vector<uint64_t>& power, const std::vector<uint8_t>& group, size_t nb_enabled, uint64_t num,
                            uint64_t den, bool* gt) {
    uint64_t total = 0, acc = 0;
    for (size_t i = 0; i < power.size(); i++) {
        const uint64_t v = i < nb_enabled ? power[i] : 0;
        if (total + v < total) return CHECK_VOTING_OVERFLOW;
        total += v;
    }
    for (size_t i = 0; i < power.size(); i++) {
        const uint64_t v = group[i] ? power[i] : 0;
        if (acc + v < acc) return CHECK_VOTING_OVERFLOW;
        acc += v;
    }
    const uint64_t sa = acc * den, st = total * num;
    if (sa / den != acc || st / num != total) return CHECK_VOTING_OVERFLOW;
    *gt = sa > st;
    return CHECK_OK;
}

// REF circuits/builder/validator.rs:73-183
static bool sign_bytes_ok(const tmx_validator& v, bool enabled, const uint8_t header[32], uint64_t height, uint64_t round) {
    const uint8_t* m = v.message;
    uint8_t le[8];
    const bool hash_in_message = memcmp(m + (round == 0 ? 16 : 25), header, 32) == 0;
    const bool precommit = m[1] == 8 && m[2] == 2;
    for (int i = 0; i < 8; i++) le[i] = (uint8_t)(height >> (8 * i));
    const bool height_ok = memcmp(m + 4, le, 8) == 0;
    for (int i = 0; i < 8; i++) le[i] = (uint8_t)(round >> (8 * i));
    const bool round_ok = round == 0 || memcmp(m + 13, le, 8) == 0;
    const bool valid = v.is_signed && enabled && hash_in_message && precommit && height_ok && round_ok;
    return (v.is_signed != 0) == valid;
}

// The gadget checks of verify_header [REF verify.rs:224-334] given the digests computed on the GPU.
static int check_header(const tmx_circuit* c, const tmx_offchain_head* h, const tmx_validator* vals, const uint8_t* aux,
                        uint32_t set, const uint8_t* root_valhash, const uint8_t* root_chain, const uint8_t* root_height,
                        uint64_t height) {
    const size_t n = c->n_max;
    if (h->round >> 63) return CHECK_ROUND_SIGN;
    for (size_t i = 0; i < n; i++)
        if (vals[i].is_signed && vals[i].message_byte_length > TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX) return CHECK_SIGNATURE;
    for (size_t i = 0; i < n; i++)
        if (!aux[AUX_SIG_OK + i]) return CHECK_SIGNATURE;
    std::vector<uint64_t> power(n);
    std::vector<uint8_t> sgn(n);
    for (size_t i = 0; i < n; i++) {
        power[i] = vals[i].voting_power;
        sgn[i] = vals[i].is_signed != 0;
        if (power[i] >> 63) return CHECK_VARINT_SIGN;
    }
    if (memcmp(aux + AUX_SET_ROOT + 32 * set, h->validators_hash_proof.leaf + 2, 32)) return CHECK_VALHASH;
    if (memcmp(root_valhash, h->header, 32)) return CHECK_VALHASH_PROOF;
    bool gt = false;
    int rc = voting_threshold(power, sgn, h->nb_validators, 2, 3, &gt);
    if (rc) return rc;
    if (!gt) return CHECK_THRESHOLD;
    for (size_t i = 0; i < n; i++)
        if (!sign_bytes_ok(vals[i], i < h->nb_validators, h->header, height, h->round)) return CHECK_SIGN_BYTES;
    // chain id [REF verify.rs:180-222]
    if ((size_t)h->chain_id_proof.enc_chain_id_byte_length + 1 > 55) return CHECK_CHAIN_ID;
    if (memcmp(root_chain, h->header, 32)) return CHECK_CHAIN_ID;
    if (c->chain_id.size() + 2 > TMX_PROTOBUF_CHAIN_ID_SIZE_BYTES ||
        memcmp(h->chain_id_proof.chain_id + 2, c->chain_id.data(), c->chain_id.size()))
        return CHECK_CHAIN_ID;
    // height [REF shared.rs:169-207]
    if (h->height_proof.height >> 63) return CHECK_VARINT_SIGN;
    if ((size_t)h->height_proof.enc_height_byte_length + 1 > 55) return CHECK_HEIGHT;
    if (memcmp(root_height, h->header, 32)) return CHECK_HEIGHT;
    if (h->height_proof.height != height) return CHECK_HEIGHT;
    return CHECK_OK;
}

static int check_statement(const tmx_circuit* c, const uint8_t* input, const uint8_t* blob, const uint8_t* aux) {
    const tmx_offchain_head* h = blob_head(blob);
    const tmx_validator* vals = blob_validators(blob);
    const size_t n = c->n_max;
    const uint8_t* proots = aux + AUX_PROOF_ROOT;
    if (c->kind == TMX_KIND_SKIP) {
        const uint64_t trusted = be64(input), target = be64(input + 40);
        const uint8_t* trusted_header = input + 8;
        if (!(target > trusted + 1 && target <= trusted + c->skip_max)) return CHECK_SKIP_DISTANCE;  // REF verify.rs:508-526
        // verify_trusted_validators [REF verify.rs:361-437]
        const tmx_hash_field* tf = blob_hash_fields(blob, c->n_max);
        if (memcmp(proots, trusted_header, 32)) return CHECK_TRUSTED_HEADER_PROOF;
        std::vector<uint64_t> power(n);
        for (size_t i = 0; i < n; i++) {
            power[i] = tf[i].voting_power;
            if (power[i] >> 63) return CHECK_VARINT_SIGN;
        }
        if (memcmp(aux + AUX_SET_ROOT, h->aux_hash_proof.leaf + 2, 32)) return CHECK_TRUSTED_VALHASH;
        std::vector<uint8_t> flag(n, 0);
        // (only enabled slots of the trusted set: what the proof enforces, see logic.cuh; the reference loops over all of them)
        for (size_t i = 0; i < n; i++)
            if (vals[i].is_signed)
                for (size_t j = 0; j < n && j < h->nb_trusted; j++)
                    if (!memcmp(vals[i].pubkey, tf[j].pubkey, 32)) flag[j] = 1;
        bool gt = false;
        int rc = voting_threshold(power, flag, h->nb_trusted, 1, 3, &gt);
        if (rc) return rc;
        if (!gt) return CHECK_TRUSTED_THRESHOLD;
        return check_header(c, h, vals, aux, 1, proots + 32, proots + 64, proots + 96, target);
    }
    const uint64_t prev = be64(input);
    const uint8_t* prev_header = input + 8;
    int rc = check_header(c, h, vals, aux, 0, proots, proots + 32, proots + 64, prev + 1);
    if (rc) return rc;
    // REF verify.rs:137-178
    if (memcmp(proots + 96, h->header, 32)) return CHECK_LAST_BLOCK_ID;
    if (memcmp(h->last_block_id_proof.leaf + 2, prev_header, 32)) return CHECK_LAST_BLOCK_ID;
    if (memcmp(proots + 128, prev_header, 32)) return CHECK_NEXT_VALHASH;
    if (memcmp(h->aux_hash_proof.leaf + 2, h->validators_hash_proof.leaf + 2, 32)) return CHECK_NEXT_VALHASH;
    return CHECK_OK;
}

}  // namespace tmx

extern "C" int tmx_last_check(void) { return g_last_check; }
namespace tmx {
void set_last_check(int check) { g_last_check = check; }  // pool.cu hands a worker's verdict to the waiting thread
}
extern "C" int tmx_circuit_build(tmx_ctx* ctx, uint32_t kind, uint32_t n_max, const char* chain_id, size_t chain_id_len,
                                 uint64_t skip_max, tmx_circuit** out) {
    if (!ctx || !out || !chain_id || kind > 1 || n_max == 0 || n_max > 4096 || chain_id_len == 0 || chain_id_len > 50)
        return fail(TMX_E_INPUT, "tmx_circuit_build: bad arguments");
    std::unique_ptr<tmx_circuit, void (*)(tmx_circuit*)> c(new tmx_circuit(), tmx_circuit_free);
    c->ctx = ctx;
    c->kind = kind;
    c->n_max = n_max;
    c->chain_id.assign(chain_id, chain_id_len);
    c->skip_max = skip_max;
    tmx_trace_dims(kind, n_max, c->dims);
    c->def = circuit_def_get(kind, n_max, c->chain_id, skip_max);
    if (!c->def) return fail(TMX_E_INPUT, std::string("tmx_circuit_build: ") + tmx_last_error());
    cudaSetDevice(ctx->device);
    for (int t = 0; t < 3; t++) {
        cudaError_t e = cudaMalloc((void**)&c->d_trace[t], c->dims[2 * t] * c->dims[2 * t + 1] * sizeof(gl));
        if (e != cudaSuccess) return fail(TMX_E_CUDA, std::string("tmx_circuit_build: cudaMalloc: ") + cudaGetErrorString(e));
    }
    if (cudaMalloc((void**)&c->d_blob, TMX_BLOB_SIZE(kind, n_max)) != cudaSuccess ||
        cudaMalloc((void**)&c->d_blob_job, TMX_BLOB_SIZE(kind, n_max)) != cudaSuccess ||
        cudaMalloc((void**)&c->d_aux, aux_bytes(n_max)) != cudaSuccess ||
        cudaMalloc(&c->d_points, witness_points_bytes(n_max)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_inputs, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_ladder, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_expand, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess)
        return fail(TMX_E_CUDA, "tmx_circuit_build: cudaMalloc failed");
    c->tstream[0] = ctx->stream;
    for (int t = 0; t < STARK_N_TABLES; t++) {
        if ((t && cudaStreamCreateWithFlags(&c->tstream[t], cudaStreamNonBlocking) != cudaSuccess) ||
            cudaEventCreateWithFlags(&c->ev_counted[t], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_done[t], cudaEventDisableTiming) != cudaSuccess)
            return fail(TMX_E_CUDA, "tmx_circuit_build: stream / event creation failed");
    }
    cudaMemset(c->d_points, 0, witness_points_bytes(n_max));  // struct padding travels to the host with the slot infos
    const AirShape sh = air_shape(kind, n_max, chain_id, chain_id_len);
    if (logic_rows(sh)) {
        c->plan = logic_plan_get(sh);
        const size_t cells = (size_t)LG_COLS * c->plan->n_rows;
        if (cudaMallocHost((void**)&c->h_slots, (size_t)n_max * sizeof(EdSlotInfo)) != cudaSuccess ||
            cudaMallocHost((void**)&c->h_logic, cells * sizeof(gl)) != cudaSuccess ||
            cudaMalloc((void**)&c->d_logic, cells * sizeof(gl)) != cudaSuccess)
            return fail(TMX_E_CUDA, "tmx_circuit_build: logic table buffers");
    }
    const int rc = c->prover.setup(ctx, c->def);
    if (rc) return rc;
    *out = c.release();
    return TMX_OK;
}

extern "C" void tmx_circuit_free(tmx_circuit* c) {
    if (!c) return;
    if (c->ctx) cudaSetDevice(c->ctx->device);
    for (int t = 0; t < 3; t++)
        if (c->d_trace[t]) cudaFree(c->d_trace[t]);
    if (c->d_blob) cudaFree(c->d_blob);
    if (c->d_blob_job) cudaFree(c->d_blob_job);
    if (c->d_aux) cudaFree(c->d_aux);
    if (c->d_points) cudaFree(c->d_points);
    if (c->h_slots) cudaFreeHost(c->h_slots);
    if (c->h_logic) cudaFreeHost(c->h_logic);
    if (c->d_logic) cudaFree(c->d_logic);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ev_inputs) cudaEventDestroy(c->ev_inputs);
    if (c->ev_ladder) cudaEventDestroy(c->ev_ladder);
    if (c->ev_expand) cudaEventDestroy(c->ev_expand);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int t = 0; t < STARK_N_TABLES; t++) {
        if (t && c->tstream[t]) cudaStreamDestroy(c->tstream[t]);
        if (c->ev_counted[t]) cudaEventDestroy(c->ev_counted[t]);
        if (c->ev_done[t]) cudaEventDestroy(c->ev_done[t]);
    }
    c->prover.release();
    delete c;
}

extern "C" int tmx_circuit_digest(const tmx_circuit* c, uint64_t out[4]) {
    if (!c || !out || !c->def) return fail(TMX_E_INPUT, "tmx_circuit_digest: NULL argument");
    for (int i = 0; i < 4; i++) out[i] = c->def->digest[i];
    return TMX_OK;
}

// build artefact (./build/main.circuit [REF succinct.json:8,15]): the circuit definition as data -- table shapes,
// constant columns and the caps of their commitments, periodic columns, constraint DAG and bus program, digest
extern "C" int tmx_circuit_save(const tmx_circuit* c, const char* path) {
    if (!c || !path || !c->def) return fail(TMX_E_INPUT, "tmx_circuit_save: NULL argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(TMX_E_IO, std::string("tmx_circuit_save: cannot open ") + path);
    const std::vector<uint64_t> w = c->def->serialize();
    const bool ok = fwrite(w.data(), sizeof(uint64_t), w.size(), f) == w.size();
    fclose(f);
    return ok ? TMX_OK : fail(TMX_E_IO, "tmx_circuit_save: short write");
}

extern "C" int tmx_circuit_load(tmx_ctx* ctx, const char* path, tmx_circuit** out) {
    if (!ctx || !path || !out) return fail(TMX_E_INPUT, "tmx_circuit_load: NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(TMX_E_IO, std::string("tmx_circuit_load: cannot open ") + path);
    std::vector<uint64_t> w;
    uint64_t buf[4096];
    size_t got;
    while ((got = fread(buf, sizeof(uint64_t), 4096, f)) > 0) w.insert(w.end(), buf, buf + got);
    fclose(f);
    // the artefact authenticates itself (its digest covers shapes, caps and constraints); this build then re-derives the
    // definition from the parameters and requires the same digest, i.e. the file was produced by a compatible version
    auto parsed = circuit_def_parse(w.data(), w.size());
    if (!parsed) return fail(TMX_E_INPUT, std::string("tmx_circuit_load: ") + tmx_last_error());
    int rc = tmx_circuit_build(ctx, parsed->kind, parsed->n_max, parsed->chain_id.data(), parsed->chain_id.size(), parsed->skip_max, out);
    if (rc) return rc;
    if (memcmp(parsed->digest, (*out)->def->digest, sizeof parsed->digest)) {
        tmx_circuit_free(*out);
        *out = nullptr;
        return fail(TMX_E_INPUT, "tmx_circuit_load: digest mismatch (built by a different version?)");
    }
    return TMX_OK;
}

// Upload the off-chain inputs once; a following tmx_prove(..., blob = NULL, 0, ...) proves from the HBM-resident
// copy (what bench.py times as `value`; the host-buffer call is the `e2e` figure).
extern "C" int tmx_circuit_set_inputs(tmx_circuit* c, const uint8_t* blob, size_t blob_len) {
    if (!c || !blob) return fail(TMX_E_INPUT, "tmx_circuit_set_inputs: NULL argument");
    if (blob_len != TMX_BLOB_SIZE(c->kind, c->n_max)) return fail(TMX_E_INPUT, "tmx_circuit_set_inputs: blob size does not match the circuit");
    TMX_CUDA(cudaSetDevice(c->ctx->device));
    c->h_blob.assign(blob, blob + blob_len);
    TMX_CUDA(cudaMemcpyAsync(c->d_blob, c->h_blob.data(), blob_len, cudaMemcpyHostToDevice, c->ctx->stream));
    TMX_CUDA(cudaStreamSynchronize(c->ctx->stream));
    c->resident = true;
    return TMX_OK;
}

// The length fields of the blob index fixed-size buffers inside the witness kernels: reject anything out of range before a
// kernel sees it [REF circuits/consts.rs:4-37 for the bounds; the reference's fixed-size array types cannot hold more].
static bool blob_lengths_ok(const tmx_circuit* c, const uint8_t* blob) {
    const tmx_offchain_head* h = blob_head(blob);
    if (h->nb_validators > c->n_max || (c->kind == TMX_KIND_SKIP && h->nb_trusted > c->n_max)) return false;
    if (h->chain_id_proof.enc_chain_id_byte_length > TMX_PROTOBUF_CHAIN_ID_SIZE_BYTES) return false;
    if (h->height_proof.enc_height_byte_length > 1 + TMX_VARINT_BYTES_LENGTH_MAX) return false;
    const tmx_validator* v = blob_validators(blob);
    for (uint32_t i = 0; i < c->n_max; i++)
        if (v[i].validator_byte_length > TMX_VALIDATOR_BYTE_LENGTH_MAX || v[i].message_byte_length > TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX)
            return false;
    if (c->kind == TMX_KIND_SKIP) {
        const tmx_hash_field* f = blob_hash_fields(blob, c->n_max);
        for (uint32_t i = 0; i < c->n_max; i++)
            if (f[i].validator_byte_length > TMX_VALIDATOR_BYTE_LENGTH_MAX) return false;
    }
    return true;
}

extern "C" int tmx_prove(tmx_circuit* c, const uint8_t* input, size_t input_len, const uint8_t* blob, size_t blob_len,
                         tmx_proof** proof_out, uint8_t out32[32]) {
    if (!c || !input || !proof_out || !out32) return fail(TMX_E_INPUT, "tmx_prove: NULL argument");
    const bool use_resident = (blob == nullptr);
    if (use_resident) {
        if (!c->resident) return fail(TMX_E_INPUT, "tmx_prove: no resident inputs (call tmx_circuit_set_inputs first)");
        blob = c->h_blob.data();
        blob_len = c->h_blob.size();
    }
    g_last_check = 0;
    *proof_out = nullptr;
    const tmx_offchain_head* h = blob_head(blob);
    if (blob_len < sizeof(tmx_offchain_head) || h->magic != TMX_BLOB_MAGIC || h->kind != c->kind || h->n_max != c->n_max ||
        blob_len != TMX_BLOB_SIZE(c->kind, c->n_max) || input_len != (c->kind == TMX_KIND_SKIP ? 48u : 40u) ||
        !blob_lengths_ok(c, blob)) {
        g_last_check = CHECK_INPUT;
        return fail(TMX_E_INPUT, "tmx_prove: input / blob does not match the circuit shape");
    }
    tmx_ctx* ctx = c->ctx;
    TMX_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Prover& pr = c->prover;
    // whatever way this call ends, no stream of the circuit may still be working on its buffers when the next one starts
    struct Quiesce {
        tmx_circuit* c;
        bool armed = true;
        ~Quiesce() {
            if (!armed) return;
            cudaStreamSynchronize(c->side);
            for (int t = 0; t < STARK_N_TABLES; t++)
                if (c->tstream[t]) cudaStreamSynchronize(c->tstream[t]);
        }
    } quiesce{c};
    // ---- witness generation on the GPU ----
    const uint8_t* d_blob = c->d_blob;
    if (!use_resident) {
        TMX_CUDA(cudaMemcpyAsync(c->d_blob_job, blob, blob_len, cudaMemcpyHostToDevice, st));
        d_blob = c->d_blob_job;
    }
    TMX_CUDA(cudaMemsetAsync(c->d_aux, 0, aux_bytes(c->n_max), st));
    TMX_CUDA(cudaMemsetAsync(pr.d_hist, 0, (BUS_HIST_SIZE + 4) * sizeof(unsigned int), st));
    WitnessArgs wa;
    int rc = witness_make_args(ctx, d_blob, c->kind, c->n_max, c->d_trace[0], c->d_trace[1], c->d_trace[2], c->d_aux, &wa);
    if (rc) return rc;
    // side stream: the sequential Ed25519 rows (one thread per validator slot, milliseconds of latency) run next to the
    // SHA-256 table's witness generation and commitment
    TMX_CUDA(cudaEventRecord(c->ev_inputs, st));
    TMX_CUDA(cudaStreamWaitEvent(c->side, c->ev_inputs, 0));
    rc = run_ed25519_ladder(ctx, wa, c->d_points, c->side);
    if (rc) return rc;
    if (c->plan) TMX_CUDA(cudaMemcpyAsync(c->h_slots, c->d_points, (size_t)c->n_max * sizeof(EdSlotInfo), cudaMemcpyDeviceToHost, c->side));
    TMX_CUDA(cudaEventRecord(c->ev_ladder, c->side));
    rc = witness_run_sha256(ctx, wa, st);
    if (rc) return rc;
    memcpy(out32, h->header, 32);
    // ---- proof: header, round 1, round 2 on the common transcript, then one tail per table on its own fork ----
    std::unique_ptr<tmx_proof> p(new tmx_proof());
    std::vector<gl>& w = p->words;
    w.push_back(STARK_PROOF_MAGIC);
    w.push_back(c->kind);
    w.push_back(c->n_max);
    w.push_back(STARK_N_TABLES);
    for (int i = 0; i < 4; i++) {
        gl x = 0;
        for (int j = 0; j < 8; j++) x |= (gl)out32[8 * i + j] << (8 * j);
        w.push_back(x);
    }
    Challenger ch;
    transcript_init(c->def->digest, input, input_len, out32, ch);
    // round 1: one stream per table; the range table waits for every histogram, the transcript for every cap
    // TMX_SERIAL_TABLES=1 puts everything back on one stream and one thread (same proof bytes): per-kernel timings with
    // CUDA events (bench.py's roofline leg, profilers) are only meaningful when nothing else shares the GPU
    const bool serial = getenv("TMX_SERIAL_TABLES") != nullptr;
    cudaStream_t ts[STARK_N_TABLES];
    for (int t = 0; t < STARK_N_TABLES; t++) ts[t] = serial ? st : c->tstream[t];
    if ((rc = pr.count_lookups(ctx, AIR_SHA256, c->d_trace[0], st))) return rc;
    TMX_CUDA(cudaEventRecord(c->ev_counted[AIR_SHA256], st));
    if ((rc = pr.commit_main(ctx, AIR_SHA256, c->d_trace[0], st))) return rc;
    // second phase of the Ed25519 / SHA-512 witness, then both tables side by side
    TMX_CUDA(cudaStreamWaitEvent(ts[AIR_ED25519], c->ev_ladder, 0));
    if ((rc = run_ed25519_expand(ctx, wa, c->d_points, ts[AIR_ED25519]))) return rc;
    TMX_CUDA(cudaEventRecord(c->ev_expand, ts[AIR_ED25519]));
    if ((rc = pr.count_lookups(ctx, AIR_ED25519, c->d_trace[2], ts[AIR_ED25519]))) return rc;
    TMX_CUDA(cudaEventRecord(c->ev_counted[AIR_ED25519], ts[AIR_ED25519]));
    if ((rc = pr.commit_main(ctx, AIR_ED25519, c->d_trace[2], ts[AIR_ED25519]))) return rc;
    TMX_CUDA(cudaStreamWaitEvent(ts[AIR_SHA512], c->ev_expand, 0));
    if ((rc = pr.count_lookups(ctx, AIR_SHA512, c->d_trace[1], ts[AIR_SHA512]))) return rc;
    TMX_CUDA(cudaEventRecord(c->ev_counted[AIR_SHA512], ts[AIR_SHA512]));
    if ((rc = pr.commit_main(ctx, AIR_SHA512, c->d_trace[1], ts[AIR_SHA512]))) return rc;
    // logic table: the host fills it while the GPU works on the tables above (it needs the slot infos of the sequential phase)
    int logic_status = 0;
    if (c->plan) {
        TMX_CUDA(cudaEventSynchronize(c->ev_ladder));
        const size_t cells = (size_t)LG_COLS * c->plan->n_rows;
        memset(c->h_logic, 0, cells * sizeof(gl));
        logic_status = logic_fill_trace(*c->plan, input, blob, c->h_slots, c->h_logic, true);
        TMX_CUDA(cudaStreamWaitEvent(ts[AIR_LOGIC], c->ev_inputs, 0));  // the histogram is cleared on the main stream
        TMX_CUDA(cudaMemcpyAsync(c->d_logic, c->h_logic, cells * sizeof(gl), cudaMemcpyHostToDevice, ts[AIR_LOGIC]));
        if ((rc = pr.count_lookups(ctx, AIR_LOGIC, c->d_logic, ts[AIR_LOGIC]))) return rc;
        TMX_CUDA(cudaEventRecord(c->ev_counted[AIR_LOGIC], ts[AIR_LOGIC]));
        if ((rc = pr.commit_main(ctx, AIR_LOGIC, c->d_logic, ts[AIR_LOGIC]))) return rc;
    }
    for (int t = 0; t < AIR_RANGE; t++)
        if (t != AIR_LOGIC || c->plan) TMX_CUDA(cudaStreamWaitEvent(ts[AIR_RANGE], c->ev_counted[t], 0));
    if ((rc = pr.fill_range_trace(ctx, ts[AIR_RANGE]))) return rc;
    if ((rc = pr.commit_main(ctx, AIR_RANGE, pr.d_range_trace, ts[AIR_RANGE]))) return rc;
    for (int t = 1; t < STARK_N_TABLES; t++) {
        TMX_CUDA(cudaEventRecord(c->ev_done[t], ts[t]));
        TMX_CUDA(cudaStreamWaitEvent(st, c->ev_done[t], 0));
    }
    std::vector<uint8_t> aux(aux_bytes(c->n_max));
    TMX_CUDA(cudaMemcpyAsync(aux.data(), c->d_aux, aux.size(), cudaMemcpyDeviceToHost, st));
    bool range_ok = true;
    if ((rc = pr.finish_round1(ctx, ch, w, st, &range_ok))) return rc;
    for (int t = 0; t < 3; t++) {
        c->phase_ms[2 * t] = pr.lde_ms[t];
        c->phase_ms[2 * t + 1] = pr.merkle_ms[t];
    }
    // The gadget checks that need the kernels' digests: where the reference's witness generation would panic.  This is a
    // PRE-CHECK for the honest prover's benefit (an early, named failure): every one of these conditions is enforced by the
    // proof itself (logic table, bus, public terms), so a prover that skips it -- TMX_DEBUG_NO_PRECHECK=1 does exactly that,
    // for tests -- only produces a proof that tmx_verify rejects.
    const bool no_precheck = getenv("TMX_DEBUG_NO_PRECHECK") != nullptr;
    const int chk = no_precheck ? 0 : check_statement(c, input, blob, aux.data());
    if (no_precheck) {
        logic_status = 0;
        range_ok = true;
    }
    if (chk) {
        g_last_check = chk;
        return fail(TMX_E_UNSAT, "tmx_prove: witness does not satisfy the circuit (check id " + std::to_string(chk) + ")");
    }
    if (logic_status) {
        g_last_check = logic_status;
        return fail(TMX_E_UNSAT, "tmx_prove: witness does not satisfy the circuit (logic table, check id " + std::to_string(logic_status) + ")");
    }
    if (!range_ok) return fail(TMX_E_UNSAT, "tmx_prove: a witness cell is outside its range table");
    const gl2 beta = ch.get_ext(), gamma = ch.get_ext();
    // round 2
    const gl* traces[STARK_N_TABLES] = {c->d_trace[0], c->d_trace[1], c->d_trace[2], c->d_logic, pr.d_range_trace};
    // (every stream is idle here: finish_round1 synchronised the main stream, which had joined the others; biggest table first)
    static const int order[STARK_N_TABLES] = {AIR_ED25519, AIR_SHA512, AIR_SHA256, AIR_LOGIC, AIR_RANGE};
    for (int i = 0; i < STARK_N_TABLES; i++)
        if ((rc = pr.commit_aux(ctx, order[i], traces[order[i]], beta, gamma, ts[order[i]]))) return rc;
    for (int t = 1; t < STARK_N_TABLES; t++) {
        TMX_CUDA(cudaEventRecord(c->ev_done[t], ts[t]));
        TMX_CUDA(cudaStreamWaitEvent(st, c->ev_done[t], 0));
    }
    if ((rc = pr.finish_round2(ctx, ch, w, st))) return rc;
    // tails: every table continues on a FORK of the transcript (the common state after round 2 plus the table index), so the
    // five tails are independent of each other and run side by side, one host thread and one stream per table; their words
    // enter the proof in table order
    std::vector<gl> tail_words[STARK_N_TABLES];
    int tail_rc[STARK_N_TABLES] = {};
    std::string tail_err[STARK_N_TABLES];
    auto run_tail = [&](int t) {
        if (!c->def->tables[t].n_main) return;
        cudaSetDevice(ctx->device);
        Challenger fork = ch;
        fork.observe((gl)t);
        tail_rc[t] = pr.prove_tail(ctx, t, beta, gamma, fork, tail_words[t], ts[t]);
        if (tail_rc[t]) tail_err[t] = tmx_last_error();
    };
    if (serial) {
        for (int t = 0; t < STARK_N_TABLES; t++) run_tail(t);
    } else {
        std::thread workers[STARK_N_TABLES];
        bool started[STARK_N_TABLES] = {};
        for (int t = 0; t < STARK_N_TABLES; t++) {
            if (t == AIR_ED25519) continue;
            try {  // nothing may unwind across the C ABI: a table whose thread cannot be created runs on this one
                workers[t] = std::thread(run_tail, t);
                started[t] = true;
            } catch (...) {
            }
        }
        run_tail(AIR_ED25519);
        for (int t = 0; t < STARK_N_TABLES; t++)
            if (t != AIR_ED25519 && !started[t]) run_tail(t);
        for (auto& th : workers)
            if (th.joinable()) th.join();
    }
    for (int t = 0; t < STARK_N_TABLES; t++) {
        if (tail_rc[t]) return fail(tail_rc[t], tail_err[t]);
        w.insert(w.end(), tail_words[t].begin(), tail_words[t].end());
    }
    quiesce.armed = false;  // every tail ended with a synchronised copy on its stream; the side stream was joined in round 1
    *proof_out = p.release();
    return TMX_OK;
}

// Device time (CUDA events on the proving stream) of the trace commitments inside the LAST tmx_prove of this circuit:
// out = {LDE table 0, Merkle table 0, LDE table 1, Merkle table 1, LDE table 2, Merkle table 2} in milliseconds.
extern "C" int tmx_circuit_last_phase_ms(const tmx_circuit* c, float out[6]) {
    if (!c || !out) return fail(TMX_E_INPUT, "tmx_circuit_last_phase_ms: NULL argument");
    for (int i = 0; i < 6; i++) out[i] = c->phase_ms[i];
    return TMX_OK;
}

// `prove input.json` with the off-chain inputs taken from a fixture directory, i.e. the whole async-hint path
// [REF circuits/skip.rs:64-101]: parse heights from the public input, assemble the blob, prove.
extern "C" int tmx_skip_inputs_from_fixture(const char*, uint32_t, uint64_t, const uint8_t*, uint64_t, uint8_t*, size_t);
extern "C" int tmx_step_inputs_from_fixture(const char*, uint32_t, uint64_t, const uint8_t*, uint8_t*, size_t);
extern "C" int tmx_prove_fixture(tmx_circuit* c, const uint8_t* input, size_t input_len, const char* fixture_dir,
                                 tmx_proof** proof_out, uint8_t out32[32]) {
    if (!c || !input || !fixture_dir || !proof_out || !out32) return fail(TMX_E_INPUT, "tmx_prove_fixture: NULL argument");
    if (input_len != (c->kind == TMX_KIND_SKIP ? 48u : 40u)) return fail(TMX_E_INPUT, "tmx_prove_fixture: wrong public input length");
    std::vector<uint8_t> blob(TMX_BLOB_SIZE(c->kind, c->n_max));
    int rc = c->kind == TMX_KIND_SKIP
                 ? tmx_skip_inputs_from_fixture(fixture_dir, c->n_max, be64(input), input + 8, be64(input + 40), blob.data(), blob.size())
                 : tmx_step_inputs_from_fixture(fixture_dir, c->n_max, be64(input), input + 8, blob.data(), blob.size());
    if (rc) return rc;
    return tmx_prove(c, input, input_len, blob.data(), blob.size(), proof_out, out32);
}

extern "C" size_t tmx_proof_size(const tmx_proof* p) { return p ? p->words.size() * sizeof(gl) : 0; }

extern "C" int tmx_proof_bytes(const tmx_proof* p, uint8_t* buf, size_t cap) {
    if (!p || !buf) return fail(TMX_E_INPUT, "tmx_proof_bytes: NULL argument");
    if (cap < p->words.size() * sizeof(gl)) return fail(TMX_E_INPUT, "tmx_proof_bytes: buffer too small");
    memcpy(buf, p->words.data(), p->words.size() * sizeof(gl));  // u64 little-endian stream
    return TMX_OK;
}

extern "C" void tmx_proof_free(tmx_proof* p) { delete p; }

static int verify_with(const CircuitDef& def, const uint8_t* proof, size_t proof_len, const uint8_t* input, size_t input_len,
                       const uint8_t out32[32]) {
    if (proof_len % 8) return fail(TMX_E_VERIFY, "tmx_verify: malformed proof length");
    std::vector<gl> w(proof_len / 8);
    memcpy(w.data(), proof, proof_len);
    // checks that depend on public data only are the verifier's own [REF verify.rs:508-526 verify_skip_distance]
    if (def.kind == TMX_KIND_SKIP && input_len == 48) {
        const uint64_t trusted = be64(input), target = be64(input + 40);
        if (!(target > trusted + 1 && target <= trusted + def.skip_max)) return fail(TMX_E_VERIFY, "tmx_verify: skip distance out of range");
    }
    const int rc = verify_proof(def, w.data(), w.size(), input, input_len, out32);
    if (rc) return fail(TMX_E_VERIFY, "tmx_verify: proof rejected (code " + std::to_string(rc) + ")");
    return TMX_OK;
}

// Verification needs no GPU: the same check from the bare circuit parameters (what a light client would hold).
extern "C" int tmx_verify_params(uint32_t kind, uint32_t n_max, const char* chain_id, size_t chain_id_len, uint64_t skip_max,
                                 const uint8_t* proof, size_t proof_len, const uint8_t* input, size_t input_len,
                                 const uint8_t out32[32]) {
    if (!chain_id || !proof || !input || !out32 || kind > 1 || n_max == 0 || n_max > 4096 || chain_id_len == 0 || chain_id_len > 50)
        return fail(TMX_E_INPUT, "tmx_verify_params: bad arguments");
    auto def = circuit_def_get(kind, n_max, std::string(chain_id, chain_id_len), skip_max);
    if (!def) return fail(TMX_E_INPUT, std::string("tmx_verify_params: ") + tmx_last_error());
    return verify_with(*def, proof, proof_len, input, input_len, out32);
}

extern "C" int tmx_verify(const tmx_circuit* c, const uint8_t* proof, size_t proof_len, const uint8_t* input, size_t input_len,
                          const uint8_t out32[32]) {
    if (!c || !proof || !input || !out32 || !c->def) return fail(TMX_E_INPUT, "tmx_verify: NULL argument");
    return verify_with(*c->def, proof, proof_len, input, input_len, out32);
}

// ------------------------------------------------------------------------------------------------
// Kernel-level entry points that need a circuit's constant / periodic columns (parity tests, ncu).
// ------------------------------------------------------------------------------------------------
static bool table_ok(const tmx_circuit* c, int table) {
    return c && c->def && table >= 0 && table < STARK_N_TABLES && c->def->tables[table].n_main != 0;
}

// rows / first-round / constant / second-round column counts of table t (0 rows: the table is absent)
extern "C" int tmx_circuit_table_shape(const tmx_circuit* c, int table, size_t out[4]) {
    if (!c || !c->def || !out || table < 0 || table >= STARK_N_TABLES) return fail(TMX_E_INPUT, "tmx_circuit_table_shape: bad arguments");
    const TableDef& td = c->def->tables[table];
    out[0] = td.n_main ? td.rows() : 0;
    out[1] = td.n_main;
    out[2] = td.n_const;
    out[3] = (size_t)td.n_aux();
    return TMX_OK;
}

// Second commitment round of one table: helper columns of its bus interactions and the running sum, for given first-round
// trace (device, [cols][n]) and bus challenges.  d_aux: [aux cols][n]; total: the table's bus contribution (extension element).
extern "C" int tmx_bus_aux(tmx_circuit* c, int table, const uint64_t* d_trace, const uint64_t beta[2], const uint64_t gamma[2],
                           uint64_t* d_aux, uint64_t total[2], void* stream) {
    if (!table_ok(c, table) || !d_trace || !beta || !gamma || !d_aux || !total) return fail(TMX_E_INPUT, "tmx_bus_aux: bad arguments");
    TMX_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t st = pick_stream(c->ctx, stream);
    Prover& pr = c->prover;
    const TableDef& td = c->def->tables[table];
    int rc = pr.commit_aux(c->ctx, table, d_trace, gl2_make(beta[0], beta[1]), gl2_make(gamma[0], gamma[1]), st);
    if (rc) return rc;
    TMX_CUDA(cudaMemcpyAsync(d_aux, pr.tab[table].d_aux, (size_t)td.n_aux() * td.rows() * sizeof(gl), cudaMemcpyDeviceToDevice, st));
    TMX_CUDA(cudaMemcpyAsync(total, pr.d_small + 66 * table + 64, 2 * sizeof(gl), cudaMemcpyDeviceToHost, st));
    TMX_CUDA(cudaStreamSynchronize(st));
    return TMX_OK;
}

// Histogram of the range lookups of one table's first-round trace: d_hist[2^16 + 2^11 + 2^8] (u32; 16-, 11-, 8-bit tables);
// accumulates, so the caller zeroes it.  *bad is set when a value is outside its table.
extern "C" int tmx_bus_count(tmx_circuit* c, int table, const uint64_t* d_trace, uint32_t* d_hist, int* bad, void* stream) {
    if (!table_ok(c, table) || !d_trace || !d_hist || !bad) return fail(TMX_E_INPUT, "tmx_bus_count: bad arguments");
    TMX_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t st = pick_stream(c->ctx, stream);
    Prover& pr = c->prover;
    TMX_CUDA(cudaMemsetAsync(pr.d_hist, 0, (BUS_HIST_SIZE + 4) * sizeof(unsigned int), st));
    int rc = pr.count_lookups(c->ctx, table, d_trace, st);
    if (rc) return rc;
    unsigned int flag = 0;
    TMX_CUDA(cudaMemcpyAsync(d_hist, pr.d_hist, BUS_HIST_SIZE * sizeof(unsigned int), cudaMemcpyDeviceToDevice, st));
    TMX_CUDA(cudaMemcpyAsync(&flag, pr.d_hist + BUS_HIST_SIZE, sizeof flag, cudaMemcpyDeviceToHost, st));
    TMX_CUDA(cudaStreamSynchronize(st));
    *bad = flag != 0;
    return TMX_OK;
}

extern "C" int tmx_quotient(tmx_circuit* c, int table, const uint64_t* d_lde_main, const uint64_t* d_lde_aux, const uint64_t total[2],
                            const uint64_t beta[2], const uint64_t gamma[2], const uint64_t alpha[2], uint64_t* d_out, void* stream) {
    if (!table_ok(c, table) || !d_lde_main || !d_lde_aux || !total || !beta || !gamma || !alpha || !d_out)
        return fail(TMX_E_INPUT, "tmx_quotient: bad arguments");
    TMX_CUDA(cudaSetDevice(c->ctx->device));
    return stark_quotient(c->ctx, c->prover, table, d_lde_main, d_lde_aux, gl2_make(total[0], total[1]), gl2_make(beta[0], beta[1]),
                          gl2_make(gamma[0], gamma[1]), alpha, d_out, pick_stream(c->ctx, stream));
}
