/*
 * tmx.h -- C ABI of libtmx.so, the B200 (sm_100a) witness generator and Goldilocks prover for
 * TendermintX's skip / step circuits.
 *
 * The reference has no FFI of its own: the boundary it sits behind is plonky2x's Rust surface
 * (`Circuit::define`, `builder.build()`, `circuit.prove()`, `circuit.verify()`, the `build` /
 * `prove input.json` CLI).  Each entry point below names the reference interface it replaces
 * (paths relative to /root/reference).  INTEGRATION.md shows the Rust `extern "C"` block a
 * maintainer would add.
 *
 * Conventions
 *   - every function returns a tmx_status (0 = TMX_OK); tmx_last_error() gives a thread-local message.
 *   - handles are opaque, created and freed by the library; caller-owned input buffers are only read
 *     during the call; output buffers are caller-allocated with explicit capacity.
 *   - "d_" parameters are DEVICE pointers in the context's GPU; `stream` is a cudaStream_t passed as
 *     void* (NULL = the context's own stream).  Field elements are canonical Goldilocks u64
 *     (little-endian), matrices are COLUMN-MAJOR: element (row, col) at col * n_rows + row.
 *   - one tmx_ctx per GPU, used by one host thread at a time; distinct contexts are independent.
 *   - nothing here falls back to a CPU implementation: without a CUDA device tmx_ctx_create fails.
 */
#ifndef TMX_H
#define TMX_H

#include <stddef.h>
#include <stdint.h>
#include "tmx_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TMX_OK = 0,
    TMX_E_INPUT = 1, /* malformed argument / input bytes / fixture (reference: panic on expect/assert) */
    TMX_E_UNSAT = 2, /* witness does not satisfy the circuit (reference: generator panic)          */
    TMX_E_CUDA = 3,  /* CUDA runtime / kernel failure, or no device                                   */
    TMX_E_IO = 4,    /* file missing / unreadable (reference: fs::read_to_string(..).unwrap())       */
    TMX_E_VERIFY = 5 /* proof rejected (reference: circuit.verify panics)                            */
} tmx_status;

typedef struct tmx_ctx tmx_ctx;

const char *tmx_last_error(void);
const char *tmx_version(void);

/* One context per GPU: owns a stream, twiddle tables, Poseidon constants and scratch memory. */
int tmx_ctx_create(int device, tmx_ctx **out);
void tmx_ctx_destroy(tmx_ctx *ctx);
int tmx_ctx_sync(tmx_ctx *ctx);
/* the context's own cudaStream_t (for CUDA-event timing of calls that take no stream argument) */
void *tmx_ctx_stream(const tmx_ctx *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches counter) */
uint64_t tmx_ctx_launch_count(const tmx_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * Kernel-level entry points (parity tests, ncu, roofline).  They replace, on the GPU, loops that run
 * inside plonky2 / starkyx when the reference calls `circuit.prove()` [circuits/skip.rs:214,244;
 * circuits/step.rs:196,223]; SURVEY.md section 8 row a23 lists them.
 * ---------------------------------------------------------------------------------------------- */

/* K1a: in-place batched NTT / inverse NTT, natural order in and out (plonky2_field fft / ifft). */
int tmx_ntt(tmx_ctx *ctx, uint64_t *d_data, size_t n_cols, unsigned log_n, int inverse, void *stream);

/* K1b: PolynomialBatch low-degree extension: values[n_cols][n] -> iNTT -> zero-pad by 2^rate_bits ->
 * coset NTT (shift 7) -> out[n_cols][n << rate_bits] in BIT-REVERSED row order (the Merkle leaf order).
 * d_coeffs (optional, may be NULL) receives the coset-scaled coefficients c_i * 7^i, [n_cols][n]. */
int tmx_lde(tmx_ctx *ctx, const uint64_t *d_values, uint64_t *d_out, uint64_t *d_coeffs, size_t n_cols,
            unsigned log_n, unsigned rate_bits, void *stream);

/* K2: Poseidon Merkle tree over the rows of a column-major matrix; leaf j = hash_or_noop(row j).
 * d_digests receives 4 u64 per node, levels concatenated: level 0 (n_rows leaf digests), level 1, ...,
 * cap level (1 << cap_height digests).  tmx_merkle_digest_count gives the total node count. */
size_t tmx_merkle_digest_count(unsigned log_rows, unsigned cap_height);
int tmx_poseidon_merkle(tmx_ctx *ctx, const uint64_t *d_cols, size_t n_cols, unsigned log_rows,
                        unsigned cap_height, uint64_t *d_digests, void *stream);

/* Poseidon permutation of n independent 12-element states (known-answer tests). */
int tmx_poseidon_permute(tmx_ctx *ctx, uint64_t *d_states, size_t n, void *stream);
/* Host-side self check (no GPU): the device permutation's arithmetic compiled for the CPU.
 * variant 0 = plain formulation, 1 = the kernels' fast path (multiplier-free linear layer),
 * 2 = the formulation the host transcript uses. */
int tmx_host_poseidon_permute(uint64_t *states, size_t n, int variant);
/* ------------------------------------------------------------------------------------------------
 * Witness tables (layout: include/tmx_trace.h).  They replace the trace generation that runs inside
 * `curta_sha256_variable`, the Tendermint Merkle gadgets and `curta_eddsa_verify_sigs_conditional` when the
 * reference proves [circuits/builder/verify.rs:147,165,202,205,248-259,285,376; validator.rs:228,248;
 * shared.rs:194,197].  d_blob is a DEVICE copy of the off-chain input blob (tmx_types.h).
 * ---------------------------------------------------------------------------------------------- */

/* dims = {rows, cols} of the SHA-256, SHA-512 and Ed25519 tables for a circuit shape */
int tmx_trace_dims(uint32_t kind, uint32_t n_max, size_t dims[6]);
/* bytes of the small result buffer the witness kernels fill: computed validators hashes [2][32], the root
 * reached by each header proof [5][32] (order: see witness_jobs.cuh), then one byte per validator slot
 * (1 = [s]B == R + [h]A holds for the slot's effective key / signature / message). */
size_t tmx_witness_aux_bytes(uint32_t n_max);
/* K6: SHA-256 table (validator leaves, validator-set trees, header inclusion proofs) */
int tmx_sha256_trace(tmx_ctx *ctx, const uint8_t *d_blob, uint32_t kind, uint32_t n_max, uint64_t *d_t256,
                     uint8_t *d_aux, void *stream);
/* K7 alone: SHA-512(R || A || M) table of every validator slot */
int tmx_sha512_trace(tmx_ctx *ctx, const uint8_t *d_blob, uint32_t kind, uint32_t n_max, uint64_t *d_t512, void *stream);
/* K7 + K8 fused, one validator per CTA: SHA-512(R || A || M) table and the Ed25519 table (joint evaluation of
 * [s]B + [h](-A), one row per scalar-bit pair) */
int tmx_ed25519_trace(tmx_ctx *ctx, const uint8_t *d_blob, uint32_t kind, uint32_t n_max, uint64_t *d_t512,
                      uint64_t *d_ted, uint8_t *d_aux, void *stream);
/* the three witness tables filled by kernels (SHA-256, SHA-512, Ed25519); the logic table is tmx_logic_trace, the range table follows from the lookups */
int tmx_witness_generate(tmx_ctx *ctx, const uint8_t *d_blob, uint32_t kind, uint32_t n_max, uint64_t *d_t256,
                         uint64_t *d_t512, uint64_t *d_ted, uint8_t *d_aux, void *stream);

/* K3 (fold): one arity-16 FRI folding step in evaluation space [plonky2 fri/prover.rs fri_committed_trees, one layer].
 * d_in: 16 << log_cosets extension elements (interleaved (a0, a1), bit-reversed order) on the coset shift * <w>;
 * d_out: 1 << log_cosets folded values (bit-reversed) on shift^16 * <w^16>. */
int tmx_fri_fold(tmx_ctx *ctx, const uint64_t *d_in, unsigned log_cosets, uint64_t shift, const uint64_t beta[2],
                 uint64_t *d_out, void *stream);

/* K9: proof-of-work grind: the SMALLEST w such that Poseidon(state with state[pos] = w)[7] has `bits` leading
 * zero bits (plonky2 fri_proof_of_work uses rayon find_any, i.e. any witness).  `state` is a host pointer. */
int tmx_pow_grind(tmx_ctx *ctx, const uint64_t state[12], int pos, unsigned bits, uint64_t *witness, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Circuit-level surface.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tmx_circuit tmx_circuit;
typedef struct tmx_proof tmx_proof;

/* `SkipCircuit::<N, C>::define` + `builder.build()` [circuits/skip.rs:113-143,165-173; circuits/step.rs:100-127]:
 * kind = TMX_KIND_SKIP / TMX_KIND_STEP; VALIDATOR_SET_SIZE_MAX, chain id and SKIP_MAX are runtime parameters here
 * (const generics / TendermintConfig in the reference, circuits/config.rs:3-32). */
int tmx_circuit_build(tmx_ctx *ctx, uint32_t kind, uint32_t n_max, const char *chain_id, size_t chain_id_len,
                      uint64_t skip_max, tmx_circuit **out);
void tmx_circuit_free(tmx_circuit *circuit);
int tmx_circuit_digest(const tmx_circuit *circuit, uint64_t out[4]);
/* the `build` artefact (./build/main.circuit in succinct.json:8,15) */
int tmx_circuit_save(const tmx_circuit *circuit, const char *path);
int tmx_circuit_load(tmx_ctx *ctx, const char *path, tmx_circuit **out);

/* The build artefact as u64 words WITHOUT a GPU (host-only: symbolic run of the constraint templates, constant columns and
 * their commitment, digest): returns the word count and copies at most `cap` words to `out` (may be NULL). */
size_t tmx_circuit_artefact(uint32_t kind, uint32_t n_max, const char *chain_id, size_t chain_id_len, uint64_t skip_max,
                            uint64_t *out, size_t cap);
/* The logic table (TMX_T_LOGIC: every plain-gate gadget of verify_skip / verify_step, one gadget instance per row) of one
 * proof, computed on the host: this is the code tmx_prove runs while the GPU commits the hash tables.  Returns the number of
 * u64 cells ([columns][rows], column-major) and copies at most `cap` of them; *status = 0 or the id of the first failing
 * check (tmx_last_check numbering); with `force` the table is completed past a failing check (adversarial tests commit to
 * such tables; tmx_verify must reject the proofs). */
size_t tmx_logic_trace(uint32_t kind, uint32_t n_max, const char *chain_id, size_t chain_id_len, const uint8_t *input,
                       const uint8_t *blob, int force, uint64_t *out, size_t cap, int *status);
/* shape of table t (TMX_T_* in tmx_trace.h): {rows, first-round columns, constant columns, second-round columns}; rows = 0
 * when the table is absent from the circuit */
int tmx_circuit_table_shape(const tmx_circuit *circuit, int table, size_t out[4]);

/* Kernel-level entry points that need a circuit's constant / periodic columns (parity tests, ncu):
 * tmx_bus_count   histogram of the range lookups of one table's first-round trace (d_trace: [cols][rows], device):
 *                 d_hist[2^16 + 2^11 + 2^8] u32 (16-, 11-, 8-bit tables); *bad is set when a value is outside its table.
 * tmx_bus_aux     second commitment round of one table: helper columns of its bus interactions and the running sum for the
 *                 given bus challenges (extension elements as two u64); d_aux: [second-round columns][rows]; total: the
 *                 table's bus contribution.
 * tmx_quotient    K5: constraint quotient of one table on its LDE coset (rate_bits 1).  d_lde_main / d_lde_aux: LDEs
 *                 ([cols][2 rows], bit-reversed rows, as tmx_lde produces them) of the first- and second-round traces;
 *                 d_out = [2][2 rows] in NATURAL order, one row per constraint challenge alpha[i]:
 *                 sum_k alpha^(M-1-k) C_k(x) / (x^n - 1) over table, helper-column and running-sum constraints. */
int tmx_bus_count(tmx_circuit *circuit, int table, const uint64_t *d_trace, uint32_t *d_hist, int *bad, void *stream);
int tmx_bus_aux(tmx_circuit *circuit, int table, const uint64_t *d_trace, const uint64_t beta[2], const uint64_t gamma[2],
                uint64_t *d_aux, uint64_t total[2], void *stream);
int tmx_quotient(tmx_circuit *circuit, int table, const uint64_t *d_lde_main, const uint64_t *d_lde_aux, const uint64_t total[2],
                 const uint64_t beta[2], const uint64_t gamma[2], const uint64_t alpha[2], uint64_t *d_out, void *stream);

/* `circuit.prove(&input)` [circuits/skip.rs:213-214,238-244]: input = 48 (skip) / 40 (step) bytes in the
 * abi.encodePacked layout of circuits/skip.rs:120-122; blob = the off-chain inputs the async hint would fetch
 * [circuits/skip.rs:64-101] (tmx_types.h).  Returns TMX_E_UNSAT (tmx_last_check() = failing check id) where the
 * reference's witness generation would panic.  out32 = the proven target / next header hash. */
int tmx_prove(tmx_circuit *circuit, const uint8_t *input, size_t input_len, const uint8_t *blob, size_t blob_len,
              tmx_proof **proof, uint8_t out32[32]);
/* Device time (CUDA events on the proving stream) of the trace commitments inside the last tmx_prove of this circuit,
 * milliseconds: {LDE (K1) table 0, Merkle (K2) table 0, LDE 1, Merkle 1, LDE 2, Merkle 2}.  Measurement aid for the
 * roofline figures; no reference counterpart. */
int tmx_circuit_last_phase_ms(const tmx_circuit *c, float out[6]);
int tmx_last_check(void);
/* keep one proof's off-chain inputs resident in HBM; tmx_prove(..., blob = NULL, 0, ...) then proves from them */
int tmx_circuit_set_inputs(tmx_circuit *circuit, const uint8_t *blob, size_t blob_len);

/* Input assembly = the fixture mode of InputDataFetcher [circuits/input/mod.rs:188-282,316-523;
 * circuits/input/conversion.rs:59-178]: reads <dir>/<height>/commit.json and validators_<page>.json (the RPC JSON
 * shapes, 100 validators per page) and fills the blob.  Sanity checks that `expect`/`assert!` in the reference
 * (header hash mismatch, set larger than n_max, missing file) return TMX_E_INPUT / TMX_E_IO. */
int tmx_header_hash_from_fixture(const char *fixture_dir, uint64_t block, uint8_t out[32]);
int tmx_skip_inputs_from_fixture(const char *fixture_dir, uint32_t n_max, uint64_t trusted_block,
                                 const uint8_t trusted_hash[32], uint64_t target_block, uint8_t *blob, size_t cap);
int tmx_step_inputs_from_fixture(const char *fixture_dir, uint32_t n_max, uint64_t prev_block, const uint8_t prev_hash[32],
                                 uint8_t *blob, size_t cap);
/* Operator-side search [circuits/input/mod.rs:158-186 find_block_to_request; circuits/input/tendermint_utils.rs:444-482
 * is_valid_skip]: is a skip from start_block to target_block possible (more than 1/3 of the target set's voting power is
 * held by start-set validators present in the target commit), and the highest block <= max_end_block reachable from
 * start_block by halving the distance (start_block + 1 = "request a step instead").  Missing fixtures -> TMX_E_IO, as
 * the reference's `expect` would abort. */
int tmx_is_valid_skip_from_fixture(const char *fixture_dir, uint64_t start_block, uint64_t target_block, int *valid);
int tmx_find_block_to_request(const char *fixture_dir, uint64_t start_block, uint64_t max_end_block, uint64_t *block);
/* tmx_prove with the off-chain inputs fetched from a fixture directory (the complete `prove input.json` path) */
int tmx_prove_fixture(tmx_circuit *circuit, const uint8_t *input, size_t input_len, const char *fixture_dir,
                      tmx_proof **proof, uint8_t out32[32]);
size_t tmx_proof_size(const tmx_proof *proof);
/* serialised proof: little-endian u64 stream, layout documented in DESIGN.md "Proof format" */
int tmx_proof_bytes(const tmx_proof *proof, uint8_t *buf, size_t cap);
void tmx_proof_free(tmx_proof *proof);

/* ------------------------------------------------------------------------------------------------
 * Prover pool: `in_flight` independent provers on one GPU (own tmx_ctx, circuit buffers and host thread each), fed from
 * one FIFO queue.  The Fiat-Shamir host round trips of one proof are filled by the kernels of the others (+25 % proofs
 * per hour on a B200 at in_flight = 4); proofs are byte-identical to tmx_prove's.  Upstream this role is played by the
 * proving platform dispatching `prove` requests to workers [bin/tendermintx.rs:169-223].  submit / wait may be called
 * from any thread; wait returns the job's tmx_prove status (TMX_E_UNSAT + tmx_last_check() on the calling thread) and
 * hands over the proof (free it with tmx_proof_free).  blob = NULL proves from the inputs of tmx_pool_set_inputs. */
typedef struct tmx_pool tmx_pool;
int tmx_pool_create(int device, uint32_t kind, uint32_t n_max, const char *chain_id, size_t chain_id_len, uint64_t skip_max,
                    unsigned in_flight, tmx_pool **pool);
/* the same with every prover loading the `build` artefact (./build/main.circuit) instead of building the circuit */
int tmx_pool_create_from_artefact(int device, const char *path, unsigned in_flight, tmx_pool **pool);
void tmx_pool_destroy(tmx_pool *pool);
int tmx_pool_set_inputs(tmx_pool *pool, const uint8_t *blob, size_t blob_len);
int tmx_pool_submit(tmx_pool *pool, const uint8_t *input, size_t input_len, const uint8_t *blob, size_t blob_len,
                    uint64_t *ticket);
int tmx_pool_wait(tmx_pool *pool, uint64_t ticket, tmx_proof **proof, uint8_t out32[32]);
unsigned tmx_pool_in_flight(const tmx_pool *pool);
uint64_t tmx_pool_launch_count(const tmx_pool *pool);
int tmx_pool_last_phase_ms(const tmx_pool *pool, unsigned prover, float out[6]);

/* `circuit.verify(&proof, &input, &output)` [circuits/skip.rs:247]: CPU verifier, 0 = accepted. */
int tmx_verify(const tmx_circuit *circuit, const uint8_t *proof, size_t proof_len, const uint8_t *input,
               size_t input_len, const uint8_t out32[32]);
/* the same check from the bare circuit parameters; needs no GPU and no tmx_ctx */
int tmx_verify_params(uint32_t kind, uint32_t n_max, const char *chain_id, size_t chain_id_len, uint64_t skip_max,
                      const uint8_t *proof, size_t proof_len, const uint8_t *input, size_t input_len,
                      const uint8_t out32[32]);

#ifdef __cplusplus
}
#endif
#endif
