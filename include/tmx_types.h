/*
 * tmx_types.h -- plain-data layout of one proof's off-chain inputs ("blob"), shared by the C ABI
 * (include/tmx.h), the C++ host and the CPU oracle.  It is the by-value image of the reference's
 * VerifySkipVariable / VerifyStepVariable witness records [REF circuits/variables.rs:69-119], i.e. what
 * the async hints SkipOffchainInputs / StepOffchainInputs write [REF circuits/skip.rs:64-101,
 * circuits/step.rs:56-88].  All integers little-endian, structs packed, no pointers.
 *
 * blob = tmx_offchain_head, then tmx_validator[n_max], then (skip only) tmx_hash_field[n_max].
 */
#ifndef TMX_TYPES_H
#define TMX_TYPES_H

#include <stdint.h>

#define TMX_KIND_STEP 0u
#define TMX_KIND_SKIP 1u
#define TMX_BLOB_MAGIC 0x31584D54u /* "TMX1" */

/* sizes from REF circuits/consts.rs:4-37 */
#define TMX_HASH_SIZE 32
#define TMX_PROTOBUF_CHAIN_ID_SIZE_BYTES 52
#define TMX_PROTOBUF_HASH_SIZE_BYTES 34
#define TMX_PROTOBUF_BLOCK_ID_SIZE_BYTES 72
#define TMX_HEADER_PROOF_DEPTH 4
#define TMX_VALIDATOR_BYTE_LENGTH_MAX 46
#define TMX_VARINT_BYTES_LENGTH_MAX 9
#define TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX 124
#define TMX_CHAIN_ID_INDEX 1
#define TMX_BLOCK_HEIGHT_INDEX 2
#define TMX_LAST_BLOCK_ID_INDEX 4
#define TMX_VALIDATORS_HASH_INDEX 7
#define TMX_NEXT_VALIDATORS_HASH_INDEX 8
#define TMX_SKIP_MAX_DEFAULT 100800u /* REF circuits/config.rs:12 */

#pragma pack(push, 1)

/* ValidatorVariable, REF circuits/variables.rs:69-79 */
typedef struct {
    uint8_t pubkey[32];     /* compressed Edwards y, little endian, sign of x in bit 255 */
    uint8_t sig_r[32];      /* EDDSASignatureVariable.r */
    uint8_t sig_s[32];      /* EDDSASignatureVariable.s as u256 little endian */
    uint8_t message[TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX]; /* sign-bytes, zero padded */
    uint32_t message_byte_length;
    uint64_t voting_power;
    uint32_t validator_byte_length;
    uint8_t is_signed;
    uint8_t pad[3];
} tmx_validator; /* 240 bytes */

/* ValidatorHashFieldVariable, REF circuits/variables.rs:82-88 */
typedef struct {
    uint8_t pubkey[32];
    uint64_t voting_power;
    uint32_t validator_byte_length;
    uint32_t pad;
} tmx_hash_field; /* 48 bytes */

/* HashInclusionProofVariable = MerkleInclusionProofVariable<4, 34>, REF circuits/variables.rs:61-62 */
typedef struct {
    uint8_t leaf[TMX_PROTOBUF_HASH_SIZE_BYTES];
    uint8_t pad[2];
    uint8_t aunts[TMX_HEADER_PROOF_DEPTH][32];
} tmx_hash_proof; /* 164 bytes */

/* BlockIDInclusionProofVariable = MerkleInclusionProofVariable<4, 72>, REF circuits/variables.rs:63-64 */
typedef struct {
    uint8_t leaf[TMX_PROTOBUF_BLOCK_ID_SIZE_BYTES];
    uint8_t aunts[TMX_HEADER_PROOF_DEPTH][32];
} tmx_block_id_proof; /* 200 bytes */

/* ChainIdProofVariable, REF circuits/variables.rs:35-41 */
typedef struct {
    uint8_t aunts[TMX_HEADER_PROOF_DEPTH][32];
    uint32_t enc_chain_id_byte_length;
    uint8_t chain_id[TMX_PROTOBUF_CHAIN_ID_SIZE_BYTES];
} tmx_chain_id_proof; /* 184 bytes */

/* HeightProofVariable, REF circuits/variables.rs:50-56 */
typedef struct {
    uint8_t aunts[TMX_HEADER_PROOF_DEPTH][32];
    uint32_t enc_height_byte_length;
    uint32_t pad;
    uint64_t height;
} tmx_height_proof; /* 144 bytes */

/* VerifySkipVariable / VerifyStepVariable minus the arrays, REF circuits/variables.rs:91-119 */
typedef struct {
    uint32_t magic;
    uint32_t kind;             /* TMX_KIND_STEP / TMX_KIND_SKIP */
    uint32_t n_max;            /* VALIDATOR_SET_SIZE_MAX of the circuit */
    uint32_t nb_validators;    /* target_block_nb_validators / next_block_nb_validators */
    uint32_t nb_trusted;       /* trusted_block_nb_validators (skip) */
    uint32_t pad;
    uint64_t round;            /* target_block_round / next_block_round */
    uint8_t header[32];        /* target_header / next_header */
    tmx_chain_id_proof chain_id_proof;
    tmx_height_proof height_proof;
    tmx_hash_proof validators_hash_proof;
    /* skip: trusted_header_validator_hash_proof; step: prev_header_next_validators_hash_proof */
    tmx_hash_proof aux_hash_proof;
    tmx_block_id_proof last_block_id_proof; /* step: next_header_last_block_id_proof; skip: zero */
} tmx_offchain_head; /* 920 bytes */

#pragma pack(pop)

#define TMX_BLOB_SIZE(kind, n_max)                                                 \
    (sizeof(tmx_offchain_head) + (size_t)(n_max) * sizeof(tmx_validator) +          \
     ((kind) == TMX_KIND_SKIP ? (size_t)(n_max) * sizeof(tmx_hash_field) : (size_t)0))

#endif
