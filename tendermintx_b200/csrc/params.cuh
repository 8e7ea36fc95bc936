// Protocol parameters of the STARK layer (Curta's configuration as recalled in SURVEY.md App. C: rate_bits 1, cap height 4,
// 16-bit proof of work, 84 queries, arity-16 FRI, final polynomial <= 2^5 coefficients).
#pragma once
#include <cstdint>

namespace tmx {

constexpr unsigned STARK_RATE_BITS = 1;
constexpr unsigned STARK_CAP_HEIGHT = 4;
constexpr unsigned STARK_POW_BITS = 16;
constexpr int STARK_NUM_QUERIES = 84;
constexpr unsigned STARK_ARITY_BITS = 4;
constexpr unsigned STARK_FINAL_POLY_BITS = 5;
constexpr uint64_t STARK_PROOF_MAGIC = 0x32504D54ULL;    // "TMP2": proof format of the bus protocol
constexpr uint64_t STARK_CIRCUIT_MAGIC = 0x32434D54ULL;  // "TMC2": build artefact

}  // namespace tmx
