// Logic table: the plain-gate gadgets of verify_skip / verify_step [REF circuits/builder/verify.rs:137-563,
// validator.rs:73-253, shared.rs:43-215, voting.rs:29-110] and the glue between the hash / signature tables, as one more
// AIR on the shared bus.  Also the table-generic accessors (rows, columns, helper counts, constant columns, AIR dispatch)
// the prover, the verifier and the circuit builder use for all five tables.
#pragma once
#include "air.cuh"

namespace tmx {

// (not built yet: the table is absent and the cross-table links are off, see TMX_BUS_LINKS in air.cuh)
TMX_HD size_t logic_rows(AirShape) { return 0; }
TMX_HD int logic_cols(AirShape) { return 0; }
TMX_HD int logic_const_cols(AirShape) { return 0; }
TMX_HD int logic_helpers(AirShape) { return 0; }
TMX_HD uint64_t logic_const_value(int, size_t, AirShape) { return 0; }
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_logic(AirShape, const Row&, const Row&, const KRow&, const Per&, Emit&, Bus&) {}

// The verifier's own bus terms: the messages that tie the tables to the public input (trusted height / header, target
// height) and to the public output (the proven header).
inline gl2 logic_public_terms(AirShape, uint64_t, const uint8_t*, const uint8_t*, gl2, gl2) { return gl2_from(0); }

// ---- table-generic accessors ----
TMX_HD size_t air_table_rows(int t, AirShape sh) { return t == AIR_LOGIC ? logic_rows(sh) : air_rows(t, sh); }
TMX_HD int air_table_cols(int t, AirShape sh) { return t == AIR_LOGIC ? logic_cols(sh) : air_cols(t); }
TMX_HD int air_table_const_cols(int t, AirShape sh) { return t == AIR_LOGIC ? logic_const_cols(sh) : air_const_cols(t); }
TMX_HD int air_table_helpers(int t, AirShape sh) { return t == AIR_LOGIC ? logic_helpers(sh) : air_helpers(t); }
TMX_HD int air_table_aux_cols(int t, AirShape sh) { return air_table_cols(t, sh) ? 2 * (air_table_helpers(t, sh) + 1) : 0; }
TMX_HD uint64_t air_table_const_value(int t, int kc, size_t row, AirShape sh) {
    return t == AIR_LOGIC ? logic_const_value(kc, row, sh) : air_const_value(t, kc, row, sh);
}
template <class F, class Row, class KRow, class Per, class Emit, class Bus>
TMX_HD void air_eval_any(int table, AirShape sh, const Row& l, const Row& n, const KRow& k, const Per& per, Emit& emit, Bus& bus) {
    if (table == AIR_LOGIC) air_logic<F>(sh, l, n, k, per, emit, bus);
    else air_eval<F>(table, l, n, k, per, emit, bus);
}

}  // namespace tmx
