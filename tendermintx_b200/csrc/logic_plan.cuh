// Row schedule of the logic table for a circuit shape, and the host-side witness of that table (logic_plan.cu).
#pragma once
#include "logic.cuh"
#include "witness_jobs.cuh"
#include <memory>
#include <vector>

namespace tmx {

struct LogicPlan {
    AirShape sh;
    size_t n_rows = 0, used_rows = 0;
    std::vector<uint64_t> K;  // constant columns [LGK_COLS][n_rows]
    size_t row_glob = 0, row_cfe = 0, row_leaf[2] = {0, 0}, row_inner[2] = {0, 0}, row_hdr = 0, row_slots = 0;
    uint64_t& k(int col, size_t row) { return K[(size_t)col * n_rows + row]; }
};

std::shared_ptr<const LogicPlan> logic_plan_get(AirShape sh);
// Fills the table ([LG_COLS][n_rows], column-major, zero-initialised) from the proof's inputs and the per-slot results of
// the sequential Ed25519 phase.  Returns 0 or the id of the check that cannot be satisfied.
int logic_fill_trace(const LogicPlan& plan, const uint8_t* input, const uint8_t* blob, const EdSlotInfo* slots, gl* trace, bool force = false);
void logic_slot_infos_host(const uint8_t* blob, uint32_t n_max, std::vector<EdSlotInfo>& out);

}  // namespace tmx
