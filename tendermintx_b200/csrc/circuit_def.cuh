// The circuit as DATA: what `Circuit::define` + `builder.build()` leave behind in the reference
// [REF circuits/skip.rs:113-143,165-173; circuits/step.rs:100-127] -- table shapes, the constant (preprocessed) columns with
// the Merkle caps of their low-degree extensions, the periodic columns, and the constraint system of every table as an
// expression DAG with its bus interactions.  The DAG is obtained by running the AIR templates of air.cuh on symbolic
// values (Sym), so there is ONE definition of the constraints: the GPU kernels and the verifier compile it, the build
// artefact (./build/main.circuit) carries it as data, and the CPU oracle interprets that data with its own engine.
// Host only.
#pragma once
#include "air.cuh"
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace tmx {

enum SymOp : uint32_t { SYM_CONST = 0, SYM_COL = 1, SYM_ADD = 2, SYM_SUB = 3, SYM_MUL = 4 };
enum SymSrc : uint32_t { SRC_LOCAL = 0, SRC_NEXT = 1, SRC_CONST = 2, SRC_PERIODIC = 3 };
struct SymNode {
    uint32_t op, a, b, deg;  // COL: a = source, b = column
    uint64_t val;            // CONST
};
// one item of a table's constraint program, flattened to u64 words in `prog`:
//   {0, node}                                       constraint  node == 0
//   {1, tag, m, len, v_0 .. v_{len-1}}              bus.one
//   {2, tag, m, len, v.., tag, m, len, v..}         bus.two
struct TableDef {
    uint32_t log_n = 0, n_main = 0, n_const = 0, n_per = 0, period = 1, n_helpers = 0, n_constraints = 0;
    std::vector<gl> periodic;    // [n_per][period]
    std::vector<gl> constants;   // [n_const][n], column-major
    std::vector<gl> const_cap;   // Merkle cap of the constant columns' LDE (4 words per digest)
    std::vector<SymNode> nodes;
    std::vector<uint64_t> prog;
    size_t rows() const { return (size_t)1 << log_n; }
    int n_aux() const { return n_main ? 2 * ((int)n_helpers + 1) : 0; }
};

struct CircuitDef {
    uint32_t kind = 0, n_max = 0;
    uint64_t skip_max = 0;
    std::string chain_id;
    TableDef tables[TMX_N_TABLES];
    gl digest[4] = {0, 0, 0, 0};
    std::vector<uint64_t> serialize() const;  // the build artefact (little-endian u64 words)
};

// Builds (or returns the process-wide cached) definition for a circuit shape.  Pure host work: symbolic run of the AIR
// templates, constant columns, their LDE + Poseidon Merkle cap, digest.
std::shared_ptr<const CircuitDef> circuit_def_get(uint32_t kind, uint32_t n_max, const std::string& chain_id, uint64_t skip_max);
// Parses an artefact; returns nullptr (and sets the error string) when malformed or when its digest does not match.
std::shared_ptr<const CircuitDef> circuit_def_parse(const uint64_t* words, size_t n_words);

// Host evaluation of a table's DAG on one row pair (tests: the compiled templates and the data must agree).
// rows: local / next main cells, local constants, periodic values; out: value of every constraint root, in order
void circuit_def_eval_constraints(const TableDef& t, const gl* local, const gl* next, const gl* consts, const gl* periodic,
                                  std::vector<gl>& out);

}  // namespace tmx
