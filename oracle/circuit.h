/*
 * oracle/circuit.h -- the circuit as data, read from the build artefact.  TEST INFRASTRUCTURE ONLY (see oracle/gl.h).
 */
#ifndef TMX_ORACLE_CIRCUIT_H
#define TMX_ORACLE_CIRCUIT_H

#include "oracle.h"
#include "../include/tmx_trace.h"

#define CIRCUIT_MAGIC 0x32434D54ULL /* "TMC2" */
enum { SYM_CONST = 0, SYM_COL = 1, SYM_ADD = 2, SYM_SUB = 3, SYM_MUL = 4 };
enum { SRC_LOCAL = 0, SRC_NEXT = 1, SRC_CONST = 2, SRC_PERIODIC = 3 };

typedef struct {
    uint32_t op, a, b, deg; /* COL: a = source, b = column */
    uint64_t val;           /* CONST */
} sym_node_t;

/* prog: {0, node} constraint | {1, tag, m, len, v..} bus.one | {2, (tag, m, len, v..) x 2} bus.two */
typedef struct {
    uint32_t present, log_n, n_main, n_const, n_per, period, n_helpers, n_constraints;
    gl_t *periodic;  /* [n_per][period] */
    gl_t *constants; /* [n_const][n] */
    gl_t *const_cap;
    size_t cap_len;
    sym_node_t *nodes;
    size_t n_nodes;
    uint64_t *prog;
    size_t prog_len;
    uint8_t *bus_mask; /* nodes the bus items depend on */
} table_def_t;

typedef struct {
    uint32_t kind, n_max;
    uint64_t skip_max;
    uint8_t chain_id[64];
    size_t chain_len;
    uint64_t params[6]; /* rate_bits, cap_height, pow_bits, n_queries, arity_bits, final_poly_bits */
    table_def_t t[TMX_N_TABLES];
    gl_t digest[4];
} circuit_def_t;

int circuit_parse(const uint64_t *words, size_t n_words, circuit_def_t *out);
void circuit_free(circuit_def_t *c);
/* values of all (or the masked) nodes for one row; v has n_nodes entries */
void circuit_eval_b(const table_def_t *t, const uint8_t *mask, const gl_t *local, const gl_t *next, const gl_t *k, const gl_t *per,
                    gl_t *v);
void circuit_eval_e(const table_def_t *t, const gl2_t *local, const gl2_t *next, const gl2_t *k, const gl2_t *per, gl2_t *v);

#endif
