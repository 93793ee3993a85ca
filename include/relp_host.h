/* relp_host.h -- C ABI of the host driver (librelp_gpu.so, relp_b200/csrc/host/).
 *
 * The host driver is the C++ mirror of relp's solver loops above the device engine of relp_gpu.h:
 * `SolveRelaxation::solve_relaxation` (algorithm/mod.rs:17-36, two_phase/mod.rs:25-109),
 * `phase_one::primal` (phase_one.rs:123-179), `remove_artificial_basis_variables`
 * (phase_one.rs:232-278) and `phase_two::primal` (phase_two.rs:22-58).  It takes a materialised
 * integer `MatrixProvider` and returns the exact result plus the pivot trace.
 */
#ifndef RELP_HOST_H
#define RELP_HOST_H

#include <stdint.h>
#include "relp_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A materialised MatrixProvider (matrix_provider/mod.rs:37-134) with integer data. */
typedef struct rh_problem {
    int32_t m, n;                 /* nr_rows(), nr_columns() */
    const int64_t* colptr;        /* n+1 */
    const int32_t* rowidx;        /* nnz, ascending within a column */
    const int64_t* vals;          /* nnz: column(j) */
    const int64_t* cost;          /* n: cost_value(j) */
    const int64_t* rhs;           /* m: right_hand_side(), >= 0 */
    int32_t n_pivots;             /* PartialInitialBasis::pivot_element_indices (phase_one.rs:66-79); */
    const int32_t* pivot_rows;    /*   -1 = trait not implemented => Fully artificial start         */
    const int32_t* pivot_cols;
    int32_t full_initial_basis;   /* FullInitialBasis (phase_one.rs:101-110): skip phase one */
    /* weights of a prescaled rational problem (rg_set_weights); all NULL for integer problems */
    const int64_t* colfac;        /* n */
    const int64_t* artfac;        /* m */
    const int64_t* colw;          /* n */
    const int64_t* artcost;       /* m */
    /* optional dense int8 block for provider columns [0, n_dense) (rg_load_dense_i8) */
    int32_t n_dense;
    int32_t reserved0;
    const int8_t* dense;          /* column-major n_dense x m */
} rh_problem;

enum { RH_OPTIMAL = 0, RH_UNBOUNDED = 1, RH_INFEASIBLE = 2 };

/* Trace entry.  Column indices are in the reference's index space of the phase: phase one counts the
 * artificial columns first (kind/artificial/partially.rs:52-80).  phase: 1, 2, or 0 for the
 * zero-level pivots of remove_artificial_basis_variables.  `row` is the row of the ORIGINAL provider
 * (the reference renumbers rows after deleting redundant ones; see `rows_removed`). */
typedef struct rh_trace_entry {
    int32_t phase, entering, row, leaving;
} rh_trace_entry;

typedef struct rh_options {
    int32_t device;
    int32_t initial_limbs;        /* 0 = default */
    int32_t rule;                 /* RG_RULE_*; the reference hard-codes RG_RULE_STEEPEST_EDGE */
    int32_t fused;                /* 1: rg_iterate (one host sync per pivot); 0: trait-shaped calls */
    int64_t max_pivots;           /* 0 = unlimited */
    int32_t profile;              /* 1: CUDA events around K1, 2: around every phase (rg_set_profile) */
    int32_t rank;                 /* row-shard rank / world (world <= 1: single GPU) */
    int32_t world;
    int32_t dense_carry;          /* see rg_options */
    const void* nccl_unique_id;   /* world > 1: see rg_options */
} rh_options;

typedef struct rh_result rh_result;   /* opaque; owns its buffers */

int rh_solve_relaxation(const rh_problem* problem, const rh_options* options, rh_result** out);
void rh_result_free(rh_result* r);
const char* rh_result_error(const rh_result* r);

int32_t rh_result_status(const rh_result* r);            /* RH_* */
int64_t rh_result_pivots(const rh_result* r);
int64_t rh_result_trace_len(const rh_result* r);
const rh_trace_entry* rh_result_trace(const rh_result* r);
int32_t rh_result_limbs(const rh_result* r);             /* 64-bit words per exported number */
/* optimal objective = -(minus_objective numerator) / denominator, over the integer (prescaled) data */
const uint64_t* rh_result_minus_objective(const rh_result* r);
const uint64_t* rh_result_denominator(const rh_result* r);
/* current_bfs (carry/mod.rs:636-645): basis column per row (provider index; negative = inert
 * artificial of a redundant row) and b numerators (m * limbs words) */
const int32_t* rh_result_basis(const rh_result* r);
const uint64_t* rh_result_b(const rh_result* r);
int32_t rh_result_nr_artificial(const rh_result* r);
int32_t rh_result_rows_removed_len(const rh_result* r);
const int32_t* rh_result_rows_removed(const rh_result* r);   /* Rank::Deficient rows (phase_one.rs:213-219) */
void rh_result_stats(const rh_result* r, rg_stats* out);
double rh_result_seconds(const rh_result* r);            /* wall time of the loops (excl. upload) */
double rh_result_device_ms(const rh_result* r);          /* same region, CUDA events on the engine's stream */
double rh_result_seconds_total(const rh_result* r);      /* create -> result exported */

#ifdef __cplusplus
}
#endif
#endif
