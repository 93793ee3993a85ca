/* relp_gpu.h -- C ABI of the B200-native exact simplex engine (librelp_gpu.so).
 *
 * This is the drop-in boundary for relp's hot path: the functions below are what a Rust
 * `GpuCarry: InverseMaintainer` / `GpuPivotRule: PivotRule` pair (bindgen over this header), the
 * C++ host driver (`relp_b200/csrc/host/`) and the Python ctypes tests all call.  Plain pointers
 * and sizes only.  Each entry point cites the reference interface it replaces (paths relative to
 * the reference's `src/algorithm/two_phase/`).
 *
 * Representation.  The (m+1) x (m+1) carry matrix of `Carry<F, BI>`
 * (tableau/inverse_maintenance/carry/mod.rs:46-66)
 *
 *        | -obj | -pi        |        row 0
 *        |  b   | B^-1       |        rows 1..m
 *
 * lives on the device as integer numerators over ONE common positive denominator D = |det B|
 * (Edmonds / Bareiss form): value = numerator / D.  Numerators are two's complement integers of
 * L x 64-bit limbs (L in {1,2,4,8,16}), stored limb-planar.  The provider's columns, right-hand
 * side and costs must be integers (the host prescales rational rows).
 *
 * Column ids.  Provider column j has id j (0 <= j < n).  The virtual artificial column of
 * phase one with reference index a (0 <= a < n_a, tableau/kind/artificial/partially.rs:52-80) has id
 * a - n_a (negative), which preserves the reference's ordering "artificials first" for Bland's rule.
 *
 * All calls are synchronous at return, single-owner and non-reentrant per context (the reference is
 * single-threaded: tableau/mod.rs:25-39).  Return value: RG_OK or a negative error code; the
 * simplex outcome of a step is reported through the `status` output where one exists.
 */
#ifndef RELP_GPU_H
#define RELP_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rg_context rg_context;

/* error codes (function return values) */
enum {
    RG_OK = 0,
    RG_ERR_CUDA = -1,        /* a CUDA runtime call failed (see rg_last_error) */
    RG_ERR_ARG = -2,         /* invalid argument */
    RG_ERR_OVERFLOW = -3,    /* a number needs more than 16 limbs: no CPU fallback exists */
    RG_ERR_STATE = -4,       /* call sequence violated (e.g. pivot before a column was generated) */
    RG_ERR_NCCL = -5
};

/* simplex step status */
enum {
    RG_STEP_PIVOTED = 0,     /* a basis change happened */
    RG_STEP_OPTIMAL = 1,     /* no entering column: `None` of select_primal_pivot_column */
    RG_STEP_UNBOUNDED = 2    /* no leaving row: `None` of select_primal_pivot_row */
};

/* pivot rules: strategy/pivot_rule.rs */
enum {
    RG_RULE_FIRST_PROFITABLE = 0,             /* :86-109  */
    RG_RULE_FIRST_PROFITABLE_WITH_MEMORY = 1, /* :113-150 */
    RG_RULE_DANTZIG = 2,                      /* SteepestDescentAlongVariable :153-187, lowest j on ties */
    RG_RULE_STEEPEST_EDGE = 3                 /* SteepestDescentAlongObjective :190-305, highest j on ties */
};

typedef struct rg_options {
    int32_t device;          /* CUDA device ordinal */
    int32_t initial_limbs;   /* a width of the ladder (1, 2, 4, 8, 10, 12, 14, 16); 0 = default (2) */
    int32_t rank;            /* row-shard rank (0 when single GPU) */
    int32_t world;           /* number of row shards (1 when single GPU) */
    int32_t dense_carry;     /* 1: always run the dense carry kernels; 0 (default): active-column mode -- carry
                                columns that still equal D*e_k are neither stored nor touched until a pivot in
                                their own row makes them general */
    int32_t reserved;
    const void* nccl_unique_id;  /* world > 1: the 128-byte ncclUniqueId shared by all ranks (rg_nccl_unique_id
                                    on rank 0, broadcast by the host, e.g. torch.distributed) */
} rg_options;

/* The limb-width ladder of the engine: 1, 2, 4, 8, 10, 12, 14, 16 limbs (powers of two up to 512 bits, then
 * steps of 128 bits).  Per-width statistics are indexed by position in this ladder (rg_stats.limb_widths). */
#define RG_NWIDTHS 8

typedef struct rg_stats {
    int64_t pivots;              /* basis changes performed */
    int64_t promotions;          /* limb-width promotions (K9) */
    int32_t limbs;               /* current limb count L */
    int32_t max_bits;            /* largest |numerator| bit length currently in the carry */
    int32_t denominator_bits;    /* bit length of D */
    int32_t reserved;            /* active-column mode: number of non-trivial carry columns; 0 in dense mode */
    int64_t kernel_launches;     /* kernels launched by this context so far */
    int64_t demotions;           /* limb-width demotions (numerators shrank by more than a width step) */
    int32_t limb_widths[RG_NWIDTHS];        /* the ladder: limb count of each per-width slot below */
    int64_t pivots_at_limbs[RG_NWIDTHS];    /* pivots performed at each width */
    /* profiling (rg_set_profile): CUDA-event time of the rank-1 update kernel (K1) per limb width */
    int64_t k1_launches_at_limbs[RG_NWIDTHS];
    double k1_ms_at_limbs[RG_NWIDTHS];
    double timer_ms;             /* rg_timer_start .. rg_timer_stop on the engine's stream */
    /* profiling: CUDA-event time per iteration phase: 0 pivot column + ratio test + row staging,
     * 1 work vector, 2 pivot scalars, 3 K1 update, 4 bookkeeping + steepest-edge update,
     * 5 pricing + column selection */
    double phase_ms[8];
    /* profiling: algorithmic work of the timed K1 launches per limb width -- bytes (every entry of the active
     * part of the carry read and written once) and IMAD.WIDE multiply-adds (two low products per entry) */
    double k1_bytes_at_limbs[RG_NWIDTHS];
    double k1_imads_at_limbs[RG_NWIDTHS];
} rg_stats;

typedef struct rg_pivot_info {   /* BasisChangeComputationInfo, tableau/mod.rs:205-234 (indices only) */
    int32_t status;              /* RG_STEP_* */
    int32_t entering;            /* column id that entered (pivot_column_index) */
    int32_t row;                 /* pivot_row_index, 0-based constraint row */
    int32_t leaving;             /* leaving_column_index (column id; negative = artificial) */
} rg_pivot_info;

/* ---- life cycle ------------------------------------------------------------------------------ */
int rg_create(const rg_options* opts, rg_context** out);   /* `IM::create_*` allocate; Drop <-> rg_destroy */
int rg_destroy(rg_context* ctx);
const char* rg_last_error(const rg_context* ctx);
/* Row-sharded operation (one process per GPU): every rank creates a context with the same `world` and the
 * same NCCL id and then issues the SAME sequence of calls; the carry rows are block-distributed, the cost
 * row, pricing data and all scalars are replicated, and every call returns identical results on every
 * rank.  Returns the id size (128) or an error. */
int rg_nccl_unique_id(void* out, int32_t bytes);

/* Device buffers of 1 MiB and more are recycled by exact size in a per-device free list of the process (repeated
 * solves allocate the same sizes; the stream-ordered pool stalled now and then on large requests).  This returns
 * the parked buffers of `device` to the driver's pool, e.g. before another library needs the memory; live contexts
 * are unaffected.  Returns the number of bytes released (>= 0) or an RG_ERR_* code. */
int64_t rg_release_cached_memory(int32_t device);

/* The block partition used for the carry rows (count = m) and for the priced columns (count = number of dense /
 * CSC columns): rank r owns [first, first + number).  Pure host arithmetic, no context or device needed. */
int rg_shard_block(int32_t count, int32_t world, int32_t rank, int32_t* first, int32_t* number);

/* ---- problem upload: MatrixProvider (matrix_provider/mod.rs:37-134) ---------------------------- */
/* All provider columns as integer CSC (column(j), :52), row indices ascending within a column.  */
int rg_load_csc(rg_context* ctx, int32_t m, int32_t n, const int64_t* colptr,
                const int32_t* rowidx, const int64_t* vals);
/* Optional dense block: provider columns [0, nd) given as int8, column-major nd x m (implicit row indices;
 * config 5).  Their CSC ranges in rg_load_csc must be empty.  Call right after rg_load_csc. */
int rg_load_dense_i8(rg_context* ctx, int32_t nd, const int8_t* colmajor);
/* right_hand_side() (:96); must be >= 0 (GeneralForm::make_b_non_negative guarantees it). */
int rg_set_rhs(rg_context* ctx, const int64_t* b);

/* Weights of a prescaled rational problem (host prescale; all 1 and not needed
 * for integer problems).  Row i of the rational problem was multiplied by r_i; column j of the integer
 * image equals the scaled column divided by w_j (unit slack / bound columns keep +-1 entries: w_j = r_i),
 * the artificial of row i is the unit column (weight r_i).  With W = lcm of all weights:
 *   colfac[j] = W / w_j,  colw[j] = w_j,  artfac[i] = W / r_i,
 *   artcost[i] = phase-one cost numerator of the artificial of row i (a common positive multiple of 1/r_i).
 * They keep Dantzig and steepest-edge choices identical to the rational problem's.  Call after rg_load_csc
 * and before rg_init_identity_basis.  All values must be in [1, 2^31). */
int rg_set_weights(rg_context* ctx, const int64_t* colfac, const int64_t* artfac, const int64_t* colw,
                   const int64_t* artcost);

/* ---- carry constructors (tableau/inverse_maintenance/mod.rs:30-130; carry/mod.rs:374-442) ----- */
/* create_for_fully_artificial / create_for_partially_artificial / from_basis_pivots with an identity
 * basis: `basis[i]` is the column id basic in row i (negative = artificial of that row, whose phase-one
 * cost is 1).  `cost` (length n, may be NULL = all zero) is the cost vector of the phase started:
 * NULL in phase one (provider columns cost 0, kind/artificial/partially.rs:52), the real integer
 * costs when starting directly in phase two (FullInitialBasis, two_phase/mod.rs:80-109).  Requires that
 * every non-artificial basic column is a unit column e_i (a positive slack), so B^-1 = I, D = 1. */
int rg_init_identity_basis(rg_context* ctx, const int32_t* basis, const int64_t* cost);

/* from_basis / from_basis_pivots on a GENERAL basis (carry/mod.rs:444-497) = BasisInverse::invert
 * (carry/basis_inverse_rows.rs:104-129, lower_upper/decomposition/mod.rs:27-143): `basis[i]` is the provider
 * column basic in row i (m distinct columns, non-singular); `cost` the cost vector of the phase started.  The
 * inverse is built on the device by fraction-free Gauss-Jordan (the engine's own rank-1 update per non-unit
 * column), D = |det B|.  Enables warm starts.  Single GPU. */
int rg_init_basis(rg_context* ctx, const int32_t* basis, const int64_t* cost);

/* from_artificial (carry/mod.rs:499-525): install the phase-two costs and rebuild -pi = -c_B^T B^-1 and
 * -obj = -c_B b.  Rows whose artificial is still basic (rank-deficient rows, phase_one.rs:232-278) stay in
 * the carry as inert rows with cost 0, which is value-identical to deleting them
 * (from_artificial_remove_rows, carry/mod.rs:527-559,673-712). */
int rg_phase_switch(rg_context* ctx, const int64_t* cost);

/* ---- pivot rule (strategy/pivot_rule.rs:23-54) -------------------------------------------------- */
/* PivotRule::new: selects the rule; for steepest edge computes every gamma_j = 1 + |B^-1 a_j|^2
 * (initial_gamma, :299-305). */
int rg_rule_new(rg_context* ctx, int32_t rule);
/* PivotRule::select_primal_pivot_column (:32-41): status RG_STEP_OPTIMAL when none; else *q. */
int rg_select_primal_pivot_column(rg_context* ctx, int32_t* status, int32_t* q);

/* ---- tableau operations (tableau/mod.rs) -------------------------------------------------------- */
/* Tableau::generate_column (:126-130) = Carry::generate_column (carry/mod.rs:613-621): computes
 * B^-1 a_q on the device and keeps it there as the current pivot column. */
int rg_generate_column(rg_context* ctx, int32_t q);
/* Tableau::select_primal_pivot_row (:287-313) on the current pivot column: min ratio, ties to the
 * lowest leaving column id (Bland).  status RG_STEP_UNBOUNDED when none; else *row (0-based). */
int rg_select_primal_pivot_row(rg_context* ctx, int32_t* status, int32_t* row);
/* Tableau::bring_into_basis (:48-64) = Carry::change_basis (carry/mod.rs:561-604) followed by
 * PivotRule::after_basis_update (pivot_rule.rs:43-53) when `update_rule` != 0.  Promotes the limb
 * width and retries when the result would not fit.  Uses the current pivot column. */
int rg_bring_into_basis(rg_context* ctx, int32_t q, int32_t row, int32_t update_rule,
                        rg_pivot_info* info);

/* One full iteration of phase_one::primal / phase_two::primal (phase_one.rs:134-178,
 * phase_two.rs:36-57): select column, generate it, ratio test, change basis, rule update -- all
 * device-driven with a single host synchronisation.  Runs up to `max_pivots` iterations; stops early
 * on OPTIMAL / UNBOUNDED.  `trace` (may be NULL) receives one rg_pivot_info per performed pivot
 * (capacity `max_pivots`); `*n_done` the number performed; `*status` the final step status. */
int rg_iterate(rg_context* ctx, int64_t max_pivots, rg_pivot_info* trace, int64_t* n_done,
               int32_t* status);

/* remove_artificial_basis_variables (phase_one.rs:232-278) for one row that still holds an
 * artificial: finds the first non-basic provider column j (ascending) with element (B^-1 a_j)[row] > 0
 * and relative cost 0 when b_row != 0, or element != 0 when b_row == 0, and pivots there at zero
 * level.  info->status = RG_STEP_PIVOTED, or RG_STEP_OPTIMAL when no column exists (row redundant). */
int rg_remove_artificial_row(rg_context* ctx, int32_t row, rg_pivot_info* info);

/* ---- exports (InverseMaintainer getters, tableau/inverse_maintenance/mod.rs:200-264) ------------ */
/* Numbers leave as two's complement little-endian limb arrays of `*limbs` 64-bit words each, all over
 * the common denominator returned by rg_get_denominator. */
int rg_get_limbs(rg_context* ctx, int32_t* limbs);
int rg_get_denominator(rg_context* ctx, uint64_t* out /* limbs words */);
int rg_get_basis(rg_context* ctx, int32_t* basis /* m column ids: basis_column_index_for_row */);
int rg_get_b(rg_context* ctx, uint64_t* out /* m * limbs words: b() numerators */);
int rg_get_minus_objective(rg_context* ctx, uint64_t* out /* limbs words */);
int rg_get_minus_pi(rg_context* ctx, uint64_t* out /* m * limbs words */);
int rg_get_basis_inverse_row(rg_context* ctx, int32_t row, uint64_t* out /* m * limbs words */);
/* current pivot column numerators (generate_column), limbs+2 words per entry, m entries */
int rg_get_pivot_column(rg_context* ctx, uint64_t* out);
/* relative cost numerators of every provider column (Tableau::relative_cost, tableau/mod.rs:106-112),
 * limbs+2 words per entry; basic columns report 0 */
int rg_get_relative_costs(rg_context* ctx, uint64_t* out);
/* steepest-edge weights gamma_j * D^2 for every provider column, 2*limbs+6 words per entry (0 for
 * basic columns) */
int rg_get_gamma(rg_context* ctx, uint64_t* out);
/* InverseMaintainer::generate_element (tableau/inverse_maintenance/mod.rs:200-215 ->
 * BasisInverse::generate_element, carry/basis_inverse_rows.rs:179-195): the single entry
 * (B^-1 a_j)[row], one numerator of limbs+2 words over the current denominator. */
int rg_get_element(rg_context* ctx, int32_t row, int32_t j, uint64_t* out);
/* BasisChangeComputationInfo (tableau/mod.rs:205-234) of the last basis change -- what a stock host-side
 * `PivotRule::after_basis_update` receives (the indices are in rg_pivot_info):
 *   column  = column_before_change, m entries of limbs+2 words, over `denominator_before`
 *   work    = work_vector = column^T B_old^-1, m entries of 2*limbs+5 words, over denominator_before^2
 *             (kept only when the device steepest-edge update ran; otherwise pass NULL)
 *   row     = basis_inverse_row = row p of the new B^-1, m entries of limbs words, over the current denominator
 *   denominator_before = limbs words.
 * Any pointer may be NULL.  Valid until the next call that generates a column or changes the basis. */
int rg_get_basis_change_info(rg_context* ctx, uint64_t* column, uint64_t* work, uint64_t* row,
                             uint64_t* denominator_before);
int rg_get_stats(rg_context* ctx, rg_stats* out);
/* measurement hooks (no reference counterpart): on = 1 puts CUDA events around every K1 launch (rg_stats
 * k1_ms_at_limbs), on = 2 additionally around every phase of an iteration (rg_stats phase_ms); and a
 * stream timer */
int rg_set_profile(rg_context* ctx, int32_t on);
int rg_timer_start(rg_context* ctx);
int rg_timer_stop(rg_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RELP_GPU_H */
