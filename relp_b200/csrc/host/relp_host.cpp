// Host driver: C++ mirror of relp's solver loops above the device engine (include/relp_gpu.h).
//
// Names follow the reference (paths relative to src/algorithm/two_phase/):
//   GpuCarry            <-> Carry<F, BI> as InverseMaintainer  (tableau/inverse_maintenance/mod.rs:30-264)
//   Tableau             <-> Tableau<IM, K>                     (tableau/mod.rs:25-39)
//   PivotRule           <-> trait PivotRule                    (strategy/pivot_rule.rs:23-54)
//   phase_one::primal, remove_artificial_basis_variables, phase_two::primal, solve_relaxation
// The carry, the pricing data and the steepest-edge weights live on the device; this layer only
// sequences the calls and keeps the index bookkeeping the reference keeps in `Kind`.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/relp_host.h"

namespace relp {

struct Error {
    int code;
    std::string what;
};

static void check(rg_context* ctx, int rc, const char* where) {
    if (rc != RG_OK) throw Error{rc, std::string(where) + ": " + (ctx ? rg_last_error(ctx) : "")};
}

// MatrixProvider (matrix_provider/mod.rs:37-134), materialised.
struct MatrixProvider {
    const rh_problem* p;
    int nr_rows() const { return p->m; }
    int nr_columns() const { return p->n; }
    bool partial_initial_basis() const { return p->n_pivots >= 0; }
    bool full_initial_basis() const { return p->full_initial_basis != 0; }
};

// The device-resident InverseMaintainer.
class GpuCarry {
public:
    rg_context* ctx = nullptr;
    int m = 0, n = 0;

    explicit GpuCarry(const MatrixProvider& mp, const rh_options& o) : m(mp.nr_rows()), n(mp.nr_columns()) {
        rg_options opts{};
        opts.device = o.device;
        opts.initial_limbs = o.initial_limbs;
        opts.world = o.world > 1 ? o.world : 1;
        opts.rank = o.world > 1 ? o.rank : 0;
        opts.nccl_unique_id = o.nccl_unique_id;
        opts.dense_carry = o.dense_carry;
        {
            int rc = rg_create(&opts, &ctx);
            if (rc != RG_OK) {          // the destructor will not run: release the half-built context here
                std::string why = ctx ? rg_last_error(ctx) : "";
                if (ctx) rg_destroy(ctx);
                ctx = nullptr;
                throw Error{rc, "rg_create: " + why};
            }
        }
        check(ctx, rg_load_csc(ctx, m, n, mp.p->colptr, mp.p->rowidx, mp.p->vals), "rg_load_csc");
        if (mp.p->n_dense > 0) check(ctx, rg_load_dense_i8(ctx, mp.p->n_dense, mp.p->dense), "rg_load_dense_i8");
        check(ctx, rg_set_rhs(ctx, mp.p->rhs), "rg_set_rhs");
        if (mp.p->colfac)
            check(ctx, rg_set_weights(ctx, mp.p->colfac, mp.p->artfac, mp.p->colw, mp.p->artcost), "rg_set_weights");
    }
    ~GpuCarry() { rg_destroy(ctx); }
    GpuCarry(const GpuCarry&) = delete;

    // create_for_fully_artificial / create_for_partially_artificial (carry/mod.rs:374-442)
    void create_for_artificial(const std::vector<int>& basis_ids) {
        check(ctx, rg_init_identity_basis(ctx, basis_ids.data(), nullptr), "rg_init_identity_basis");
    }
    // from_basis_pivots (carry/mod.rs:480-497), identity bases
    void from_basis_pivots(const std::vector<int>& basis_ids, const int64_t* cost) {
        check(ctx, rg_init_identity_basis(ctx, basis_ids.data(), cost), "rg_init_identity_basis");
    }
    // from_artificial / from_artificial_remove_rows (carry/mod.rs:499-559)
    void from_artificial(const int64_t* cost) { check(ctx, rg_phase_switch(ctx, cost), "rg_phase_switch"); }
};

// PivotRule (strategy/pivot_rule.rs:23-54): a handle on the device-side rule state.
class PivotRule {
public:
    GpuCarry& im;
    int rule;
    PivotRule(GpuCarry& carry, int rule_) : im(carry), rule(rule_) {   // PivotRule::new
        check(im.ctx, rg_rule_new(im.ctx, rule), "rg_rule_new");
    }
    // select_primal_pivot_column: false <=> None
    bool select_primal_pivot_column(int* q) {
        int32_t st = 0;
        check(im.ctx, rg_select_primal_pivot_column(im.ctx, &st, q), "rg_select_primal_pivot_column");
        return st == RG_STEP_PIVOTED;
    }
};

// Tableau<IM, K>: K is Artificial (phase one) while `nr_artificial > 0 || in_phase_one`.
class Tableau {
public:
    GpuCarry& im;
    int nr_artificial = 0;                 // Artificial::nr_artificial_variables
    std::vector<int> column_to_row;        // Partially::column_to_row (partially.rs:17-21)
    std::vector<int> basis;                // basis_indices in engine ids (negative = artificial)

    explicit Tableau(GpuCarry& carry) : im(carry) {}

    void generate_column(int q) { check(im.ctx, rg_generate_column(im.ctx, q), "rg_generate_column"); }
    bool select_primal_pivot_row(int* row) {
        int32_t st = 0;
        check(im.ctx, rg_select_primal_pivot_row(im.ctx, &st, row), "rg_select_primal_pivot_row");
        return st == RG_STEP_PIVOTED;
    }
    rg_pivot_info bring_into_basis(int q, int row, bool update_rule) {
        rg_pivot_info info{};
        check(im.ctx, rg_bring_into_basis(im.ctx, q, row, update_rule ? 1 : 0, &info), "rg_bring_into_basis");
        basis[row] = q;
        return info;
    }
    bool has_artificial_in_basis() const {
        for (int j : basis) if (j < 0) return true;
        return false;
    }
    // reference column index of an engine id in the current phase
    int ref_index(int id, bool phase_one) const { return phase_one ? id + nr_artificial : id; }
};

struct Outcome {
    int status = RH_OPTIMAL;
    std::vector<rh_trace_entry> trace;
    std::vector<int> rows_removed;
    int nr_artificial = 0;
    int64_t pivots = 0;
};

namespace phase_two { enum Result { FiniteOptimum, Unbounded, Limit }; }

// The shared loop body of phase_one::primal (phase_one.rs:134-178) and phase_two::primal
// (phase_two.rs:36-57).
static phase_two::Result simplex_loop(Tableau& t, int rule_id, bool phase_one, const rh_options& o,
                                      Outcome& out) {
    PivotRule rule(t.im, rule_id);
    const int phase = phase_one ? 1 : 2;
    auto budget = [&]() -> int64_t {
        return o.max_pivots > 0 ? o.max_pivots - out.pivots : (int64_t)1 << 62;
    };
    if (o.fused) {
        std::vector<rg_pivot_info> buf(4096);
        for (;;) {
            int64_t want = std::min<int64_t>((int64_t)buf.size(), budget());
            if (want <= 0) return phase_two::Limit;
            int64_t done = 0;
            int32_t st = 0;
            check(t.im.ctx, rg_iterate(t.im.ctx, want, buf.data(), &done, &st), "rg_iterate");
            for (int64_t k = 0; k < done; ++k) {
                t.basis[buf[k].row] = buf[k].entering;
                out.trace.push_back({phase, t.ref_index(buf[k].entering, phase_one), buf[k].row,
                                     t.ref_index(buf[k].leaving, phase_one)});
            }
            out.pivots += done;
            if (st == RG_STEP_OPTIMAL) return phase_two::FiniteOptimum;
            if (st == RG_STEP_UNBOUNDED) return phase_two::Unbounded;
        }
    }
    for (;;) {
        if (budget() <= 0) return phase_two::Limit;
        int q = -1, row = -1;
        if (!rule.select_primal_pivot_column(&q)) return phase_two::FiniteOptimum;
        t.generate_column(q);
        if (!t.select_primal_pivot_row(&row)) return phase_two::Unbounded;
        rg_pivot_info info = t.bring_into_basis(q, row, true);   // + rule.after_basis_update
        out.trace.push_back({phase, t.ref_index(q, phase_one), row, t.ref_index(info.leaving, phase_one)});
        out.pivots++;
    }
}

// remove_artificial_basis_variables (phase_one.rs:232-278)
static std::vector<int> remove_artificial_basis_variables(Tableau& t, Outcome& out) {
    std::vector<int> rows_to_remove;
    for (int row = 0; row < t.im.m; ++row) {
        if (t.basis[row] >= 0) continue;
        rg_pivot_info info{};
        check(t.im.ctx, rg_remove_artificial_row(t.im.ctx, row, &info), "rg_remove_artificial_row");
        if (info.status == RG_STEP_PIVOTED) {
            t.basis[row] = info.entering;
            out.trace.push_back({0, t.ref_index(info.entering, true), row, t.ref_index(info.leaving, true)});
            out.pivots++;
        } else {
            rows_to_remove.push_back(row);
        }
    }
    return rows_to_remove;
}

static bool objective_is_zero(GpuCarry& im) {
    int32_t L = 0;
    check(im.ctx, rg_get_limbs(im.ctx, &L), "rg_get_limbs");
    std::vector<uint64_t> v(L);
    check(im.ctx, rg_get_minus_objective(im.ctx, v.data()), "rg_get_minus_objective");
    for (uint64_t x : v) if (x) return false;
    return true;
}

// SolveRelaxation::solve_relaxation (two_phase/mod.rs:25-109)
static void solve_relaxation(const MatrixProvider& mp, const rh_options& o, GpuCarry& im, Outcome& out) {
    const int m = mp.nr_rows();
    Tableau t(im);
    t.basis.assign(m, 0);
    if (mp.full_initial_basis()) {
        // two_phase/mod.rs:80-109
        for (int k = 0; k < mp.p->n_pivots; ++k) t.basis[mp.p->pivot_rows[k]] = mp.p->pivot_cols[k];
        im.from_basis_pivots(t.basis, mp.p->cost);
    } else {
        // Tableau::<_, Fully>::new (fully.rs:82-97) / Tableau::<_, Partially>::new (partially.rs:125-205)
        std::vector<int> real_col(m, -1);
        if (mp.partial_initial_basis())
            for (int k = 0; k < mp.p->n_pivots; ++k) real_col[mp.p->pivot_rows[k]] = mp.p->pivot_cols[k];
        for (int i = 0; i < m; ++i) if (real_col[i] < 0) t.column_to_row.push_back(i);
        t.nr_artificial = (int)t.column_to_row.size();
        int a = 0;
        for (int i = 0; i < m; ++i) t.basis[i] = real_col[i] >= 0 ? real_col[i] : (a++) - t.nr_artificial;
        im.create_for_artificial(t.basis);
        out.nr_artificial = t.nr_artificial;

        // phase_one::primal (phase_one.rs:123-179)
        phase_two::Result r = simplex_loop(t, o.rule, true, o, out);
        if (r == phase_two::Limit) { out.status = -1; return; }
        if (r == phase_two::Unbounded) throw Error{RG_ERR_STATE, "Artificial cost can not be unbounded."};
        if (!objective_is_zero(im)) { out.status = RH_INFEASIBLE; return; }
        if (t.has_artificial_in_basis()) out.rows_removed = remove_artificial_basis_variables(t, out);
        // Tableau::from_artificial[_removing_rows] (non_artificial.rs:151-226)
        im.from_artificial(mp.p->cost);
        t.nr_artificial = 0;
    }
    phase_two::Result r = simplex_loop(t, o.rule, false, o, out);
    out.status = r == phase_two::FiniteOptimum ? RH_OPTIMAL : (r == phase_two::Unbounded ? RH_UNBOUNDED : -1);
}

}  // namespace relp

struct rh_result {
    std::string err;
    relp::Outcome out;
    int32_t limbs = 0;
    std::vector<uint64_t> minus_obj, denom, b;
    std::vector<int32_t> basis;
    rg_stats stats{};
    double seconds = 0, seconds_total = 0;
};

extern "C" int rh_solve_relaxation(const rh_problem* problem, const rh_options* options, rh_result** outp) {
    if (!problem || !options || !outp) return RG_ERR_ARG;
    rh_result* res = new rh_result();
    *outp = res;
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    try {
        relp::MatrixProvider mp{problem};
        relp::GpuCarry im(mp, *options);
        if (options->profile) rg_set_profile(im.ctx, options->profile);
        auto t1 = clk::now();
        rg_timer_start(im.ctx);
        relp::solve_relaxation(mp, *options, im, res->out);
        rg_timer_stop(im.ctx);
        auto t2 = clk::now();
        res->seconds = std::chrono::duration<double>(t2 - t1).count();
        relp::check(im.ctx, rg_get_limbs(im.ctx, &res->limbs), "rg_get_limbs");
        res->minus_obj.resize(res->limbs);
        res->denom.resize(res->limbs);
        res->b.resize((size_t)res->limbs * problem->m);
        res->basis.resize(problem->m);
        relp::check(im.ctx, rg_get_minus_objective(im.ctx, res->minus_obj.data()), "rg_get_minus_objective");
        relp::check(im.ctx, rg_get_denominator(im.ctx, res->denom.data()), "rg_get_denominator");
        relp::check(im.ctx, rg_get_b(im.ctx, res->b.data()), "rg_get_b");
        relp::check(im.ctx, rg_get_basis(im.ctx, res->basis.data()), "rg_get_basis");
        rg_get_stats(im.ctx, &res->stats);
        res->seconds_total = std::chrono::duration<double>(clk::now() - t0).count();
        if (getenv("RG_HOSTPROF"))
            fprintf(stderr, "[hostprof] create+upload %.3f s, loops %.3f s, export %.3f s\n",
                    std::chrono::duration<double>(t1 - t0).count(), res->seconds,
                    std::chrono::duration<double>(clk::now() - t2).count());
    } catch (const relp::Error& e) {
        res->err = e.what;
        return e.code;
    } catch (const std::exception& e) {     // nothing may propagate across the C boundary
        res->err = std::string("host driver: ") + e.what();
        return RG_ERR_STATE;
    } catch (...) {
        res->err = "host driver: unknown exception";
        return RG_ERR_STATE;
    }
    return RG_OK;
}

extern "C" void rh_result_free(rh_result* r) { delete r; }
extern "C" const char* rh_result_error(const rh_result* r) { return r ? r->err.c_str() : ""; }
extern "C" int32_t rh_result_status(const rh_result* r) { return r->out.status; }
extern "C" int64_t rh_result_pivots(const rh_result* r) { return r->out.pivots; }
extern "C" int64_t rh_result_trace_len(const rh_result* r) { return (int64_t)r->out.trace.size(); }
extern "C" const rh_trace_entry* rh_result_trace(const rh_result* r) { return r->out.trace.data(); }
extern "C" int32_t rh_result_limbs(const rh_result* r) { return r->limbs; }
extern "C" const uint64_t* rh_result_minus_objective(const rh_result* r) { return r->minus_obj.data(); }
extern "C" const uint64_t* rh_result_denominator(const rh_result* r) { return r->denom.data(); }
extern "C" const int32_t* rh_result_basis(const rh_result* r) { return r->basis.data(); }
extern "C" const uint64_t* rh_result_b(const rh_result* r) { return r->b.data(); }
extern "C" int32_t rh_result_nr_artificial(const rh_result* r) { return r->out.nr_artificial; }
extern "C" int32_t rh_result_rows_removed_len(const rh_result* r) { return (int32_t)r->out.rows_removed.size(); }
extern "C" const int32_t* rh_result_rows_removed(const rh_result* r) { return r->out.rows_removed.data(); }
extern "C" void rh_result_stats(const rh_result* r, rg_stats* out) { *out = r->stats; }
extern "C" double rh_result_seconds(const rh_result* r) { return r->seconds; }
extern "C" double rh_result_device_ms(const rh_result* r) { return r->stats.timer_ms; }
extern "C" double rh_result_seconds_total(const rh_result* r) { return r->seconds_total; }
