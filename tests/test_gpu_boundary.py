"""Entry-by-entry carry parity through the trait-shaped C ABI (include/relp_gpu.h): after EVERY pivot the
device carry (-obj, -pi, b, every row of B^-1), the relative costs, the steepest-edge weights, the generated
pivot column, single elements and the exported BasisChangeComputationInfo are compared with the oracle's
`Carry` / `Tableau` / `PivotRule` state -- and, on problem_1 / problem_2, with the staged golden carries the
reference's own tests hold (src/tests/problem_1.rs:237-292,380-422, src/tests/problem_2.rs:117-174).
"""
from fractions import Fraction as F

import numpy as np
import pytest

from oracle import relp_oracle as ro
from tests.common import problem_from_provider

pytestmark = pytest.mark.gpu


def _dense_row(d, m):
    return [d.get(k, F(0)) for k in range(m)]


class Lockstep:
    """Drives the oracle and the GPU engine through the same calls and compares the complete state."""

    def __init__(self, provider, rule, initial_limbs=0, dense_carry=False):
        from tests.gpu_engine import Engine
        self.provider = provider
        self.rule_name = rule
        self.rule_cls = ro.PIVOT_RULES[rule]
        self.prob = problem_from_provider(provider)
        self.eng = Engine(self.prob, initial_limbs=initial_limbs, dense_carry=dense_carry)
        self.m, self.n = self.prob.m, self.prob.n
        self.trace = []
        self.checked = 0

    def close(self):
        self.eng.close()

    # ---- comparisons ---------------------------------------------------------------------------------
    def compare_state(self, t, rule, na):
        e, m, n = self.eng, self.m, self.n
        im = t.im
        assert e.minus_objective() == im.minus_objective
        assert e.minus_pi() == im.minus_pi
        assert e.b() == im.b
        ids = e.basis()
        assert [i + na for i in ids] == im.basis_indices
        for r in range(m):
            assert e.basis_inverse_row(r) == _dense_row(im.rows[r], m), f"B^-1 row {r}"
        rc = e.relative_costs()
        for j in range(n):
            if not t.is_in_basis(na + j):
                assert rc[j] == t.relative_cost(na + j), f"relative cost of column {j}"
            else:
                assert rc[j] == 0
        if self.rule_name == "steepest_edge" and rule is not None:
            g = e.gamma()
            for j in range(n):
                want = rule.gamma[na + j]
                if want is not None:
                    assert g[j] == want, f"gamma of column {j}"
        # single elements (generate_element): a handful per state
        rng = np.random.default_rng(self.checked)
        for _ in range(4):
            r, j = int(rng.integers(0, m)), int(rng.integers(0, n))
            want = t.generate_element(r, na + j)
            assert e.element(r, j) == (want if want is not None else 0), f"element ({r},{j})"
        self.checked += 1

    def loop(self, t, na, phase):
        e = self.eng
        rule = self.rule_cls(t)
        e.rule_new(self.rule_name)
        while True:
            self.compare_state(t, rule, na)
            sel = rule.select_primal_pivot_column(t)
            q = e.select_column()
            if sel is None:
                assert q is None
                return "optimal"
            assert q is not None and q + na == sel[0]
            column = t.generate_column(sel[0])
            e.generate_column(q)
            assert e.pivot_column() == _dense_row(column, self.m)
            p = t.select_primal_pivot_row(column)
            row = e.select_row()
            if p is None:
                assert row is None
                return "unbounded"
            assert row == p
            info = t.bring_into_basis(sel[0], p, dict(column), sel[1])
            entering, prow, leaving = e.bring_into_basis(q, row, True)
            assert (entering + na, prow, leaving + na) == (sel[0], p, info["leaving_column_index"])
            # BasisChangeComputationInfo (tableau/mod.rs:205-234)
            se = self.rule_name == "steepest_edge"
            col_b, work, row_p = e.basis_change_info(want_work=se)
            assert col_b == _dense_row(info["column_before_change"], self.m)
            assert row_p == _dense_row(info["basis_inverse_row"], self.m)
            if se:
                assert work == _dense_row(info["work_vector"], self.m)
            rule.after_basis_update(info, t)
            self.trace.append((phase, sel[0], p, info["leaving_column_index"]))

    def run(self, hooks=None):
        """hooks: optional dict name -> callable(tableau) called at the staged points of the golden tests"""
        hooks = hooks or {}
        provider, e, m = self.provider, self.eng, self.m
        cost = [int(provider.cost_value(j)) for j in range(self.n)]
        if provider.has_full_initial_basis:
            pivots = provider.pivot_element_indices()
            basis = [0] * m
            for r, c in pivots:
                basis[r] = c
            e.init_identity_basis(basis, cost)
            t = ro.Tableau(provider, ro.Carry.from_basis_pivots(pivots, provider), [c for _, c in pivots], None)
        else:
            art = (ro.Tableau.new_partially(provider) if provider.has_partial_initial_basis
                   else ro.Tableau.new_fully(provider))
            na = art.nr_artificial_variables()
            ids = [j - na for j in art.im.basis_indices]
            e.init_identity_basis(ids, None)
            if "phase_one_start" in hooks:
                hooks["phase_one_start"](art)
            assert self.loop(art, na, 1) == "optimal"
            if art.objective_function_value() != 0:
                return "infeasible", None
            if "phase_one_end" in hooks:
                hooks["phase_one_end"](art)
            removed = []
            if art.has_artificial_in_basis():
                rows = [r for r, _ in art.artificial_basis_columns()]
                removed = ro.remove_artificial_basis_variables(art, None)
                for r in rows:
                    pivoted, *_ = e.remove_artificial_row(r)
                    assert pivoted == (r not in removed)
                self.compare_state(art, None, na)
            assert not removed, "rank-deficient cases are covered by tests/test_gpu_parity.py"
            t = ro.Tableau.from_artificial(art.im, na, set(art.basis_columns), provider)
            e.phase_switch(cost)
        if "phase_two_start" in hooks:
            hooks["phase_two_start"](t)
        status = self.loop(t, 0, 2)
        return status, t


@pytest.mark.parametrize("rule", ["steepest_edge", "dantzig", "first_profitable"])
@pytest.mark.parametrize("dense_carry", [False, True])
def test_problem_2_staged_carries(rule, dense_carry):
    """src/tests/problem_2.rs:117-174: the carry before phase one, after phase one, and the optimum 9/2."""
    from tests.test_oracle_golden import problem_2
    ls = Lockstep(problem_2(), rule, dense_carry=dense_carry)
    seen = {}

    def start(t):
        assert t.im.minus_objective == -8 and t.im.minus_pi == [-1, -1, -1] and t.im.b == [1, 3, 4]
        assert ls.eng.minus_objective() == -8 and ls.eng.minus_pi() == [-1, -1, -1] and ls.eng.b() == [1, 3, 4]
        seen["start"] = True

    try:
        status, t = ls.run({"phase_one_start": start})
        assert seen.get("start") and status == "optimal"
        assert t.objective_function_value() == F(9, 2)
        assert -ls.eng.minus_objective() == F(9, 2)
        assert ls.checked >= 3
    finally:
        ls.close()


@pytest.mark.parametrize("rule", ["steepest_edge", "dantzig", "first_profitable_with_memory"])
def test_problem_1_every_pivot(rule):
    """src/tests/problem_1.rs: Partially-artificial start (slack pivots), optimum 58 at x = (4, 0, 6, ..)."""
    from tests.test_oracle_golden import problem_1
    ls = Lockstep(problem_1(), rule)
    try:
        status, t = ls.run()
        assert status == "optimal"
        assert -ls.eng.minus_objective() == t.objective_function_value()
        assert ls.checked >= 3
    finally:
        ls.close()


@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("limbs", [1, 4])
def test_random_lp_every_pivot(seed, limbs):
    from tests.test_gpu_parity import random_matrix_data
    rng = np.random.default_rng(4200 + seed)
    md = random_matrix_data(rng, 6, (2, 1, 2, 1))
    ls = Lockstep(md, "steepest_edge", initial_limbs=limbs)
    try:
        status, _ = ls.run()
        assert status in ("optimal", "unbounded", "infeasible")
    finally:
        ls.close()


def test_bounded_lp_with_dense_block_every_pivot():
    """FullInitialBasis start, dense int8 block, active-column carry: every getter after every pivot."""
    from relp_b200.generators import bounded_lp
    from tests.common import provider_from_problem
    prob = bounded_lp(24, 32, k_bounding=8, dense=True, seed=11, dense_block=True)
    ls = Lockstep(provider_from_problem(prob), "steepest_edge", initial_limbs=1)
    ls.prob = prob                      # keep the dense block (problem_from_provider would rebuild CSC)
    ls.eng.close()
    from tests.gpu_engine import Engine
    ls.eng = Engine(prob, initial_limbs=1)
    try:
        status, _ = ls.run()
        assert status == "optimal" and ls.checked >= 5
    finally:
        ls.close()


def test_zero_pivot_is_rejected():
    """rg_bring_into_basis on a row whose pivot-column entry is zero must fail with RG_ERR_ARG, not corrupt
    the carry (reference: debug_assert on the pivot value, carry/mod.rs:303-304)."""
    from tests.gpu_engine import Engine
    from tests.test_oracle_golden import problem_2
    prob = problem_from_provider(problem_2())
    with Engine(prob) as e:
        t = ro.Tableau.new_fully(problem_2())
        na = t.nr_artificial_variables()
        e.init_identity_basis([j - na for j in t.im.basis_indices], None)
        e.rule_new("dantzig")
        # find a (column, row) with a zero entry
        for q in range(prob.n):
            col = t.generate_column(na + q)
            zero_rows = [r for r in range(prob.m) if r not in col]
            if zero_rows:
                e.generate_column(q)
                with pytest.raises(RuntimeError, match="pivot element is zero"):
                    e.bring_into_basis(q, zero_rows[0], False)
                break
        else:
            pytest.skip("no zero entry in this fixture")
        # the carry is untouched and still usable
        assert e.b() == t.im.b and e.minus_pi() == t.im.minus_pi


# ---- general-basis constructor: from_basis (carry/mod.rs:444-478), BasisInverse::invert ---------------------
def _state_matches(e, t, m, n, rule=None):
    im = t.im
    assert e.minus_objective() == im.minus_objective
    assert e.minus_pi() == im.minus_pi
    assert e.b() == im.b
    assert e.basis() == im.basis_indices
    for r in range(m):
        assert e.basis_inverse_row(r) == _dense_row(im.rows[r], m), f"B^-1 row {r}"
    if rule is not None:
        g = e.gamma()
        for j in range(n):
            if rule.gamma[j] is not None:
                assert g[j] == rule.gamma[j], f"gamma {j}"


def _oracle_tableau_from_basis(provider, basis):
    """oracle restatement of from_basis: basis[i] is the column basic in row i"""
    carry = ro.Carry.from_basis_pivots([(i, j) for i, j in enumerate(basis)], provider)
    return ro.Tableau(provider, carry, list(basis), None)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("shuffle", [False, True])
def test_init_basis_mid_solve_and_continue(seed, shuffle):
    """Warm start: the basis an oracle solve has reached after a few pivots is handed to rg_init_basis (rows in
    the oracle's order, or shuffled: exercises the device row permutation); the rebuilt carry equals the
    oracle's from_basis carry entry by entry, and the continued solve walks the oracle's pivots."""
    from relp_b200.generators import bounded_lp
    from tests.common import provider_from_problem
    from tests.gpu_engine import Engine
    prob = bounded_lp(36, 48, k_bounding=10, nnz_per_col=5, seed=40 + seed)
    provider = provider_from_problem(prob)
    # reach a mid-solve basis with the oracle (phase two from the slack basis)
    pivots = provider.pivot_element_indices()
    t = ro.Tableau(provider, ro.Carry.from_basis_pivots(pivots, provider), [c for _, c in pivots], None)
    rule = ro.SteepestDescentAlongObjective(t)
    for _ in range(6):
        sel = rule.select_primal_pivot_column(t)
        if sel is None:
            break
        col = t.generate_column(sel[0])
        p = t.select_primal_pivot_row(col)
        info = t.bring_into_basis(sel[0], p, col, sel[1])
        rule.after_basis_update(info, t)
    basis = list(t.im.basis_indices)
    if shuffle:
        rng = np.random.default_rng(seed)
        rng.shuffle(basis)
    t2 = _oracle_tableau_from_basis(provider, basis)
    cost = [int(provider.cost_value(j)) for j in range(prob.n)]
    with Engine(prob, initial_limbs=1) as e:
        e.init_basis(basis, cost)
        rule2 = ro.SteepestDescentAlongObjective(t2)
        e.rule_new("steepest_edge")                 # general-basis steepest-edge initialisation
        _state_matches(e, t2, prob.m, prob.n, rule2)
        # continue to the optimum in lockstep
        steps = 0
        while True:
            sel = rule2.select_primal_pivot_column(t2)
            q = e.select_column()
            if sel is None:
                assert q is None
                break
            assert q == sel[0]
            col = t2.generate_column(q)
            e.generate_column(q)
            p = t2.select_primal_pivot_row(col)
            assert e.select_row() == p
            info = t2.bring_into_basis(q, p, col, sel[1])
            e.bring_into_basis(q, p, True)
            rule2.after_basis_update(info, t2)
            steps += 1
        _state_matches(e, t2, prob.m, prob.n, rule2)


def test_init_basis_rejects_singular_basis():
    from relp_b200.generators import bounded_lp
    from tests.gpu_engine import Engine
    prob = bounded_lp(12, 16, k_bounding=4, nnz_per_col=3, seed=3)
    ns = prob.n - prob.m
    # two structural columns with identical support cannot both... use a duplicated column id instead
    basis = [ns + i for i in range(prob.m)]
    basis[1] = basis[0]
    with Engine(prob) as e:
        with pytest.raises(RuntimeError, match="distinct"):
            e.init_basis(basis, list(prob.cost))


def test_steepest_edge_init_on_general_basis_with_dense_block():
    """PivotRule::new(steepest edge) after phase-one pivots with a dense int8 block loaded (the
    row-wise tensor-core initialisation).  Rows 0..3 get artificials (their slack pivots are withheld), so phase
    one pivots before the phase-two rule is created; full trace / objective / solution against the oracle."""
    import relp_b200
    from relp_b200.generators import bounded_lp
    from tests.common import oracle_trace, provider_from_problem
    prob = bounded_lp(40, 56, k_bounding=12, dense=True, seed=21, dense_block=True, full_initial_basis=False)
    prob.pivots = [rc for rc in prob.pivots if rc[0] >= 4]
    provider = provider_from_problem(prob)
    ores, otrace = oracle_trace(provider, "steepest_edge")
    assert any(ph == 1 for ph, *_ in otrace)
    for limbs in (1, 4):
        g = relp_b200.solve_relaxation(prob, rule="steepest_edge", initial_limbs=limbs)
        assert g.status == ores.status and g.trace == otrace
        assert g.objective == ores.objective and g.bfs == ores.bfs


def test_release_cached_memory_returns_the_recycled_buffers():
    """buffers >= 1 MiB are parked by exact size when a context lets go of them; rg_release_cached_memory hands them
    back to the driver, and a solve afterwards still works (and parks them again)"""
    import ctypes as C
    import relp_b200
    from relp_b200 import _lib
    from relp_b200.generators import bounded_lp
    lib = _lib.load()
    prob = bounded_lp(1100, 640, k_bounding=24, dense=True, seed=9, dense_block=True)
    g1 = relp_b200.solve_relaxation(prob, rule="steepest_edge")
    freed = lib.rg_release_cached_memory(0)
    assert freed > 0
    assert lib.rg_release_cached_memory(0) == 0
    assert lib.rg_release_cached_memory(99) < 0
    g2 = relp_b200.solve_relaxation(prob, rule="steepest_edge")
    assert g2.trace == g1.trace and g2.objective == g1.objective
    assert lib.rg_release_cached_memory(0) > 0
