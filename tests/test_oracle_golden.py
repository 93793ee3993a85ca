"""Pins the CPU oracle against the reference's own golden fixtures.

Each test names the reference test it restates (paths under /root/reference).
"""
from fractions import Fraction as F

import pytest

from oracle import relp_oracle as ro


def fr(*xs):
    return [F(x) if not isinstance(x, tuple) else F(*x) for x in xs]


# ---- src/tests/problem_2.rs --------------------------------------------------------------------
def problem_2():
    cols = ro.columns_from_rows([[3, 2, 1, 0, 0], [5, 1, 1, 1, 0], [2, 5, 1, 0, 1]], 5)
    variables = [ro.Variable(1) for _ in range(5)]
    return ro.MatrixData(cols, [1, 3, 4], [], 3, 0, 0, 0, variables)


def rows_dense(carry):
    m = carry.m
    return [[carry.rows[i].get(k, F(0)) for k in range(m)] for i in range(m)]


def test_problem_2_artificial_tableau():
    # src/tests/problem_2.rs:117-137 (artificial_tableau_form)
    t = ro.Tableau.new_partially(problem_2())
    assert t.im.minus_objective == -8
    assert t.im.minus_pi == fr(-1, -1, -1)
    assert t.im.b == fr(1, 3, 4)
    assert t.im.basis_indices == [0, 1, 2]
    assert t.column_to_row == [0, 1, 2]
    assert rows_dense(t.im) == [fr(1, 0, 0), fr(0, 1, 0), fr(0, 0, 1)]


def test_problem_2_pipeline_first_profitable():
    # src/tests/problem_2.rs:29-67 (conversion_pipeline)
    provider = problem_2()
    art = ro.Tableau.new_partially(provider)
    res = ro.phase_one_primal(art, ro.FirstProfitable)
    assert res[0] == "feasible" and res[1] == []
    t = ro.Tableau.from_artificial(res[3], res[2], res[4], provider)
    # tableau_form, problem_2.rs:139-174
    assert t.im.minus_objective == F(-9, 2)
    assert t.im.minus_pi == fr((5, 2), -1, -1)
    assert t.im.b == fr((1, 2), (5, 2), (3, 2))
    assert t.im.basis_indices == [1, 3, 4]
    assert rows_dense(t.im) == [fr((1, 2), 0, 0), fr((-1, 2), 1, 0), fr((-5, 2), 0, 1)]
    assert t.basis_columns == {1, 3, 4}
    out = ro.phase_two_primal(t, ro.FirstProfitable)
    assert out == ("optimal", [(1, F(1, 2)), (3, F(5, 2)), (4, F(3, 2))])


def test_two_phase_simplex_and_solve_matrix():
    # src/algorithm/two_phase/test.rs:18-44 (simplex, solve_matrix)
    r = ro.solve_relaxation(problem_2())
    assert r.status == "optimal"
    assert r.bfs == [(1, F(1, 2)), (3, F(5, 2)), (4, F(3, 2))]
    assert r.objective == F(9, 2)
    for rule in ro.PIVOT_RULES:
        r = ro.solve_relaxation(problem_2(), rule)
        assert r.objective == F(9, 2)


def tableau_mod_tableau(provider):
    # tableau/mod.rs:462-489
    carry = ro.Carry(-6, fr(1, -1, -1), fr(1, 2, 3), [2, 3, 4],
                     [{0: F(1)}, {0: F(-1), 1: F(1)}, {0: F(-1), 2: F(1)}])
    return ro.Tableau(provider, carry, {2, 3, 4}, None)


def test_tableau_cost_and_relative_cost():
    # tableau/mod.rs:491-520 (cost, relative_cost)
    provider = problem_2()
    art = ro.Tableau.new_partially(provider)
    assert art.objective_function_value() == 8
    assert art.relative_cost(0) == 0
    assert art.relative_cost(art.nr_artificial_variables() + 0) == -10
    t = tableau_mod_tableau(provider)
    assert t.objective_function_value() == 6
    assert [t.relative_cost(j) for j in range(3)] == fr(-3, -3, 0)


def test_tableau_generate_column():
    # tableau/mod.rs:522-543
    provider = problem_2()
    art = ro.Tableau.new_partially(provider)
    j = art.nr_artificial_variables()
    assert art.generate_column(j) == {0: F(3), 1: F(5), 2: F(2)}
    assert art.relative_cost(j) == -10
    t = tableau_mod_tableau(provider)
    assert t.generate_column(0) == {0: F(3), 1: F(2), 2: F(-1)}
    assert t.relative_cost(0) == -3


def test_tableau_bring_into_basis():
    # tableau/mod.rs:545-566
    provider = problem_2()
    art = ro.Tableau.new_partially(provider)
    column = art.nr_artificial_variables()
    data = art.generate_column(column)
    row = art.select_primal_pivot_row(data)
    cost = art.relative_cost(column)
    art.bring_into_basis(column, row, data, cost)
    assert art.is_in_basis(column) and not art.is_in_basis(0)
    assert art.objective_function_value() == F(14, 3)
    t = tableau_mod_tableau(provider)
    data = t.generate_column(1)
    row = t.select_primal_pivot_row(data)
    t.bring_into_basis(1, row, data, t.relative_cost(1))
    assert t.is_in_basis(1)
    assert t.objective_function_value() == F(9, 2)


def test_tableau_create_tableau():
    # tableau/mod.rs:568-603 (bfs_tableau + create_tableau)
    provider = problem_2()
    m = 3
    carry = ro.Carry(0, fr(1, 1, 1), fr(1, 2, 3), [m + 2, m + 3, m + 4],
                     [{0: F(1)}, {0: F(-1), 1: F(1)}, {0: F(-1), 2: F(1)}])
    t = ro.Tableau(provider, carry, {m + 2, m + 3, m + 4}, None)
    assert ro.FirstProfitable(t).select_primal_pivot_column(t) is None


def test_pivot_rule_fixtures():
    # strategy/pivot_rule.rs:314-344 (find_profitable_column, find_pivot_row)
    provider = problem_2()
    art = ro.Tableau.new_partially(provider)
    sel = ro.FirstProfitable(art).select_primal_pivot_column(art)
    assert sel[0] == 3
    carry = ro.Carry(F(-9, 2), fr((5, 2), -1, -1), fr((1, 2), (5, 2), (3, 2)), [1, 3, 4],
                     [{0: F(1, 2)}, {0: F(-1, 2), 1: F(1)}, {0: F(-5, 2), 2: F(1)}])
    t = ro.Tableau(provider, carry, {1, 3, 4}, None)
    assert ro.FirstProfitable(t).select_primal_pivot_column(t) is None

    def col(*xs):
        return {i: F(x) for i, x in enumerate(xs) if x}
    assert art.select_primal_pivot_row(col(3, 5, 2)) == 0
    assert art.select_primal_pivot_row(col(2, 1, 5)) == 0
    assert t.select_primal_pivot_row(col(3, 2, -1)) == 0
    assert t.select_primal_pivot_row(col(2, -1, 3)) == 0


# ---- src/tests/problem_1.rs --------------------------------------------------------------------
def problem_1():
    # create_matrix_data_data + matrix_data_form, problem_1.rs:292-348
    cols = ro.columns_from_rows([[0, -1, 1], [1, 0, 1]], 3)
    variables = [ro.Variable(1, 4), ro.Variable(4, 2), ro.Variable(9)]
    return ro.MatrixData(cols, [6, 10], [], 1, 0, 0, 1, variables)


def test_problem_1_pipeline():
    # src/tests/problem_1.rs:36-108 from the MatrixData stage on
    provider = problem_1()
    assert provider.nr_rows() == 4 and provider.nr_columns() == 6
    art = ro.Tableau.new_partially(provider)
    # artificial_tableau_form, problem_1.rs:350-376
    assert art.im.minus_objective == -16
    assert art.im.minus_pi == fr(-1, -1, 0, 0)
    assert art.im.b == fr(6, 10, 4, 2)
    assert art.im.basis_indices == [0, 1, 2 + 4, 2 + 5]
    assert art.column_to_row == [0, 1]
    res = ro.phase_one_primal(art, ro.FirstProfitable)
    assert res[0] == "feasible" and res[1] == []
    t = ro.Tableau.from_artificial(res[3], res[2], res[4], provider)
    # tableau_form, problem_1.rs:378-404
    assert t.im.minus_objective == -58
    assert t.im.minus_pi == fr(4, -13, 12, 0)
    assert t.im.b == fr(6, 0, 4, 2)
    assert t.im.basis_indices == [2, 1, 0, 5]
    assert rows_dense(t.im) == [fr(0, 1, -1, 0), fr(-1, 1, -1, 0), fr(0, 0, 1, 0), fr(1, -1, 1, 1)]
    out = ro.phase_two_primal(t, ro.FirstProfitable)
    assert out == ("optimal", [(0, F(4)), (2, F(6)), (5, F(2))])
    # objective 54 = reduced objective 58 + fixed cost -4 (problem_1.rs:101-105,287)
    assert t.objective_function_value() - 4 == 54


def test_problem_1_all_rules():
    for rule in ro.PIVOT_RULES:
        r = ro.solve_relaxation(problem_1(), rule)
        assert r.status == "optimal" and r.objective == 58
        assert r.bfs == [(0, F(4)), (2, F(6)), (5, F(2))]


# ---- src/algorithm/two_phase/test.rs -----------------------------------------------------------
def test_solve_relaxation_1():
    # two_phase/test.rs:46-94
    cols = ro.columns_from_rows([[1, 0], [1, 1]], 2)
    data = ro.MatrixData(cols, fr((3, 2), (5, 2)), [], 0, 0, 2, 0,
                         [ro.Variable(-2), ro.Variable(-1)])
    r = ro.solve_relaxation(data)
    assert r.bfs == [(0, F(3, 2)), (1, F(1))]


def _two_var(rows, b, counts):
    cols = ro.columns_from_rows(rows, 2)
    variables = [ro.Variable(-2, F(3, 4)), ro.Variable(-1)]
    return ro.MatrixData(cols, b, [], *counts, variables)


def test_redundant_row():
    # two_phase/test.rs:96-134
    r = ro.solve_relaxation(_two_var([[1, 1], [1, 1], [1, 1]], [1, 1, 1], (3, 0, 0, 0)))
    assert r.status == "optimal"
    assert r.bfs == [(0, F(3, 4)), (1, F(1, 4))]
    assert r.tableau.nr_columns() == 3
    assert len(r.rows_removed) == 2


def test_empty_row_at_eq():
    # two_phase/test.rs:136-173
    r = ro.solve_relaxation(_two_var([[1, 1], [0, 0]], [1, 0], (2, 0, 0, 0)))
    assert r.bfs == [(0, F(3, 4)), (1, F(1, 4))]
    assert r.tableau.nr_columns() == 3


def test_empty_row_at_ineq():
    # two_phase/test.rs:175-212
    r = ro.solve_relaxation(_two_var([[1, 1], [0, 0]], [1, 1], (1, 0, 1, 0)))
    assert r.bfs == [(0, F(3, 4)), (1, F(1, 4)), (2, F(1))]
    assert r.tableau.nr_columns() == 4


# ---- examples ----------------------------------------------------------------------------------
def test_max_flow_example():
    # examples/max_flow.rs:261-283
    adj = ro.adjacency_from_rows([[0, 0, 0, 0], [2, 0, 0, 0], [1, 1, 0, 0], [0, 1, 2, 0]])
    problem = ro.MaxFlowPrimal(adj, 0, 3)
    r = ro.solve_relaxation(problem)
    assert r.status == "optimal"
    dense = [F(0)] * problem.nr_columns()
    for j, v in r.bfs:
        dense[j] = v
    assert dense == fr(2, 1, 1, 1, 2, 0, 0, 0, 0, 0)


def test_shortest_path_example():
    # examples/shortest_path.rs:150-167
    adj = ro.adjacency_from_rows([[0, 0, 0, 0], [1, 0, 0, 0], [2, 2, 0, 0], [0, 3, 1, 0]])
    problem = ro.ShortestPathPrimal(adj, 0, 3)
    r = ro.solve_relaxation(problem)
    dense = [F(0)] * problem.nr_columns()
    for j, v in r.bfs:
        dense[j] = v
    assert dense == fr(0, 1, 0, 0, 1)


def test_steepest_edge_recurrence_matches_recompute():
    # pivot_rule.rs:290 debug_assert_eq!(gamma, initial_gamma(j, tableau))
    class Checked(ro.SteepestDescentAlongObjective):
        def __init__(self, t):
            super().__init__(t, check=True)
    for prob in (problem_1(), problem_2()):
        r = ro.solve_relaxation(prob, Checked)
        assert r.status == "optimal"


def test_from_basis_invert_fixtures():
    """BasisInverseRows::invert fixtures of the reference (carry/basis_inverse_rows.rs:292-323): identity columns
    give the identity; [TwoSlack((0,1),(1,1)), Slack((1,1))] gives rows e_0 and (-1, 1)."""
    p = ro.ExplicitProvider(2, [[(0, 1)], [(1, 1)]], [0, 0], [1, 1])
    c = ro.Carry.from_basis([0, 1], p)
    assert rows_dense(c) == [fr(1, 0), fr(0, 1)]
    p = ro.ExplicitProvider(2, [[(0, 1), (1, 1)], [(1, 1)]], [0, 0], [3, 5])
    c = ro.Carry.from_basis([0, 1], p)
    assert rows_dense(c) == [fr(1, 0), fr(-1, 1)]
    assert c.b == fr(3, 2)
