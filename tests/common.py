"""Shared test helpers: oracle provider <-> integer problem, traces."""
from fractions import Fraction as F

from oracle import relp_oracle as ro


def provider_from_problem(prob):
    """IntegerProblem -> oracle ExplicitProvider (same columns, costs, rhs, pivots)."""
    cols = [prob.column(j) for j in range(prob.n)]
    return ro.ExplicitProvider(prob.m, cols, [int(c) for c in prob.cost], [int(b) for b in prob.rhs],
                               prob.pivots, prob.full_initial_basis)


def problem_from_provider(provider):
    """oracle provider with integer data -> IntegerProblem"""
    from relp_b200 import IntegerProblem
    m, n = provider.nr_rows(), provider.nr_columns()
    cols = []
    for j in range(n):
        col = []
        for i, v in provider.column(j):
            assert F(v).denominator == 1
            col.append((i, int(v)))
        cols.append(col)
    cost = [provider.cost_value(j) for j in range(n)]
    rhs = provider.right_hand_side()
    assert all(F(c).denominator == 1 for c in cost) and all(F(b).denominator == 1 for b in rhs)
    pivots = provider.pivot_element_indices() if provider.has_partial_initial_basis else None
    return IntegerProblem.from_columns(m, cols, [int(c) for c in cost], [int(b) for b in rhs], pivots,
                                       provider.has_full_initial_basis)


def oracle_trace(provider, rule, limit=None):
    """Runs the oracle; returns (result, [(phase, entering, row_in_original_space, leaving)])."""
    trace = ro.Trace(limit)
    try:
        res = ro.solve_relaxation(provider, rule, trace)
    except ro.PivotLimit:
        res = None
    removed = sorted(res.rows_removed) if res is not None else []
    keep = [i for i in range(provider.nr_rows()) if i not in set(removed)]
    out = []
    for phase, q, p, leaving, _obj in trace.pivots:
        row = keep[p] if (phase == 2 and removed) else p
        out.append((phase, q, row, leaving))
    return res, out
