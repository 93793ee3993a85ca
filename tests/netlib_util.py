"""Builds oracle providers / GPU problems from MPS text through relp_b200.frontend."""
from fractions import Fraction

from oracle import relp_oracle as ro
from relp_b200 import frontend


def provider_from_mps(text):
    lp = frontend.canonicalize(frontend.parse_mps(text))
    variables = [ro.Variable(c, u) for c, u in zip(lp.costs, lp.upper)]
    md = ro.MatrixData(lp.constraint_columns, lp.b, lp.ranges, *lp.counts, variables)
    return lp, md


def scaled_from_provider(provider):
    m, n = provider.nr_rows(), provider.nr_columns()
    cols = [provider.column(j) for j in range(n)]
    costs = [provider.cost_value(j) for j in range(n)]
    rhs = provider.right_hand_side()
    pivots = provider.pivot_element_indices() if provider.has_partial_initial_basis else None
    return frontend.prescale(m, n, cols, costs, rhs, pivots, provider.has_full_initial_basis)
