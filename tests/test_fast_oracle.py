"""Pins the C++ fast oracle (oracle/fast_oracle.cpp) to the Python oracle, which is pinned to the
reference's golden fixtures: identical status, trace, objective and solution."""
import random
from fractions import Fraction as F

import numpy as np
import pytest

from oracle import fast_oracle as fo
from oracle import relp_oracle as ro
from tests.common import oracle_trace, provider_from_problem
from tests.test_oracle_golden import problem_1, problem_2, _two_var


def compare(provider, rules=("first_profitable", "first_profitable_with_memory", "dantzig", "steepest_edge")):
    for rule in rules:
        ores, otrace = oracle_trace(provider, rule)
        f = fo.solve_provider(provider, rule)
        assert f.status == ores.status, rule
        assert f.trace == otrace, rule
        if ores.status == "optimal":
            assert f.objective == ores.objective, rule
            assert f.bfs == ores.bfs, rule
        assert sorted(f.rows_removed) == sorted(ores.rows_removed)


def test_rational_arithmetic():
    lib = fo.load()
    rng = random.Random(5)
    for _ in range(300):
        an, ad, bn, bd = (rng.randint(1, 10 ** 9) for _ in range(4))
        an *= rng.choice([-1, 1]); bn *= rng.choice([-1, 1])
        pa, pb = rng.randint(1, 9), rng.randint(1, 9)
        a, b = F(an, ad) ** pa, F(bn, bd) ** pb
        for op, want in ((0, a + b), (1, a - b), (2, a * b), (3, a / b), (4, F((a > b) - (a < b)))):
            got = fo.parse_rational(lib.fo_arith(op, an, ad, pa, bn, bd, pb))
            assert got == want, (op, an, ad, pa, bn, bd, pb)


def test_golden_fixtures():
    compare(problem_1())
    compare(problem_2())
    compare(_two_var([[1, 1], [1, 1], [1, 1]], [1, 1, 1], (3, 0, 0, 0)))   # redundant_row
    compare(_two_var([[1, 1], [0, 0]], [1, 0], (2, 0, 0, 0)))               # empty_row_at_eq
    compare(_two_var([[1, 1], [0, 0]], [1, 1], (1, 0, 1, 0)))               # empty_row_at_ineq
    adj = ro.adjacency_from_rows([[0, 0, 0, 0], [2, 0, 0, 0], [1, 1, 0, 0], [0, 1, 2, 0]])
    compare(ro.MaxFlowPrimal(adj, 0, 3))
    adj = ro.adjacency_from_rows([[0, 0, 0, 0], [1, 0, 0, 0], [2, 2, 0, 0], [0, 3, 1, 0]])
    compare(ro.ShortestPathPrimal(adj, 0, 3))


@pytest.mark.parametrize("seed", range(10))
def test_random_lps(seed):
    from tests.test_gpu_parity import random_matrix_data
    rng = np.random.default_rng(500 + seed)
    nv = int(rng.integers(3, 10))
    counts = tuple(int(rng.integers(0, 4)) for _ in range(4))
    if sum(counts) == 0:
        counts = (1, 1, 1, 0)
    compare(random_matrix_data(rng, nv, counts))


def test_synthetic_and_maxflow():
    from relp_b200.generators import bounded_lp, max_flow
    for prob in (bounded_lp(60, 120, k_bounding=20, nnz_per_col=4, seed=2),
                 bounded_lp(30, 40, k_bounding=10, dense=True, seed=1),
                 max_flow(24, 3, 3, 9)):
        provider = provider_from_problem(prob)
        for rule in ("steepest_edge", "dantzig"):
            ores, otrace = oracle_trace(provider, rule)
            f = fo.solve_problem(prob, rule)
            assert f.status == ores.status and f.trace == otrace
            assert f.objective == ores.objective and f.bfs == ores.bfs
