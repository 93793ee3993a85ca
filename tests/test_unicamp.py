"""The reference's unicamp integration tests (tests/unicamp/test.rs, harness tests/unicamp/mod.rs:48-87) through the
restated pipeline: MPS reader -> GeneralForm -> presolve (which may solve the problem outright) -> standardize ->
MatrixData -> exact simplex -> solution reconstruction.  Expected objectives and variable values are the constants
the reference's tests hold (GLPK-checked there); the ignored reference tests are not reproduced.
CPU leg: the oracle solves; GPU leg (marked): the CUDA engine solves the same MatrixData."""
import os
from fractions import Fraction as F

import pytest

from relp_b200 import frontend, presolve
from tests.netlib_util import scaled_from_provider
from oracle import relp_oracle as ro

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "unicamp")

# (objective, values or None, mode): tests/unicamp/test.rs:8-138
EXPECTED = {
    "model_data_1": (F(123, 38), [("COL01", F(5, 2)), ("COL02", 0), ("COL03", 0), ("COL04", F(9, 14)),
                                  ("COL05", F(1, 2)), ("COL06", 4), ("COL07", 0), ("COL08", F(5, 19))], "probably"),
    "model_data_3_1": (F(70), [("SUP1", F(200, 3)), ("SUP2", F(100, 3)), ("SUP3", 100)], "exact"),
    "model_data_3_2": (F(180), [("SUP1", 25), ("SUP2", 75)], "exact"),
    "model_data_3_3": (F(245), [("SUP1", 100), ("SUP2", 150)], "exact"),
    "model_data_3_4": (F(2250), None, "objective"),
    "model_data_4": (F(7), [("COL01", 1), ("COL02", 2), ("COL03", 2)], "exact"),
    "model_data_6": (F(28), [(f"X{a}{b}", v) for a, row in enumerate([[0, 1, 1, 0, 0, 0, 0, 0], [1, 0, 0, 2, 0, 0, 0, 0],
                                                                       [1, 0, 0, 3, 0, 0, 0, 0]])
                             for b, v in enumerate(row)], "probably"),
}


def solve(name, solver):
    text = open(os.path.join(DIR, name + ".mps")).read()
    mps = frontend.parse_mps(text)
    try:
        lp = frontend.canonicalize(mps)
    except presolve.FiniteOptimum as e:              # tests/unicamp/mod.rs:63-72
        return frontend.Solution(e.objective, e.values)
    variables = [ro.Variable(c, u) for c, u in zip(lp.costs, lp.upper)]
    md = ro.MatrixData(lp.constraint_columns, lp.b, lp.ranges, *lp.counts, variables)
    bfs, objective = solver(md)
    return frontend.recover(lp, bfs, objective)


def check(name, sol):
    want_obj, want_values, mode = EXPECTED[name]
    assert sol.objective_value == want_obj
    if mode == "objective":
        return
    expected = frontend.Solution(want_obj, [(k, F(v)) for k, v in want_values])
    assert [k for k, _ in sol.solution_values] == [k for k, _ in want_values]
    if mode == "exact":
        assert sol.solution_values == expected.solution_values
    else:
        assert expected.is_probably_equal_to(sol, 0.5)


def cpu_solver(md):
    from oracle import fast_oracle as fo
    ref = fo.solve_provider(md, "steepest_edge")
    assert ref.status == "optimal"
    return ref.bfs, ref.objective


def gpu_solver(md):
    import relp_b200
    g = relp_b200.solve_relaxation(scaled_from_provider(md).problem, rule="steepest_edge")
    assert g.status == "optimal"
    return g.bfs, g.objective


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_unicamp_with_cpu_oracle(name):
    check(name, solve(name, cpu_solver))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_unicamp_with_gpu_engine(name):
    check(name, solve(name, gpu_solver))


def _nazareth():
    """tests/burkardt/test.rs:155-167: the presolved, standardized problem is unbounded"""
    text = open(os.path.join(os.path.dirname(DIR), "nazareth.mps")).read()
    lp = frontend.canonicalize(frontend.parse_mps(text))
    variables = [ro.Variable(c, u) for c, u in zip(lp.costs, lp.upper)]
    return ro.MatrixData(lp.constraint_columns, lp.b, lp.ranges, *lp.counts, variables)


def test_nazareth_is_unbounded_with_cpu_oracle():
    from oracle import fast_oracle as fo
    assert fo.solve_provider(_nazareth(), "steepest_edge").status == "unbounded"


@pytest.mark.gpu
def test_nazareth_is_unbounded_with_gpu_engine():
    import relp_b200
    from oracle import fast_oracle as fo
    md = _nazareth()
    ref = fo.solve_provider(md, "steepest_edge")
    g = relp_b200.solve_relaxation(scaled_from_provider(md).problem, rule="steepest_edge")
    assert g.status == ref.status == "unbounded" and g.trace == ref.trace


def _cook():
    """tests/cook/test.rs:17-39 (Cook et al., small example): optimum -143/2"""
    text = open(os.path.join(os.path.dirname(DIR), "cook_small_example.mps")).read()
    lp = frontend.canonicalize(frontend.parse_mps(text))
    variables = [ro.Variable(c, u) for c, u in zip(lp.costs, lp.upper)]
    return lp, ro.MatrixData(lp.constraint_columns, lp.b, lp.ranges, *lp.counts, variables)


def test_cook_small_example_with_cpu_oracle():
    lp, md = _cook()
    assert frontend.recover(lp, *cpu_solver(md)).objective_value == F(-143, 2)


@pytest.mark.gpu
def test_cook_small_example_with_gpu_engine():
    lp, md = _cook()
    assert frontend.recover(lp, *gpu_solver(md)).objective_value == F(-143, 2)
