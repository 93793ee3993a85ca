"""Solution back-substitution: the basic feasible solution of the solved MatrixData is
mapped back to the ORIGINAL variables by name and compared with the solutions the reference's own tests assert
(tests/burkardt/test.rs:53-191): AFIRO's 32 values through `Solution::is_probably_equal_to(.., 0.1)` exactly like
the reference does (:112), maros and testprob with `assert_eq!` semantics (:142-150, :184-191) -- plus an exact
feasibility check against the raw MPS rows, which the reference does not do.
CPU leg: the C++ oracle solves; GPU leg (marked): the CUDA engine solves."""
import os
from fractions import Fraction as F

import pytest

from relp_b200 import frontend
from tests.netlib_util import provider_from_mps, scaled_from_provider

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

AFIRO = [("X01", F(80)), ("X02", F(51, 2)), ("X03", F(109, 2)), ("X04", F(424, 5)), ("X06", F(255, 14)),
         ("X07", 0), ("X08", 0), ("X09", 0), ("X10", 0), ("X11", 0), ("X12", 0), ("X13", 0), ("X14", F(255, 14)),
         ("X15", 0), ("X16", F(999)), ("X22", F(500)), ("X23", F(11898, 25)), ("X24", F(602, 25)), ("X25", 0),
         ("X26", F(215)), ("X28", 0), ("X29", 0), ("X30", 0), ("X31", 0), ("X32", 0), ("X33", 0), ("X34", 0),
         ("X35", 0), ("X36", F(11898, 35)), ("X37", F(11898, 35)), ("X38", 0), ("X39", 0)]
EXPECTED = {
    "afiro.mps": (F(-406659, 875), AFIRO, "probably"),
    "AFIRO.SIF": (F(-406659, 875), AFIRO, "probably"),
    "maros.mps": (F(385, 3), [("VOL1", F(10, 3)), ("VOL2", F(40, 3)), ("VOL3", F(20)), ("VOL4", F(0))], "exact"),
    "testprob.mps": (F(54), [("X1", F(4)), ("X2", F(-1)), ("X3", F(6))], "exact"),
}


def feasible_in_original(mps, values):
    """every row and bound of the raw MPS holds exactly; returns the objective row's value"""
    x = dict(values)
    act = {r: F(0) for r in mps["rows"]}
    obj = F(0)
    for col, entries in mps["columns"].items():
        for r, v in entries.items():
            if r == mps["objective"]:
                obj += v * x[col]
            elif r in act:
                act[r] += v * x[col]
    for r in mps["rows"]:
        lo, hi = mps["interval"][r]
        assert lo is None or act[r] >= lo, r
        assert hi is None or act[r] <= hi, r
    for col in mps["col_order"]:
        lo, hi = mps["bounds"].get(col, [F(0), frontend.INF])
        assert lo is frontend.INF or x[col] >= lo, col
        assert hi is frontend.INF or x[col] <= hi, col
    return obj


def check_solution(name, bfs, objective, lp, mps):
    want_obj, want_values, mode = EXPECTED[name]
    sol = frontend.recover(lp, bfs, objective)
    expected = frontend.Solution(want_obj, [(k, F(v)) for k, v in want_values])
    assert sol.objective_value == want_obj
    assert [k for k, _ in sol.solution_values] == [k for k, _ in want_values]
    if mode == "exact":
        assert sol.solution_values == expected.solution_values
    else:
        assert expected.is_probably_equal_to(sol, 0.1)          # tests/burkardt/test.rs:112
    assert feasible_in_original(mps, sol.solution_values) == want_obj


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_recover_with_cpu_oracle(name):
    from oracle import fast_oracle as fo
    text = open(os.path.join(GOLD, name)).read()
    mps = frontend.parse_mps(text)
    lp, md = provider_from_mps(text)
    ref = fo.solve_provider(md, "steepest_edge")
    assert ref.status == "optimal"
    check_solution(name, ref.bfs, ref.objective, lp, mps)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_recover_with_gpu_engine(name):
    import relp_b200
    text = open(os.path.join(GOLD, name)).read()
    mps = frontend.parse_mps(text)
    lp, md = provider_from_mps(text)
    sp = scaled_from_provider(md)
    g = relp_b200.solve_relaxation(sp.problem, rule="steepest_edge")
    assert g.status == "optimal"
    check_solution(name, g.bfs, g.objective, lp, mps)


def test_solution_probably_equal_semantics():
    a = frontend.Solution(F(1), [(f"x{i}", F(i)) for i in range(12)])
    b = frontend.Solution(F(1), [(f"x{i}", F(i if i < 2 else -i)) for i in range(12)])
    assert a.is_probably_equal_to(b, 0.1) and not a.is_probably_equal_to(b, 0.5)
    assert not a.is_probably_equal_to(frontend.Solution(F(2), a.solution_values), 0.0)
