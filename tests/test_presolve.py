"""`GeneralForm::presolve` as restated in relp_b200/presolve.py + general_form.py, against the reference's own
presolve tests:

* tests/golden/presolve_changes.json -- the 29 cases of `presolve/test/changes.rs` (each: a `GeneralForm` literal and
  the exact `Changes` / `Err(..)` that `compute_presolve_changes` must return), extracted mechanically by
  scripts/gen_presolve_fixtures.py;
* `presolve/test/with_application.rs` -- a six-variable LP that the presolve solves completely (hand-ported);
* `presolve/test/per_rule.rs` -- the fixed-variable rule in isolation (hand-ported);
* the TESTPROB pipeline of src/tests/problem_1.rs: presolve + standardize must give the expected standardized
  general form and `MatrixData` arguments (:262-372) exactly."""
import json
import os
from fractions import Fraction as F

import pytest

from relp_b200 import mps, presolve
from relp_b200.general_form import GeneralForm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "presolve_changes.json")))["cases"]


def dec(x):
    if isinstance(x, str) and "/" in x and x.replace("/", "").replace("-", "").isdigit():
        p, q = x.split("/")
        return F(int(p), int(q))
    if isinstance(x, list):
        return [dec(v) for v in x]
    return x


def build(case):
    rows = dec(case["rows"])
    n = case["ncols"]
    cols = [[(i, rows[i][j]) for i in range(len(rows)) if rows[i][j] != 0] for j in range(n)]
    variables = []
    for v in case["variables"]:
        var = mps.Variable("continuous", dec(v["cost"]))
        var.lower_bound, var.upper_bound = dec(v["lower"]), dec(v["upper"])
        variables.append(var)
    types = [tuple(t) if isinstance(t, list) else t for t in dec(case["constraint_types"])]
    data = mps.GeneralFormData(case["objective"], cols, len(rows), types, dec(case["b"]), variables,
                               [f"x{j}" for j in range(n)], [f"r{i}" for i in range(len(rows))])
    data.fixed_cost = dec(case["fixed_cost"])
    return GeneralForm(data)


def as_tuple(t):
    return tuple(as_tuple(v) for v in t) if isinstance(t, list) else t


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_compute_presolve_changes_matches_the_reference_fixture(case):
    g = build(case)
    expect = case["expect"]
    if expect["kind"] == "err":
        with pytest.raises(presolve.Infeasible if expect["value"] == "infeasible" else presolve.Unbounded):
            presolve.compute_presolve_changes(g)
        return
    want = {k: dec(v) for k, v in expect["value"]}
    got = presolve.compute_presolve_changes(g)
    assert got["fixed_cost"] == want["fixed_cost"]
    assert got["constraints_marked_removed"] == want["constraints_marked_removed"]
    assert got["b"] == {i: v for i, v in want["b"]}
    assert got["constraints"] == {i: (tuple(t) if isinstance(t, list) else t) for i, t in want["constraints"]}
    assert got["bounds"] == {tuple(k): v for k, v in want["bounds"]}
    want_removed = [(j, as_tuple(sol)) for j, sol in want["removed_variables"]]
    got_removed = [(j, as_tuple([sol[0], sol[1], [list(t) for t in sol[2]]]) if sol[0] == "function" else sol)
                   for j, sol in got["removed_variables"]]
    assert got_removed == want_removed


def _v(cost, lo, up):
    v = mps.Variable("continuous", F(cost))
    v.lower_bound, v.upper_bound = lo, up
    return v


def test_presolve_solves_the_with_application_problem():
    """presolve/test/with_application.rs:25-127"""
    rows = [[2, 0, 0, 0, 0, 0], [3, 5, 0, 0, 0, 0], [7, 11, 13, 0, 0, 0], [17, 19, 23, 0, 29, 31]]
    cols = [[(i, F(rows[i][j])) for i in range(4) if rows[i][j]] for j in range(6)]
    variables = [_v(211, None, None), _v(223, (F(103) - F(101) / 2 * 3) / 5, None), _v(227, None, None),
                 _v(-229, None, F(131)), _v(233, F(-30736, 65 * 29), F(123)), _v(0, F(5), None)]
    names = ["XONE", "XTWO", "XTHREE", "XFOUR", "XFIVE", "XSIX"]
    data = mps.GeneralFormData("minimize", cols, 4, ["E", "L", "G", "E"], [F(101), F(103), F(107), F(109)],
                               variables, names, ["r0", "r1", "r2", "r3"])
    data.fixed_cost = F(1)
    with pytest.raises(presolve.FiniteOptimum) as e:
        GeneralForm(data).presolve()
    assert e.value.objective == (F(1) + F(211 * 101, 2) + F(223 * -97, 10) + F(227 * -699, 65) + F(-229 * 131)
                                 + F(233 * -30736, 1885))
    assert e.value.values == [("XONE", F(101, 2)), ("XTWO", (F(103) - F(101) / 2 * 3) / 5),
                              ("XTHREE", (F(-3601, 5) + F(29 * 30736, 1885)) / 23), ("XFOUR", F(131)),
                              ("XFIVE", F(-30736, 65 * 29)), ("XSIX", F(5))]


def test_fixed_variable_rule_in_isolation():
    """presolve/test/per_rule.rs:14-44 (feasible) and :46-68 (infeasible)"""
    def index(types):
        data = mps.GeneralFormData("minimize", [[(0, F(1)), (1, F(2))]], 2, types, [F(1), F(1)], [_v(1, F(1), F(1))],
                                   ["X"], ["a", "b"])
        data.fixed_cost = F(7)
        return presolve.Index(GeneralForm(data))
    ix = index(["E", "G"])
    ix.presolve_fixed_variable(0)
    assert ix.count_constraint == [0, 0] and ix.count_variable == [0]
    assert ix.constraints_marked_removed == [0, 1]
    assert ix.removed_variables == [(0, ("solved", F(1)))]
    assert ix.b == {0: F(0), 1: F(-1)} and ix.fixed_cost == F(1)
    with pytest.raises(presolve.Infeasible):
        index(["E", "E"]).presolve_fixed_variable(0)


def test_problem_1_presolved_and_standardized_is_the_reference_fixture():
    """src/tests/problem_1.rs:262-372: rows MYEQN (b = 6) and LIM2 (b = 10), LIM1 removed by domain propagation,
    shifts / bounds / fixed cost as expected, MatrixData::new(.., 1, 0, 0, 1, ..)"""
    from tests.test_mps_reader import PROBLEM_1
    g = GeneralForm(mps.parse(PROBLEM_1).to_general_form())
    g.presolve()
    counts = g.standardize()
    assert counts == [1, 0, 0, 1]
    assert g.columns == [[(1, F(1))], [(0, F(-1))], [(0, F(1)), (1, F(1))]]
    assert g.constraint_types == ["E", "G"] and g.b == [F(6), F(10)]
    assert g.fixed_cost == F(-4)
    got = [(v.variable_type, v.cost, v.lower_bound, v.upper_bound, v.shift, v.flipped) for v in g.variables]
    assert got == [("continuous", F(1), F(0), F(4), F(0), False), ("integer", F(4), F(0), F(2), F(1), False),
                   ("continuous", F(9), F(0), None, F(0), False)]
    cols, b, ranges, ne, nr, nu, nl, variables = g.derive_matrix_data(counts)
    assert (ne, nr, nu, nl) == (1, 0, 0, 1) and ranges == []
    cost, values = g.compute_full_solution_with_reduced_solution({0: F(4), 2: F(6)})
    assert cost == F(54) and values == [("XONE", F(4)), ("YTWO", F(-1)), ("ZTHREE", F(6))]
