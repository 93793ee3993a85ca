"""`GeneralForm::standardize` / `derive_matrix_data` / solution reconstruction as restated in
relp_b200/general_form.py (reference: src/data/linear_program/general_form/mod.rs:262-332,506-684,800-905).

Fixture: the reference's own TESTPROB (src/tests/problem_1.rs).  Its expected `general_form_standardized()`
(:262-313) and `MatrixData` (:315-372) are taken AFTER presolve, which removes the redundant row LIM1
(x + y <= 5 with x <= 4, y <= 1); the restated pipeline skips presolve, so its output must equal that fixture
once the LIM1 row is dropped -- and must carry the same shifts, bounds and fixed cost."""
from fractions import Fraction as F

from relp_b200 import mps
from relp_b200.general_form import GeneralForm
from tests.test_mps_reader import PROBLEM_1


def _lp(rows, cols, rhs="", bounds="", ranges="", sense=""):
    text = "NAME t\n" + sense + "ROWS\n N  c\n" + rows + "COLUMNS\n" + cols
    if rhs:
        text += "RHS\n" + rhs
    if ranges:
        text += "RANGES\n" + ranges
    if bounds:
        text += "BOUNDS\n" + bounds
    return mps.parse(text + "ENDATA").to_general_form()


def test_problem_1_standardized_matches_the_reference_fixture_up_to_presolve():
    g = GeneralForm(mps.parse(PROBLEM_1).to_general_form())
    counts = g.standardize()
    assert counts == [1, 0, 1, 1]
    assert g.row_names == ["MYEQN", "LIM1", "LIM2"]
    assert g.b == [F(6), F(6), F(10)]
    assert g.fixed_cost == F(-4)                                     # problem_1.rs:312: -R64!(1 * 4)
    assert g.columns == [[(1, F(1)), (2, F(1))], [(0, F(-1)), (1, F(1))], [(0, F(1)), (2, F(1))]]
    got = [(v.variable_type, v.cost, v.lower_bound, v.upper_bound, v.shift, v.flipped) for v in g.variables]
    assert got == [("continuous", F(1), F(0), F(4), F(0), False),      # problem_1.rs:278-303
                   ("integer", F(4), F(0), F(2), F(1), False),
                   ("continuous", F(9), F(0), None, F(0), False)]
    # dropping the presolve-redundant LIM1 row gives the reference's presolved fixture (:262-276)
    keep = {0: 0, 2: 1}
    reduced = [[(keep[i], v) for i, v in col if i in keep] for col in g.columns]
    assert reduced == [[(1, F(1))], [(0, F(-1))], [(0, F(1)), (1, F(1))]]
    assert [g.b[0], g.b[2]] == [F(6), F(10)] and [g.constraint_types[0], g.constraint_types[2]] == ["E", "G"]
    cols, b, ranges, ne, nr, nu, nl, variables = g.derive_matrix_data(counts)
    assert (ne, nr, nu, nl) == (1, 0, 1, 1) and ranges == []
    assert variables == [(F(1), F(4)), (F(4), F(2)), (F(9), None)]
    # optimum of the reference's test (problem_1.rs:94-107): 54 at x = (4, -1, 6), reduced y' = y + 1 = 0
    cost, values = g.compute_full_solution_with_reduced_solution({0: F(4), 2: F(6)})
    assert cost == F(54) and values == [("XONE", F(4)), ("YTWO", F(-1)), ("ZTHREE", F(6))]


def test_free_variables_are_split_with_the_negative_halves_appended():
    gf = _lp(" L  r1\n G  r2\n", "    X         r1   1   c   2\n    Y         r1   3   r2   1\n    Z         r2   5\n",
             rhs="    R         r1   4   r2   1\n", bounds=" FR B         X\n FR B         Z\n")
    g = GeneralForm(gf)
    counts = g.standardize()
    assert counts == [0, 0, 1, 1]
    assert len(g.variables) == 5 and g.from_active_to_original == [0, 1, 2, 0, 2]
    assert g.original_variables == [("active_free", 0, 3), ("active", 1), ("active_free", 2, 4)]
    assert g.columns[3] == [(i, -v) for i, v in g.columns[0]] and g.columns[4] == [(i, -v) for i, v in g.columns[2]]
    assert [v.cost for v in g.variables] == [F(2), F(0), F(0), F(-2), F(0)]
    cost, values = g.compute_full_solution_with_reduced_solution({3: F(7), 2: F(1)})
    assert values == [("X", F(-7)), ("Y", F(0)), ("Z", F(1))] and cost == F(-14)


def test_flip_shift_negative_rhs_and_range_rows():
    gf = _lp(" L  a\n G  b\n E  e\n L  r\n",
             "    X         a    1   b   1\n    X         e    1   r   1\n    X         c    3\n"
             "    Y         a    1   r   2\n",
             rhs="    R         a   -2   b  -3\n    R         e   -4   r   10\n",
             ranges="    G         r    4\n",
             bounds=" MI B         X\n UP B         X                   -1\n LO B         Y                    2\n")
    assert gf.constraint_types == ["L", "G", "E", ("R", F(4))] and gf.b == [F(-2), F(-3), F(-4), F(10)]
    g = GeneralForm(gf)
    counts = g.standardize()
    x, y = g.variables
    # X in (-inf, -1]: flipped, x' = -x - 1 >= 0; Y >= 2: y' = y - 2
    assert (x.flipped, x.shift, x.lower_bound, x.upper_bound, x.cost) == (True, F(-1), F(0), None, F(-3))
    assert (y.flipped, y.shift, y.lower_bound, y.upper_bound) == (False, F(-2), F(0), None)
    assert g.fixed_cost == F(-3)                   # cost 3 at x = -1
    # rows after the substitution x = -1 - x', y = 2 + y':  a: -x' + y' <= -3 -> x' - y' >= 3 (a G row);
    # b: -x' >= -2 -> x' <= 2 (an L row); e: -x' = -3 -> x' = 3; r in [6, 10]: -x' + 2y' in [3, 7] (b = 7, range 4)
    assert counts == [1, 1, 1, 1]
    assert g.row_names == ["e", "r", "b", "a"]
    assert g.b == [F(3), F(7), F(2), F(3)]
    assert g.constraint_types == ["E", ("R", F(4)), "L", "G"]
    assert g.columns == [[(0, F(1)), (1, F(-1)), (2, F(1)), (3, F(1))], [(1, F(2)), (3, F(-1))]]
    cost, values = g.compute_full_solution_with_reduced_solution({0: F(3)})
    assert values == [("X", F(-4)), ("Y", F(2))] and cost == F(-12)
    # a ranged row whose upper end is negative: the interval is mirrored, b = r - b_old
    g2 = GeneralForm(_lp(" L  r\n", "    X         r    1\n", rhs="    R         r   -2\n", ranges="    G         r    3\n"))
    assert g2.b == [F(-2)]
    g2.standardize()
    assert g2.b == [F(5)] and g2.constraint_types == [("R", F(3))] and g2.columns == [[(0, F(-1))]]


def test_maximisation_negates_the_costs_only():
    g = GeneralForm(_lp(" L  r\n", "    X         r    1   c   5\n", rhs="    R         r    2\n", sense="OBJSENSE\n  MAX\n"))
    g.standardize()
    assert g.objective == "minimize" and g.variables[0].cost == F(-5)
