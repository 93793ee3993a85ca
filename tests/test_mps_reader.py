"""MPS reader (relp_b200/mps.py) against the reference's own parser tests and fixtures:
src/io/mps/parse/mod.rs:834-962 (line filter, program name, row section, consistency), src/io/mps/number/parse.rs:
128-305 (number reading), src/io/mps/convert.rs:524-585 (compute_b), src/tests/problem_1.rs:110-262 (the expected
`MPS` and `GeneralForm` of the TESTPROB file), and the netlib / Burkardt files the reference's integration tests read
(tests/netlib, tests/burkardt): fixed and free mode must agree on them."""
import os
from fractions import Fraction as F

import pytest

from relp_b200 import mps

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# src/io/mps/parse/mod.rs:841-868
MPS_LITERAL_STRING = """
* Start of the file
NAME          TESTPROB

ROWS
* This is the cost row
 N  COST
 L  LIM1
 G  LIM2
 E  MYEQN
COLUMNS
    XONE      COST                 1   LIM1                 1
    XONE      LIM2                 1
    MARKER0   'MARKER'      'INTORG'
    YTWO      COST                 4   LIM1                 1
    YTWO      MYEQN               -1
    MARKER0   'MARKER'      'INTEND'
    ZTHREE    COST                 9   LIM2                 1
    ZTHREE    MYEQN                1
RHS
    RHS1      LIM1                 5   LIM2                10
    RHS1      MYEQN                7
BOUNDS
 UP BND1      XONE                 4
 LO BND1      YTWO                -1
 UP BND1      YTWO                 1
ENDATA"""

# src/tests/problem_1.rs:110-132 (no comments, no blank lines)
PROBLEM_1 = "\n".join(l for l in MPS_LITERAL_STRING.splitlines() if l and not l.startswith("*"))


def test_into_lines():
    """parse/mod.rs:870-880"""
    lines = mps.into_lines(MPS_LITERAL_STRING)
    assert lines[:3] == [(3, "NAME          TESTPROB"), (5, "ROWS"), (7, " N  COST")]
    assert lines[-1] == (27, "ENDATA")


@pytest.mark.parametrize("cr", [mps.Free, mps.Fixed])
def test_parse_program_name(cr):
    """parse/mod.rs:882-895"""
    assert mps.parse_program_name(mps.into_lines(MPS_LITERAL_STRING)[0], cr) == "TESTPROB"
    with pytest.raises(mps.ParseError):
        mps.parse_program_name(None, cr)
    with pytest.raises(mps.ParseError):
        mps.parse_program_name((1, "NAM"), cr)
    with pytest.raises(mps.ParseError):
        mps.parse_program_name((1, "ROWS and more"), cr)


@pytest.mark.parametrize("parse", [mps.parse_free, mps.parse_fixed])
def test_row_section_and_expected_mps(parse):
    """parse/mod.rs:897-926 (rows LIM1 / LIM2 / MYEQN, cost row COST) and the whole expected `MPS` of
    src/tests/problem_1.rs:135-198.  The literal has its marker keyword in field 4; the fixed-column reader wants it
    in field 5 (columns 40-47, parse/fixed.rs:62-70) and rejects the literal as the reference's does, so the
    fixed-mode run uses the same file with the keyword moved."""
    if parse is mps.parse_fixed:
        with pytest.raises(mps.ParseError):
            parse(PROBLEM_1)
        moved = PROBLEM_1.replace("'MARKER'      'INTORG'", "'MARKER'                 'INTORG'")
        moved = moved.replace("'MARKER'      'INTEND'", "'MARKER'                 'INTEND'")
        got = parse(moved)
    else:
        got = parse(PROBLEM_1)
    expected = mps.MPS(
        "TESTPROB", "minimize", "COST", [(0, F(1)), (1, F(4)), (2, F(9))],
        [("LIM1", "L"), ("LIM2", "G"), ("MYEQN", "E")],
        [("XONE", "continuous", [(0, F(1)), (1, F(1))]),
         ("YTWO", "integer", [(0, F(1)), (2, F(-1))]),
         ("ZTHREE", "continuous", [(1, F(1)), (2, F(1))])],
        [("RHS1", [(0, F(5)), (1, F(10)), (2, F(7))])],
        [],
        [("BND1", [(0, ("UP", F(4))), (1, ("LO", F(-1))), (1, ("UP", F(1)))])])
    assert got == expected
    if parse is mps.parse_free:
        assert parse(MPS_LITERAL_STRING) == expected     # comments and empty lines are skipped


def test_expected_general_form_of_problem_1():
    """src/tests/problem_1.rs:201-258"""
    gf = mps.parse(PROBLEM_1).to_general_form()
    assert gf.objective == "minimize" and gf.fixed_cost == 0
    assert gf.columns == [[(0, F(1)), (1, F(1))], [(0, F(1)), (2, F(-1))], [(1, F(1)), (2, F(1))]]
    assert gf.constraint_types == ["L", "G", "E"]
    assert gf.b == [F(5), F(10), F(7)]
    v = gf.variables
    assert [(x.variable_type, x.cost, x.lower_bound, x.upper_bound, x.shift, x.flipped) for x in v] == [
        ("continuous", F(1), F(0), F(4), F(0), False),
        ("integer", F(4), F(-1), F(1), F(0), False),
        ("continuous", F(9), F(0), None, F(0), False)]
    assert gf.variable_names == ["XONE", "YTWO", "ZTHREE"]


def test_check_row_section_consistency():
    """parse/mod.rs:928-961: no cost row; duplicate rows; cost row among the rows; a valid one.  Rows come out sorted
    by name."""
    head = "NAME x\nROWS\n"
    tail = "COLUMNS\n    X         b                    1\nENDATA"
    with pytest.raises(mps.Inconsistency):
        mps.parse(head + " E  b\n" + tail)                                   # no cost row
    with pytest.raises(mps.Inconsistency):
        mps.parse(head + " N  a\n E  b\n E  b\n" + tail)                     # duplicate
    with pytest.raises(mps.Inconsistency):
        mps.parse(head + " N  a\n E  a\n E  b\n" + tail)                     # cost row name among the rows
    got = mps.parse(head + " N  a\n E  c\n E  b\n" + tail)
    assert got.rows == [("b", "E"), ("c", "E")]
    assert got.columns == [("X", "continuous", [(0, F(1))])]
    with pytest.raises(mps.ParseError):
        mps.parse(head + " N  a\n N  a2\n E  b\n" + tail)                    # second cost row


def test_number_parsing():
    """number/parse.rs:130-300: Raw::try_from and its conversion"""
    cases = {"1": F(1), "2.": F(2), ".3": F(3, 10), "-1": F(-1), "-2.": F(-2), "-.3": F(-3, 10), "16456": F(16456),
             "64896848.": F(64896848), ".984654684": F(984654684, 10 ** 9), "-.95": F(-95, 100),
             "15465.2": F(154652, 10), "1234.56789": F(123456789, 10 ** 5), "1.24654": F(124654, 10 ** 5),
             "0": F(0), "0.": F(0), ".0": F(0), "-0": F(0)}
    for text, value in cases.items():
        assert mps.parse_number(text) == value, text
    for bad in ("1e5", "+1", "1.2.3", "abc", "1 2", "--1", ""):
        with pytest.raises(mps.ParseError):
            mps.parse_number(bad)


def test_compute_b():
    """convert.rs:524-585 and the RANGES table of io/mps/mod.rs:221-229"""
    assert mps.compute_b([], [], [], 0) == []
    assert mps.compute_b([("R", [(0, F(1))])], ["E"], [("", "E")], 1) == [F(1)]
    assert mps.compute_b([("R1", [(0, F(1))]), ("R2", [(0, F(2))])], ["G"], [("", "G")], 1) == [F(2)]
    assert mps.compute_b([("R1", [(0, F(1))]), ("R2", [(0, F(2))])], ["L"], [("", "L")], 1) == [F(1)]
    assert mps.compute_b([("R", [(0, F(1)), (1, F(5))])], ["G", "E"], [("", "G"), ("", "E")], 2) == [F(1), F(5)]
    with pytest.raises(mps.Inconsistency):
        mps.compute_b([("R1", [(0, F(1))]), ("R2", [(0, F(2))])], ["E"], [("", "E")], 1)
    # ranges: b is the upper end of [h, u]
    for t, r, b, u in (("G", F(3), F(10), F(13)), ("G", F(-3), F(10), F(13)), ("L", F(3), F(10), F(10)),
                       ("L", F(-3), F(10), F(10)), ("E", F(3), F(10), F(13)), ("E", F(-3), F(10), F(10))):
        c = [("R", r)]
        assert mps.compute_b([("R", [(0, b)])], c, [("", t)], 1) == [u]
        assert c == [("R", abs(r))]
    assert mps.compute_b([], ["L"], [("", "L")], 1) == [F(0)]                # default right-hand side


def test_bounds_semantics():
    """convert.rs:107-224: GLPK's rule for the implied zero lower bound, free / bounded conflicts, integer bounds"""
    def bounds_of(lines):
        text = ("NAME b\nROWS\n N  c\n L  r\nCOLUMNS\n    X         r                    1\n"
                "    Y         r                    1\nBOUNDS\n" + lines + "ENDATA")
        v = mps.parse(text).to_general_form().variables
        return [(x.lower_bound, x.upper_bound, x.variable_type) for x in v]
    assert bounds_of(" UP B         X                   -4\n") == [(F(0), F(-4), "continuous"), (F(0), None, "continuous")]
    assert bounds_of(" LO B         X                   -4\n UP B         X                    9\n")[0] == (F(-4), F(9), "continuous")
    assert bounds_of(" MI B         X\n")[0] == (None, F(0), "continuous")
    assert bounds_of(" PL B         X\n")[0] == (F(0), None, "continuous")
    assert bounds_of(" FR B         X\n")[0] == (None, None, "continuous")
    assert bounds_of(" FX B         X                    3\n")[0] == (F(3), F(3), "continuous")
    assert bounds_of(" BV B         X\n")[0] == (F(0), F(1), "integer")
    assert bounds_of(" UI B         X                    7\n")[0] == (F(0), F(7), "integer")
    assert bounds_of(" LI B         X                    2\n")[0] == (F(2), None, "integer")
    assert bounds_of(" UP B         X                    5\n UP B2        X                    3\n")[0] == (F(0), F(3), "continuous")
    with pytest.raises(mps.Inconsistency):
        bounds_of(" UP B         X                    5\n FR B         X\n")
    with pytest.raises(mps.Inconsistency):
        bounds_of(" FR B         X\n UP B         X                    5\n")
    with pytest.raises(mps.Inconsistency):
        bounds_of(" UP B         Z                    5\n")                 # unknown column
    with pytest.raises(mps.ParseError):
        bounds_of(" XX B         X                    5\n")


def test_unknown_names_sections_and_trailing_lines():
    base = "NAME t\nROWS\n N  c\n L  r\nCOLUMNS\n    X         r                    1\n"
    with pytest.raises(mps.Inconsistency):
        mps.parse(base.replace("    X         r ", "    X         q ") + "ENDATA")
    with pytest.raises(mps.Inconsistency):
        mps.parse(base + "RHS\n    R         q                    1\nENDATA")
    with pytest.raises(mps.Inconsistency):          # the cost row is not a row of the RHS section (parse/mod.rs:598-600)
        mps.parse(base + "RHS\n    R         c                    1\nENDATA")
    with pytest.raises(mps.ParseError):
        mps.parse(base + "BOUNDS\n UP B         X                    1\nRHS\nENDATA")   # section order
    with pytest.raises(mps.ParseError):
        mps.parse(base + "ENDATA\n junk")
    with pytest.raises(mps.Inconsistency):
        mps.parse(base + "RANGES\n    R         r                    1\n    R2        r                    2\nENDATA")
    with pytest.raises(mps.Inconsistency):
        mps.parse(base + "    X         r                    2\nENDATA")               # duplicate row in a column
    got = mps.parse("NAME t\nOBJSENSE\n  MAX\nROWS\n N  c\n L  r\nCOLUMNS\n    X         r    1   c   2\nENDATA")
    assert got.objective == "maximize" and got.cost_values == [(0, F(2))]


@pytest.mark.parametrize("fname", ["afiro.mps", "adlittle.mps", "maros.mps", "testprob.mps", "AFIRO.SIF",
                                   "ADLITTLE.SIF", "SC205.SIF"])
def test_fixed_and_free_agree_on_the_reference_fixtures(fname):
    """the files of tests/netlib and tests/burkardt of the reference parse identically in both modes; rows are in
    name order and the shapes are the published ones"""
    text = open(os.path.join(GOLDEN, fname)).read()
    a, b = mps.parse_free(text), mps.parse_fixed(text)
    assert a == b
    assert [r[0] for r in a.rows] == sorted(r[0] for r in a.rows)
    shapes = {"afiro": (27, 32, 83), "adlittle": (56, 97, 383), "sc205": (205, 203, 551)}
    key = fname.split(".")[0].lower()
    if key in shapes:
        m, n, nnz = shapes[key]
        assert len(a.rows) == m and len(a.columns) == n
        assert sum(len(c[2]) for c in a.columns) == nnz          # nnz(A) of the structural part
    gf = a.to_general_form()
    assert len(gf.b) == len(a.rows) and len(gf.variables) == len(a.columns)
