"""Host-side front-end feeding the engine from MPS/SIF files (SURVEY.md section 8f "next" rows 1-3).

The pipeline is the reference's own (tests/netlib/mod.rs:48-70), restated: `relp_b200/mps.py` reads the file (parse ->
`MPS` -> `GeneralForm` data, rows sorted by name, the reference's bound and range semantics), `relp_b200/presolve.py`
+ `relp_b200/general_form.py` presolve and standardize it and derive the `MatrixData` layout
(src/algorithm/two_phase/matrix_provider/matrix_data.rs:46-61,291-329: equality, range, <=, >= rows; variable-bound
rows; x >= 0 with optional upper bounds).  This module glues them (`parse_mps`, `canonicalize`), maps solutions back
(`recover`) and prescales the rational rows to integers for the device (`prescale`).
"""
from fractions import Fraction
from math import gcd

import numpy as np

from .solver import IntegerProblem

INF = None


def parse_mps(text, mode="auto"):
    """MPS text -> the dict `canonicalize` consumes, through the restatement of relp's own reader (`relp_b200/mps.py`:
    parse -> `MPS` -> `GeneralForm` data; src/io/mps/{parse,convert}.rs).  mode "free" is `io::import` /
    `io::mps::parse` (mod.rs:38-42), "fixed" is `parse_fixed` (what the reference's netlib tests call,
    tests/netlib/mod.rs:53); "auto" tries the flexible mode first and the fixed-column mode when that fails (names
    with blanks).

    Keys: name, objective (cost row name), sense, rows (constraint rows in the reference's order: SORTED BY NAME),
    row_type, interval {row: [lo, hi]} (None = unbounded side; ranges and duplicate right-hand sides resolved as
    convert.rs does), columns {col: {row: Fraction}} (cost row entries included), col_order, bounds {col: [lo, hi]}
    with the reference's bound semantics (GLPK's rule for negative upper bounds)."""
    from . import mps as reader
    if mode == "free":
        m = reader.parse_free(text)
    elif mode == "fixed":
        m = reader.parse_fixed(text)
    else:
        try:
            m = reader.parse_free(text)
        except (reader.ParseError, reader.Inconsistency):
            m = reader.parse_fixed(text)
    gf = m.to_general_form()
    rows = list(gf.row_names)
    row_type, interval = {}, {}
    for r, (_, t), rel, b in zip(rows, m.rows, gf.constraint_types, gf.b):
        row_type[r] = t
        if isinstance(rel, tuple):
            interval[r] = [b - rel[1], b]
        elif rel == "E":
            interval[r] = [b, b]
        elif rel == "L":
            interval[r] = [INF, b]
        else:
            interval[r] = [b, INF]
    columns, col_order, bounds = {}, [], {}
    for (cname, _vt, _vals), values, var in zip(m.columns, gf.columns, gf.variables):
        entries = {rows[i]: v for i, v in values}
        if var.cost != 0:
            entries[m.cost_row_name] = var.cost
        columns[cname] = entries
        col_order.append(cname)
        bounds[cname] = [var.lower_bound, var.upper_bound]         # None = unbounded = INF
    return dict(name=m.name, objective=m.cost_row_name, sense=m.objective, rows=rows, row_type=row_type,
                interval=interval, columns=columns, col_order=col_order, bounds=bounds, rhs={}, ranges={},
                general_form=m.to_general_form())     # a fresh copy: `canonicalize` standardizes it in place


class Solution:
    """`Solution` of the reference (data/linear_program/solution.rs:12-80): objective value and the value of
    every variable of the ORIGINAL problem by name."""

    def __init__(self, objective, values):
        self.objective_value = objective
        self.solution_values = list(values)          # [(name, Fraction)] in the file's column order

    def is_probably_equal_to(self, other, min_equal):
        """solution.rs:47-79: equal objective, equal variable names and more than `min_equal` of the values
        equal (LPs with several optimal vertices: the reference's own tests compare like this)."""
        if self.objective_value != other.objective_value:
            return False
        a, b = dict(self.solution_values), dict(other.solution_values)
        if len(self.solution_values) != len(other.solution_values) or set(a) != set(b):
            return False
        if len(a) < 10:
            return True
        return sum(1 for k in a if a[k] == b[k]) / len(a) > min_equal


class CanonicalLP:
    """min c x + constant over the MatrixData layout; `recover` maps a reduced solution back."""

    def __init__(self):
        self.constraint_columns = []
        self.b = []
        self.ranges = []
        self.counts = (0, 0, 0, 0)
        self.costs = []
        self.upper = []
        self.constant = Fraction(0)
        self.var_map = []      # per canonical column: (original column, +1 / -1, shift)
        self.col_order = []    # original variable names, file order
        self.fixed = {}        # original variables fixed by their bounds (substituted out): name -> value
        self.name = ""
        self.general_form = None   # general_form.GeneralForm after `standardize` (reference path)
        self.maximize = False  # OBJSENSE MAX: costs negated for the solve, objective negated back in `recover`


def recover(lp, bfs, objective):
    """Solution back-substitution (reference: MatrixData::reconstruct_solution, matrix_data.rs:402-411, keeps the
    structural columns of the basic feasible solution; GeneralForm::compute_full_solution_with_reduced_solution +
    reshift_solution, general_form/mod.rs:808-935, undo the shifts, sign flips, free-variable splits and
    substituted fixed variables of the canonicalisation).

    lp: CanonicalLP; bfs: [(provider column, value)] of the solved MatrixData; objective: its optimum.
    Returns a `Solution` over the original variables (file order) with the constant added back."""
    nv = len(lp.constraint_columns)
    reduced = {j: v for j, v in bfs if j < nv}                 # slack / bound-slack columns are dropped
    if getattr(lp, "general_form", None) is not None:           # the reference's reconstruction, mod.rs:800-905
        cost, values = lp.general_form.compute_full_solution_with_reduced_solution(reduced)
        assert cost == objective + lp.constant
        return Solution(cost, values)
    values = {name: Fraction(fixed) for name, fixed in lp.fixed.items()}
    for k, (orig, sign, shift) in enumerate(lp.var_map):
        x = reduced.get(k, Fraction(0))
        if orig in values and orig not in lp.fixed:
            values[orig] += sign * x                           # second half of a split free variable
        else:
            values[orig] = shift + sign * x
    total = objective + lp.constant
    return Solution(-total if getattr(lp, "maximize", False) else total, [(name, values[name]) for name in lp.col_order])


def canonicalize(mps, presolve=True):
    """MPS dict -> CanonicalLP.  With the reference reader's `GeneralForm` data this is the reference's own pipeline
    (tests/netlib/mod.rs:55-59): `GeneralForm::presolve` (`relp_b200/presolve.py`; `presolve=False` skips it),
    `standardize` and `derive_matrix_data` (`relp_b200/general_form.py`): free variables split with the negative
    halves appended, upper-bounded-only variables flipped, lower bounds shifted to zero, negative right-hand sides
    negated, rows ordered ==, ranges, <=, >= (stable).  A presolve that solves the whole problem raises
    `presolve.FiniteOptimum`, an infeasible / unbounded one `presolve.Infeasible` / `Unbounded`.  Dicts without the
    general form (hand-built in tests) take the legacy interval-based path below."""
    if mps.get("general_form") is not None:
        from .general_form import GeneralForm
        g = GeneralForm(mps["general_form"])
        if presolve:
            g.presolve()
        counts = g.standardize()
        cols, b, ranges, ne, nr, nu, nl, variables = g.derive_matrix_data(counts)
        lp = CanonicalLP()
        lp.name = mps["name"]
        lp.col_order = list(g.variable_names)
        lp.constraint_columns = [list(c) for c in cols]
        lp.b, lp.ranges, lp.counts = list(b), list(ranges), (ne, nr, nu, nl)
        lp.costs = [c for c, _ in variables]
        lp.upper = [u for _, u in variables]
        lp.constant = g.fixed_cost
        lp.var_map = [(g.variable_names[g.from_active_to_original[j]], -1 if v.flipped else 1,
                       v.shift if v.flipped else -v.shift) for j, v in enumerate(g.variables)]
        lp.general_form = g
        return lp
    lp = CanonicalLP()
    lp.name = mps["name"]
    lp.col_order = list(mps["col_order"])
    lp.fixed = {}
    rows = mps["rows"]
    obj = mps["objective"]
    lp.constant = Fraction(0)          # the reference's reader rejects a right-hand side on the cost row
    lp.maximize = mps.get("sense", "minimize") == "maximize"
    sense = -1 if lp.maximize else 1   # a maximisation is solved as the minimisation of the negated costs
    interval = {r: list(mps["interval"][r]) for r in rows}
    # variables -> x' >= 0
    new_cols = []   # (entries {row: v}, cost, upper, (orig, sign, shift))
    for col in mps["col_order"]:
        entries = dict(mps["columns"][col])
        cost = sense * entries.pop(obj, Fraction(0))
        entries = {r: v for r, v in entries.items() if r in interval}
        lo, hi = mps["bounds"].get(col, [Fraction(0), INF])
        if lo is not INF:
            shift = lo
            up = None if hi is INF else hi - lo
            if shift != 0:
                lp.constant += cost * shift
                for r, v in entries.items():
                    for k in (0, 1):
                        if interval[r][k] is not INF:
                            interval[r][k] -= v * shift
            if up is not None and up == 0:
                lp.fixed[col] = lo
                continue   # fixed variable: substituted out
            new_cols.append((entries, cost, up, (col, 1, shift)))
        elif hi is not INF:
            # x = hi - x'
            lp.constant += cost * hi
            for r, v in entries.items():
                for k in (0, 1):
                    if interval[r][k] is not INF:
                        interval[r][k] -= v * hi
            new_cols.append(({r: -v for r, v in entries.items()}, -cost, None, (col, -1, hi)))
        else:
            new_cols.append((entries, cost, None, (col, 1, Fraction(0))))
            new_cols.append(({r: -v for r, v in entries.items()}, -cost, None, (col, -1, Fraction(0))))
    # classify rows, make b >= 0
    groups = {"E": [], "R": [], "L": [], "G": []}
    flip = {}
    for r in rows:
        lo, hi = interval[r]
        if lo is not INF and hi is not INF and lo == hi:
            flip[r] = lo < 0
            groups["E"].append((r, -lo if flip[r] else lo, None))
        elif lo is not INF and hi is not INF:
            f = hi < 0
            flip[r] = f
            nlo, nhi = (-hi, -lo) if f else (lo, hi)
            groups["R"].append((r, nhi, nhi - nlo))
        elif hi is not INF:          # a x <= hi
            f = hi < 0
            flip[r] = f
            groups["G" if f else "L"].append((r, -hi if f else hi, None))
        else:                        # a x >= lo
            f = lo < 0
            flip[r] = f
            groups["L" if f else "G"].append((r, -lo if f else lo, None))
    order = groups["E"] + groups["R"] + groups["L"] + groups["G"]
    row_index = {r: i for i, (r, _, _) in enumerate(order)}
    lp.b = [b for _, b, _ in order]
    lp.ranges = [rg for _, _, rg in groups["R"]]
    lp.counts = (len(groups["E"]), len(groups["R"]), len(groups["L"]), len(groups["G"]))
    for entries, cost, up, vm in new_cols:
        col = sorted((row_index[r], (-v if flip[r] else v)) for r, v in entries.items())
        lp.constraint_columns.append(col)
        lp.costs.append(cost)
        lp.upper.append(up)
        lp.var_map.append(vm)
    return lp


def _lcm(a, b):
    return a * b // gcd(a, b)


class ScaledProblem:
    """Integer image of a rational provider + the weights that keep the pivoting rules invariant.

    rows are multiplied by r_i = lcm of the denominators in row i (coefficients and right-hand side);
    single-entry +-1 columns of row i (slacks, bound slacks) and the artificial columns stay UNIT
    columns, which is a column scaling by 1/r_i and is compensated by weights.
    """

    def __init__(self):
        self.problem = None            # IntegerProblem
        self.row_scale = []            # r_i
        self.col_weight = []           # w_j (1 for structural columns, r_i for unit columns of row i)
        self.cost_scale = 1            # integer costs = cost_scale * c_j / w_j
        self.W = 1                     # lcm of all weights


def prescale(m, n, columns, costs, rhs, pivots, full_initial_basis=False):
    """columns: list of [(row, Fraction)], costs, rhs: Fractions."""
    sp = ScaledProblem()
    r = [1] * m
    for j in range(n):
        col = columns[j]
        if len(col) == 1 and abs(col[0][1]) == 1:
            continue                   # unit column: stays +-1, gets a weight instead
        for i, v in col:
            r[i] = _lcm(r[i], Fraction(v).denominator)
    for i in range(m):
        r[i] = _lcm(r[i], Fraction(rhs[i]).denominator)
    w = [1] * n
    int_cols = []
    for j in range(n):
        col = columns[j]
        if len(col) == 1 and abs(col[0][1]) == 1:
            w[j] = r[col[0][0]]
            int_cols.append([(col[0][0], int(col[0][1]))])
        else:
            ic = []
            for i, v in col:
                x = Fraction(v) * r[i]
                assert x.denominator == 1
                ic.append((i, int(x)))
            int_cols.append(ic)
    W = 1
    for x in w + r:
        W = _lcm(W, x)
    cs = 1
    for j in range(n):
        cs = _lcm(cs, (Fraction(costs[j]) / w[j]).denominator)
    int_cost = [int(Fraction(costs[j]) * cs / w[j]) for j in range(n)]
    int_rhs = [int(Fraction(rhs[i]) * r[i]) for i in range(m)]
    prob = IntegerProblem.from_columns(m, int_cols, int_cost, int_rhs, pivots, full_initial_basis)
    prob.col_weight = np.array(w, dtype=np.int64)
    prob.row_scale = np.array(r, dtype=np.int64)
    prob.W = W
    prob.cost_scale = cs
    if W >= 2 ** 31:
        raise ValueError("weight lcm exceeds 2^31: rows need a coarser common scale")
    sp.problem, sp.row_scale, sp.col_weight, sp.cost_scale, sp.W = prob, r, w, cs, W
    return sp
