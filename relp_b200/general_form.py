"""`GeneralForm::standardize` and `derive_matrix_data`, restated (the canonicalisation
half): reference `src/data/linear_program/general_form/mod.rs`

    standardize            :325-332   transform_variables, make_b_non_negative, make_minimization_problem,
                                      reorder_constraints_by_type
    transform_variables    :506-547   split free variables, flip upper-bounded-only ones, shift lower bounds to 0
    split_free_variables   :549-581   the negative duplicates are APPENDED after all original columns
    make_b_non_negative    :583-611
    reorder_constraints_by_type :623-684   stable order  ==, =r=, <=, >=
    derive_matrix_data     :262-302
    reshift_solution / compute_full_solution_with_reduced_solution :800-905

`GeneralForm.presolve` (mod.rs:333-505, rules in `relp_b200/presolve.py`) is optional, as in the reference: without
it fixed variables stay as columns with upper bound 0 and redundant rows stay.  Values are `fractions.Fraction`.
Host-side only.
"""
from fractions import Fraction


class GeneralForm:
    """Mutable working copy of `mps.GeneralFormData` with the bookkeeping `standardize` needs.

    original_variables: per original variable ("active", j), ("active_free", j_positive, j_negative) or
    ("removed", ("solved", value) | ("function", constant, [(original index, coefficient)])) --
    `OriginalVariable::{Active, ActiveFree, Removed}` (mod.rs:72-99)."""

    def __init__(self, data):
        self.objective = data.objective
        self.columns = [list(col) for col in data.columns]            # column-major [(row, value)]
        self.nr_rows = data.nr_rows
        self.constraint_types = list(data.constraint_types)           # "E" | "L" | "G" | ("R", r)
        self.b = list(data.b)
        self.variables = data.variables                               # mps.Variable (mutated in place)
        self.variable_names = list(data.variable_names)
        self.row_names = list(data.row_names)
        self.fixed_cost = Fraction(data.fixed_cost)
        self.original_variables = [("active", j) for j in range(len(self.variables))]
        self.from_active_to_original = list(range(len(self.variables)))

    # -- presolve, mod.rs:333-505 (rules: relp_b200/presolve.py) ---------------------------------------
    def presolve(self):
        """`GeneralForm::presolve` (mod.rs:352-369).  Raises presolve.Infeasible / Unbounded, or FiniteOptimum when
        every variable was solved.  Must run on a fresh problem, before `standardize` (as in the reference's
        pipeline, tests/netlib/mod.rs:55-57)."""
        from . import presolve as ps
        ch = ps.compute_presolve_changes(self)
        # update_values_that_remain, mod.rs:404-436
        for i, v in ch["b"].items():
            self.b[i] = v
        for i, t in ch["constraints"].items():
            self.constraint_types[i] = t
        self.fixed_cost += ch["fixed_cost"]
        for j, solution in ch["removed_variables"]:
            self.original_variables[j] = ("removed", solution)
        for (j, direction), value in ch["bounds"].items():
            if direction == ps.LOWER:
                self.variables[j].lower_bound = value
            else:
                self.variables[j].upper_bound = value
        self.remove_rows_and_columns(ch["constraints_marked_removed"], [j for j, _ in ch["removed_variables"]])
        self.compute_solution_where_possible()
        solution = self.get_solution()
        if solution is not None:
            raise ps.FiniteOptimum(*solution)

    def remove_rows_and_columns(self, constraints, variables):
        """mod.rs:438-475"""
        gone_v, gone_c = set(variables), set(constraints)
        keep_v = [j for j in range(len(self.variables)) if j not in gone_v]
        self.columns = [self.columns[j] for j in keep_v]
        self.variables = [self.variables[j] for j in keep_v]
        self.from_active_to_original = [self.from_active_to_original[j] for j in keep_v]
        if variables:
            for new_index, orig in enumerate(self.from_active_to_original):
                assert self.original_variables[orig][0] == "active"
                self.original_variables[orig] = ("active", new_index)
        new_row, k = {}, 0
        for i in range(len(self.b)):
            if i not in gone_c:
                new_row[i] = k
                k += 1
        self.columns = [[(new_row[i], v) for i, v in col if i not in gone_c] for col in self.columns]
        keep_c = [i for i in range(len(self.b)) if i not in gone_c]
        self.constraint_types = [self.constraint_types[i] for i in keep_c]
        self.b = [self.b[i] for i in keep_c]
        self.row_names = [self.row_names[i] for i in keep_c]
        self.nr_rows = len(self.b)

    def _solution_value(self, j, cache):
        """compute_solution_value, mod.rs:712-737: None while it depends on a variable still in the problem"""
        ov = self.original_variables[j]
        if ov[0] != "removed":
            return None
        sol = ov[1]
        if sol[0] == "solved":
            return sol[1]
        if j not in cache:
            total = Fraction(0)
            for k, coefficient in sol[2]:
                v = self._solution_value(k, cache)
                if v is None:
                    total = None
                    break
                total += coefficient * v
            cache[j] = None if total is None else sol[1] - total
        return cache[j]

    def compute_solution_where_possible(self):
        """mod.rs:686-710"""
        cache = {}
        for j, ov in enumerate(self.original_variables):
            if ov[0] == "removed" and ov[1][0] == "function":
                self._solution_value(j, cache)
        for j, value in cache.items():
            if value is not None:
                self.original_variables[j] = ("removed", ("solved", value))

    def get_solution(self):
        """mod.rs:739-752: (fixed cost, [(name, value)]) once every variable is solved"""
        values = []
        for name, ov in zip(self.variable_names, self.original_variables):
            if ov[0] == "removed" and ov[1][0] == "solved":
                values.append((name, ov[1][1]))
            else:
                return None
        return self.fixed_cost, values

    # -- transform_variables, mod.rs:506-547 -------------------------------------------------------
    def split_free_variables(self):
        """mod.rs:549-581"""
        from .mps import Variable
        free = [j for j, v in enumerate(self.variables) if v.lower_bound is None and v.upper_bound is None]
        new_columns = [[(i, -val) for i, val in self.columns[j]] for j in free]
        self.columns.extend(new_columns)
        for j in free:
            orig = self.from_active_to_original[j]
            self.original_variables[orig] = ("active_free", j, len(self.from_active_to_original))
            self.from_active_to_original.append(orig)
            neg = Variable(self.variables[j].variable_type, -self.variables[j].cost)
            neg.lower_bound = Fraction(0)
            self.variables.append(neg)
            self.variables[j].lower_bound = Fraction(0)

    def transform_variables(self):
        self.split_free_variables()
        for j, var in enumerate(self.variables):
            if var.lower_bound is None and var.upper_bound is not None:      # flip: at least a lower bound
                var.flipped = not var.flipped
                var.shift = -var.shift
                var.cost = -var.cost
                var.lower_bound = -var.upper_bound
                var.upper_bound = None
                self.columns[j] = [(i, -val) for i, val in self.columns[j]]
            if var.lower_bound is not None:                                   # shift: the lower bound becomes 0
                lower = var.lower_bound
                var.shift -= lower
                if var.upper_bound is not None:
                    var.upper_bound -= lower
                self.fixed_cost += lower * var.cost
                for i, val in self.columns[j]:
                    self.b[i] -= val * lower
                var.lower_bound = Fraction(0)

    # -- make_b_non_negative, mod.rs:583-611 ---------------------------------------------------------
    def make_b_non_negative(self):
        negate = {i for i, v in enumerate(self.b) if v < 0}
        self.columns = [[(i, -val if i in negate else val) for i, val in col] for col in self.columns]
        for i in sorted(negate):
            t = self.constraint_types[i]
            if t == "L":
                self.constraint_types[i] = "G"
                self.b[i] = -self.b[i]
            elif t == "E":
                self.b[i] = -self.b[i]
            elif t == "G":
                self.constraint_types[i] = "L"
                self.b[i] = -self.b[i]
            else:                       # ("R", r): b is the upper end; the negated interval's upper end is r - b
                self.b[i] = t[1] - self.b[i]

    def make_minimization_problem(self):
        """mod.rs:613-621 (the costs stay negated: a maximisation reports the minimum of the negated objective, as the
        reference does)"""
        if self.objective == "maximize":
            self.objective = "minimize"
            for var in self.variables:
                var.cost = -var.cost

    # -- reorder_constraints_by_type, mod.rs:623-684 -------------------------------------------------
    def reorder_constraints_by_type(self):
        kind = ["R" if isinstance(t, tuple) else t for t in self.constraint_types]
        order = {"E": 0, "R": 1, "L": 2, "G": 3}
        counts = [sum(1 for k in kind if k == key) for key in ("E", "R", "L", "G")]
        starts = [0, counts[0], counts[0] + counts[1], counts[0] + counts[1] + counts[2]]
        seen = [0, 0, 0, 0]
        destination = []
        for k in kind:                                 # stable within a type
            g = order[k]
            destination.append(starts[g] + seen[g])
            seen[g] += 1
        m = len(self.b)
        new_b, new_types, new_names = [None] * m, [None] * m, [None] * m
        for i in range(m):
            new_b[destination[i]] = self.b[i]
            new_types[destination[i]] = self.constraint_types[i]
            new_names[destination[i]] = self.row_names[i]
        self.b, self.constraint_types, self.row_names = new_b, new_types, new_names
        self.columns = [sorted((destination[i], val) for i, val in col) for col in self.columns]
        return counts

    def standardize(self):
        """mod.rs:325-332; returns [nr_equality, nr_range, nr_upper(<=), nr_lower(>=)]"""
        self.transform_variables()
        self.make_b_non_negative()
        self.make_minimization_problem()
        return self.reorder_constraints_by_type()

    # -- derive_matrix_data, mod.rs:262-302 ------------------------------------------------------------
    def derive_matrix_data(self, counts):
        """Arguments of `MatrixData::new`: (constraint columns, b, ranges, nr_equality, nr_range, nr_upper,
        nr_lower, [(cost, upper bound)])"""
        ne, nr, nu, nl = counts
        assert ne + nr + nu + nl == len(self.b)
        ranges = [t[1] for t in self.constraint_types[ne:ne + nr]]
        return (self.columns, self.b, ranges, ne, nr, nu, nl, [(v.cost, v.upper_bound) for v in self.variables])

    # -- solution reconstruction, mod.rs:800-905 ---------------------------------------------------------
    def compute_full_solution_with_reduced_solution(self, reduced):
        """reduced: {active variable index: value} (zeros absent) -- `MatrixData::reconstruct_solution` has already
        dropped the slack columns (matrix_data.rs:402-411).  Returns (objective value, [(name, value)])."""
        cost = sum((v * self.variables[j].cost for j, v in reduced.items()), Fraction(0)) + self.fixed_cost
        x = {}
        for j, var in enumerate(self.variables):            # reshift_solution
            v = reduced.get(j, Fraction(0)) - var.shift
            if var.flipped:
                v = -v
            x[j] = v
        done = {}

        def value_of(k):                                    # compute_solution_value_with_bfs, mod.rs:856-905
            if k not in done:
                ov = self.original_variables[k]
                if ov[0] == "active":
                    done[k] = x[ov[1]]
                elif ov[0] == "active_free":
                    done[k] = x[ov[1]] - x[ov[2]]
                elif ov[1][0] == "solved":
                    done[k] = ov[1][1]
                else:
                    done[k] = ov[1][1] - sum((c * value_of(q) for q, c in ov[1][2]), Fraction(0))
            return done[k]

        values = [(name, value_of(k)) for k, name in enumerate(self.variable_names)]
        return cost, values
