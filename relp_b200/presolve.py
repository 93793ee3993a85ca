"""`GeneralForm::presolve`, restated (the presolve half): reference
`src/data/linear_program/general_form/presolve/**` and `general_form/mod.rs:333-505`.

    Index / presolve_step / queues by counter     presolve/mod.rs:22-259
    Counters (row / column counts, activity)      presolve/counters.rs
    Queues (three stacks, one FIFO set)           presolve/queues.rs
    Updates (pending changes, bound lookups)      presolve/updates.rs
    rules: fixed variable, bound constraint,      presolve/rule/{fixed_variable,bound_constraint,
           slack, domain propagation                             slack,domain_propagation}.rs
    applying the changes                          general_form/mod.rs:378-505

The order of everything is the reference's: the substitution, bound and slack queues are `Vec`s used as stacks, the
activity queue is a FIFO with set semantics (crate `fifo-set`: a push of an element already queued is ignored, pop
takes the oldest), a step applies at most one rule, the loop stops when the queues are empty or after as many
consecutive not-meaningful steps as there are rows and columns left.  Hash maps of the reference are only ever
iterated where the result does not depend on the order.

Outcomes the reference returns as `Err(LinearProgramType::..)` are exceptions here: `Infeasible`, `Unbounded`,
`FiniteOptimum` (the presolve solved the whole problem).  Host-side only, `fractions.Fraction` arithmetic.
"""
from fractions import Fraction

LOWER, UPPER = 0, 1
MEANINGFUL, NOT_MEANINGFUL, NO_CHANGE = "meaningful", "not_meaningful", "none"


class Infeasible(Exception):
    pass


class Unbounded(Exception):
    pass


class FiniteOptimum(Exception):
    """The presolve solved the problem: `.objective`, `.values` = [(name, value)] (general_form/mod.rs:366-368)."""

    def __init__(self, objective, values):
        super().__init__("solved by presolve")
        self.objective = objective
        self.values = values


def _sign(v):
    return 1 if v > 0 else -1


def _times(direction, coefficient):
    """`BoundDirection * NonZeroSign`, elements.rs:149-158: a negative coefficient flips the side"""
    return direction if coefficient > 0 else 1 - direction


class FIFOSet:
    def __init__(self, items=()):
        self.items, self.members = [], set()
        for x in items:
            self.push(x)

    def push(self, x):
        if x not in self.members:
            self.members.add(x)
            self.items.append(x)

    def pop(self):
        if not self.items:
            return None
        x = self.items.pop(0)
        self.members.discard(x)
        return x

    def __len__(self):
        return len(self.items)


def is_empty_constraint_feasible(rhs, ctype):
    """presolve/mod.rs:261-287"""
    if ctype == "E":
        return rhs == 0
    if isinstance(ctype, tuple):
        return rhs >= 0 and rhs - ctype[1] <= 0
    if ctype == "L":
        return rhs >= 0
    return rhs <= 0


def optimize_independent_column(objective, cost, lower, upper):
    """updates.rs:357-379"""
    assert cost != 0
    if (objective == "minimize") == (cost > 0):
        if lower is None:
            raise Unbounded()
        return lower
    if upper is None:
        raise Unbounded()
    return upper


def _feasible_value(lower, upper):
    """Variable::get_feasible_value (mod.rs:1046-1050) / Updates::variable_feasible_value (updates.rs:131-149)"""
    if lower is None and upper is None:
        return Fraction(0)
    if lower is None:
        return upper
    if upper is None:
        return lower
    return upper if lower <= upper else None


class Index:
    """presolve/mod.rs:22-59"""

    def __init__(self, gf):
        self.gf = gf
        n, m = len(gf.variables), len(gf.b)
        self.rows = [[] for _ in range(m)]                    # Counters: row-major copy, entries by column index
        for j, col in enumerate(gf.columns):
            for i, v in col:
                self.rows[i].append((j, v))
        self.count_variable = [len(col) for col in gf.columns]
        self.count_constraint = [len(r) for r in self.rows]
        self.count_activity = []
        for row in self.rows:                                 # counters.rs:37-56: bounds missing on either side
            lo = up = 0
            for j, c in row:
                var = gf.variables[j]
                lower, upper = (var.lower_bound, var.upper_bound) if c > 0 else (var.upper_bound, var.lower_bound)
                lo += lower is None
                up += upper is None
            self.count_activity.append([lo, up])
        # Updates::new, updates.rs:44-94
        self.b, self.constraints, self.bounds, self.activity_var_bounds = {}, {}, {}, {}
        self.fixed_cost = Fraction(0)
        self.removed_variables = []
        for j in range(n):
            if self.count_variable[j] == 0:
                var = gf.variables[j]
                if var.cost == 0:
                    value = _feasible_value(var.lower_bound, var.upper_bound)
                else:
                    value = optimize_independent_column(gf.objective, var.cost, var.lower_bound, var.upper_bound)
                    self.fixed_cost += var.cost * value
                self.removed_variables.append((j, ("solved", value)))
        self.constraints_marked_removed = []
        for i in range(m):
            if self.count_constraint[i] == 0:
                if not is_empty_constraint_feasible(gf.b[i], gf.constraint_types[i]):
                    raise Infeasible()
                self.constraints_marked_removed.append(i)
        # Queues::new, queues.rs:33-66
        self.q_bound = [i for i in range(m) if self.count_constraint[i] == 1]
        self.q_activity = FIFOSet()
        for i in range(m):
            if self.count_constraint[i] > 1:
                if self.count_activity[i][LOWER] <= 1:
                    self.q_activity.push((i, LOWER))
                if self.count_activity[i][UPPER] <= 1:
                    self.q_activity.push((i, UPPER))
        self.q_slack = [j for j in range(n) if self.count_variable[j] == 1 and gf.variables[j].cost == 0]
        self.q_substitution = [j for j in range(n) if self.count_variable[j] > 0 and self._original_fixed(j)]
        self.activity_bounds = [[None, None] for _ in range(m)]

    def _original_fixed(self, j):
        var = self.gf.variables[j]
        return var.lower_bound is not None and var.lower_bound == var.upper_bound

    # -- counters -----------------------------------------------------------------------------------
    def constraint_active(self, i):
        return self.count_constraint[i] > 0

    def variable_active(self, j):
        return self.count_variable[j] > 0

    def iter_active_column(self, j):
        return [(i, v) for i, v in self.gf.columns[j] if self.count_constraint[i] > 0]

    def iter_active_row(self, i):
        return [(j, v) for j, v in self.rows[i] if self.count_variable[j] > 0]

    def queues_empty(self):
        return not (len(self.q_activity) or self.q_slack or self.q_bound or self.q_substitution)

    # -- updates (updates.rs) -------------------------------------------------------------------------
    def get_b(self, i):
        return self.b.get(i, self.gf.b[i])

    def change_b(self, i, change):
        self.b[i] = self.get_b(i) + change

    def constraint_type(self, i):
        return self.constraints.get(i, self.gf.constraint_types[i])

    def variable_bound(self, j, direction):
        """updates.rs:151-170: activity-derived, then derived, then original"""
        key = (j, direction)
        if key in self.activity_var_bounds:
            return self.activity_var_bounds[key]
        if key in self.bounds:
            return self.bounds[key]
        var = self.gf.variables[j]
        return var.lower_bound if direction == LOWER else var.upper_bound

    def is_variable_fixed(self, j):
        lo, up = self.variable_bound(j, LOWER), self.variable_bound(j, UPPER)
        return lo if lo is not None and lo == up else None

    def variable_feasible_value(self, j):
        return _feasible_value(self.variable_bound(j, LOWER), self.variable_bound(j, UPPER))

    @staticmethod
    def _compare_and_update(key, new, existing, table):
        """bound_compare_and_update, updates.rs:335-355"""
        if (new > existing) if key[1] == LOWER else (new < existing):
            table[key] = new
            return ("shift", new - existing)
        return ("none",)

    def update_bound(self, j, direction, new):
        """updates.rs:172-211"""
        key = (j, direction)
        if key in self.bounds:
            compare_with = self.bounds[key]
        elif key in self.activity_var_bounds:
            compare_with = self.bounds[key] = self.activity_var_bounds.pop(key)
        else:
            var = self.gf.variables[j]
            original = var.lower_bound if direction == LOWER else var.upper_bound
            if original is None:
                self.bounds[key] = new
                return ("new",)
            compare_with = original
        return self._compare_and_update(key, new, compare_with, self.bounds)

    def update_activity_variable_bound(self, j, direction, new):
        """updates.rs:213-255"""
        key = (j, direction)
        if key in self.activity_var_bounds:
            return self._compare_and_update(key, new, self.activity_var_bounds[key], self.activity_var_bounds)
        if key in self.bounds:
            return self._compare_and_update(key, new, self.bounds[key], self.bounds)
        var = self.gf.variables[j]
        original = var.lower_bound if direction == LOWER else var.upper_bound
        if original is None:
            self.activity_var_bounds[key] = new
            return ("new",)
        return self._compare_and_update(key, new, original, self.activity_var_bounds)

    def optimize_column_independently(self, j):
        var = self.gf.variables[j]
        value = optimize_independent_column(self.gf.objective, var.cost, self.variable_bound(j, LOWER),
                                            self.variable_bound(j, UPPER))
        self.fixed_cost += var.cost * value
        return ("solved", value)

    def nr_variables_remaining(self):
        return len(self.gf.variables) - len(self.removed_variables)

    def nr_constraints_remaining(self):
        return len(self.gf.b) - len(self.constraints_marked_removed)

    def into_changes(self):
        """updates.rs:277-325"""
        for i in self.constraints_marked_removed:
            self.b.pop(i, None)
            self.constraints.pop(i, None)
        for j, _ in self.removed_variables:
            for d in (LOWER, UPPER):
                self.bounds.pop((j, d), None)
                self.activity_var_bounds.pop((j, d), None)
        free_to_be_restricted = {j for (j, _d) in self.activity_var_bounds
                                 if self.gf.variables[j].lower_bound is None and self.gf.variables[j].upper_bound is None
                                 and (j, LOWER) not in self.bounds and (j, UPPER) not in self.bounds}
        for (j, d), value in list(self.activity_var_bounds.items()):
            if j in free_to_be_restricted:
                self.bounds[(j, d)] = value
        b = {i: v for i, v in self.b.items() if v != self.gf.b[i]}
        constraints = {i: t for i, t in self.constraints.items() if t != self.gf.constraint_types[i]}
        return dict(b=b, constraints=constraints, fixed_cost=self.fixed_cost, bounds=dict(self.bounds),
                    removed_variables=sorted(self.removed_variables, key=lambda t: t[0]),
                    constraints_marked_removed=sorted(self.constraints_marked_removed))

    # -- the step (presolve/mod.rs:61-93) ---------------------------------------------------------------
    def presolve_step(self):
        if self.q_substitution:
            variable = self.q_substitution.pop()
            if self.variable_active(variable):
                self.presolve_fixed_variable(variable)
                return MEANINGFUL
        while self.q_bound:
            constraint = self.q_bound.pop()
            if self.constraint_active(constraint):
                self.presolve_bound_constraint(constraint)
                return MEANINGFUL
        while self.q_slack:
            variable = self.q_slack.pop()
            if self.variable_active(variable):
                self.presolve_slack(variable)
                return MEANINGFUL
        while len(self.q_activity):
            constraint, direction = self.q_activity.pop()
            if self.constraint_active(constraint):
                return self.presolve_domain_propagation(constraint, direction)
        return NOT_MEANINGFUL

    # -- shared bookkeeping (presolve/mod.rs:95-259) -----------------------------------------------------
    def after_bound_change(self, variable, direction, change):
        if self.is_variable_fixed(variable) is not None and self.variable_active(variable):
            self.q_substitution.append(variable)
        if change is not None:
            self.update_activity_bounds(variable, direction, change)
        else:
            self.update_activity_counters(variable, direction)

    def update_activity_bounds(self, variable, direction, by_how_much):
        for row, coefficient in self.iter_active_column(variable):
            if not self.constraint_active(row):
                continue
            side = _times(direction, coefficient)
            if self.activity_bounds[row][side] is not None:
                self.activity_bounds[row][side] += by_how_much * coefficient
                self.q_activity.push((row, side))

    def update_activity_counters(self, variable, direction):
        for constraint, coefficient in self.iter_active_column(variable):
            side = _times(direction, coefficient)
            self.count_activity[constraint][side] -= 1
            if self.count_activity[constraint][side] <= 1:
                self.q_activity.push((constraint, side))

    def remove_constraint_values(self, constraint):
        for variable in [j for j, _ in self.iter_active_row(constraint)]:
            self.count_constraint[constraint] -= 1
            self.count_variable[variable] -= 1
            self.queue_variable_by_counter(variable)
        assert self.count_constraint[constraint] == 0

    def queue_variable_by_counter(self, variable):
        c = self.count_variable[variable]
        if c == 0:
            if self.gf.variables[variable].cost == 0:
                value = ("solved", self.variable_feasible_value(variable))
            else:
                value = self.optimize_column_independently(variable)
            self.removed_variables.append((variable, value))
        elif c == 1 and self.gf.variables[variable].cost == 0:
            self.q_slack.append(variable)

    def queue_constraint_by_counter(self, constraint):
        c = self.count_constraint[constraint]
        if c == 0:
            if not is_empty_constraint_feasible(self.get_b(constraint), self.constraint_type(constraint)):
                raise Infeasible()
            self.constraints_marked_removed.append(constraint)
            return MEANINGFUL
        if c == 1:
            self.q_bound.append(constraint)
        return NO_CHANGE

    # -- rule: fixed variable (rule/fixed_variable.rs) -----------------------------------------------------
    def presolve_fixed_variable(self, variable):
        value = self.is_variable_fixed(variable)
        for constraint, coefficient in self.iter_active_column(variable):
            self.change_b(constraint, -coefficient * value)
        self.fixed_cost += self.gf.variables[variable].cost * value
        for constraint in [i for i, _ in self.iter_active_column(variable)]:
            self.count_variable[variable] -= 1
            self.count_constraint[constraint] -= 1
            self.queue_constraint_by_counter(constraint)
        self.removed_variables.append((variable, ("solved", value)))

    # -- rule: bound constraint (rule/bound_constraint.rs) ---------------------------------------------------
    def presolve_bound_constraint(self, constraint):
        (variable, coefficient), = self.iter_active_row(constraint)
        rhs, ctype = self.get_b(constraint), self.constraint_type(constraint)
        bound_value = rhs / coefficient
        positive = coefficient > 0
        if isinstance(ctype, tuple):
            bound1 = (rhs - ctype[1]) / coefficient
            changes = [(LOWER, bound1), (UPPER, bound_value)] if positive else [(LOWER, bound_value), (UPPER, bound1)]
        elif ctype == "E":
            changes = [(LOWER, bound_value), (UPPER, bound_value)]
        elif (ctype == "G") == positive:
            changes = [(LOWER, bound_value)]
        else:
            changes = [(UPPER, bound_value)]
        self.count_variable[variable] -= 1
        self.count_constraint[constraint] -= 1
        self.constraints_marked_removed.append(constraint)
        for direction, value in changes:
            change = self.update_bound(variable, direction, value)
            if change[0] == "new":
                self.after_bound_change(variable, direction, None)
            elif change[0] == "shift":
                self.after_bound_change(variable, direction, change[1])
        if self.variable_feasible_value(variable) is None:
            raise Infeasible()
        self.queue_variable_by_counter(variable)

    # -- rule: slack (rule/slack.rs) ---------------------------------------------------------------------------
    def compute_removed_variable_solution(self, constraint, variable, coefficient):
        constant = self.get_b(constraint) / coefficient
        coefficients = [(self.gf.from_active_to_original[j], other / coefficient)
                        for j, other in self.iter_active_row(constraint) if j != variable]
        return ("function", constant, coefficients)

    def presolve_slack(self, variable):
        (constraint, coefficient), = self.iter_active_column(variable)
        ctype = self.constraint_type(constraint)
        lower, upper = self.variable_bound(variable, LOWER), self.variable_bound(variable, UPPER)
        has = (lower is not None, upper is not None)
        positive = coefficient > 0
        kind = "R" if isinstance(ctype, tuple) else ctype
        # the variable can absorb the whole constraint: both disappear (slack.rs:41-63)
        absorbs = (has == (False, False)
                   or (kind == "G" and has == (True, False) and positive) or (kind == "L" and has == (False, True) and positive)
                   or (kind == "L" and has == (True, False) and not positive)
                   or (kind == "G" and has == (False, True) and not positive))
        if absorbs:
            solution = self.compute_removed_variable_solution(constraint, variable, coefficient)
            for other in [j for j, _ in self.iter_active_row(constraint)]:
                self.count_constraint[constraint] -= 1
                self.count_variable[other] -= 1
                if other != variable:
                    self.queue_variable_by_counter(other)
            self.removed_variables.append((variable, solution))
            self.constraints_marked_removed.append(constraint)
            return
        if has == (True, True):
            if kind == "E":
                new_type, bound = (("R", coefficient * (upper - lower)), lower) if positive else \
                                  (("R", coefficient * (lower - upper)), upper)
            elif kind == "R":
                new_type, bound = (("R", ctype[1] + coefficient * (upper - lower)), lower) if positive else \
                                  (("R", ctype[1] + coefficient * (lower - upper)), upper)
            elif kind == "L":
                new_type, bound = ("L", lower) if positive else ("L", upper)
            else:
                new_type, bound = ("G", upper) if positive else ("G", lower)
        elif has == (True, False):
            # (Less | Equal | Range, (Some, None), Positive) -> Less; (Equal | Greater | Range, (Some, None), Negative)
            # -> Greater; the remaining (Greater, Positive) / (Less, Negative) cases were absorbed above
            new_type, bound = ("L", lower) if positive else ("G", lower)
        else:
            new_type, bound = ("G", upper) if positive else ("L", upper)
        change = -coefficient * bound
        if kind in ("E", "R"):
            removed = self.compute_removed_variable_solution(constraint, variable, coefficient)
        else:
            removed = ("solved", bound)
        self.count_variable[variable] -= 1
        self.removed_variables.append((variable, removed))
        # update_activity_queues_if_needed, slack.rs:118-138
        none_lower, none_upper = lower is None, upper is None
        if (none_lower and positive) or (none_upper and not positive):
            self.count_activity[constraint][LOWER] -= 1
            if self.count_activity[constraint][LOWER] <= 1:
                self.q_activity.push((constraint, LOWER))
        if (none_upper and positive) or (none_lower and not positive):
            self.count_activity[constraint][UPPER] -= 1
            if self.count_activity[constraint][UPPER] <= 1:
                self.q_activity.push((constraint, UPPER))
        self.count_constraint[constraint] -= 1
        self.queue_constraint_by_counter(constraint)
        self.change_b(constraint, change)
        self.constraints[constraint] = new_type

    # -- rule: domain propagation (rule/domain_propagation.rs) ------------------------------------------------------
    def presolve_domain_propagation(self, constraint, direction):
        counter = self.count_activity[constraint][direction]
        missing = sum(1 for j, c in self.iter_active_row(constraint)
                      if self.variable_bound(j, _times(direction, c)) is None)
        assert missing == counter, (missing, counter)
        if counter == 0:
            return self.for_entire_constraint(constraint, direction)
        assert counter == 1
        return self.create_variable_bound(constraint, direction)

    def compute_activity_bound_if_needed(self, constraint, direction):
        if self.activity_bounds[constraint][direction] is None:
            self.activity_bounds[constraint][direction] = sum(
                (c * self.variable_bound(j, _times(direction, c)) for j, c in self.iter_active_row(constraint)),
                Fraction(0))
        return self.activity_bounds[constraint][direction]

    def constraint_update(self, constraint, bound_value, direction):
        """domain_propagation.rs:165-222: None | "remove" | ("replace", relation, rhs shift) | "set_to_bound\""""
        rhs, ctype = self.get_b(constraint), self.constraint_type(constraint)
        kind = "R" if isinstance(ctype, tuple) else ctype
        cmp = (rhs > bound_value) - (rhs < bound_value)
        if direction == LOWER:
            if cmp < 0 and kind in ("E", "R", "L"):
                raise Infeasible()
            if cmp == 0 and kind in ("E", "L"):
                return "set_to_bound"
            if kind == "G" and cmp <= 0:
                return "remove"
            if kind == "R" and cmp > 0:
                lower_bound = rhs - ctype[1]
                return None if bound_value < lower_bound else ("replace", "L", Fraction(0))
            if kind == "R" and cmp == 0:
                raise AssertionError("a zero range is an equality")
            return None
        if cmp > 0 and kind in ("E", "G"):
            raise Infeasible()
        if cmp == 0 and kind in ("E", "G"):
            return "set_to_bound"
        if kind == "L" and cmp >= 0:
            return "remove"
        if kind == "R" and cmp == 0:
            return ("replace", "G", -ctype[1])
        if kind == "R" and cmp > 0:
            lower_bound = rhs - ctype[1]
            if bound_value < lower_bound:
                raise Infeasible()
            if bound_value == lower_bound:
                return "set_to_bound"
            return ("replace", "G", -ctype[1])
        return None

    def constraint_part(self, constraint, bound, direction, made_change):
        update = self.constraint_update(constraint, bound, direction)
        if update is None:
            return False, True
        if update == "remove":
            result = (True, True)
        elif update == "set_to_bound":
            counters_to_update = []
            for variable, coefficient in self.iter_active_row(constraint):
                vdir = _times(direction, coefficient)
                value = self.variable_bound(variable, vdir)
                if (variable, vdir) in self.activity_var_bounds:
                    self.bounds[(variable, vdir)] = self.activity_var_bounds.pop((variable, vdir))
                change = self.update_bound(variable, 1 - vdir, value)
                if change[0] == "new":
                    counters_to_update.append((variable, 1 - vdir))
                assert self.is_variable_fixed(variable) is not None
                self.q_substitution.append(variable)
            for variable, d in counters_to_update:
                self.update_activity_counters(variable, d)
            result = (True, False)
        else:
            _, relation, shift = update
            self.constraints[constraint] = relation
            self.change_b(constraint, shift)
            result = (False, True)
        made_change[0] = MEANINGFUL
        return result

    def can_variable_rule_be_applied(self, constraint, direction):
        rhs, ctype = self.get_b(constraint), self.constraint_type(constraint)
        if ctype == "E":
            return rhs
        if isinstance(ctype, tuple):
            return rhs if direction == LOWER else rhs - ctype[1]
        if ctype == "L":
            return rhs if direction == LOWER else None
        return None if direction == LOWER else rhs

    def variable_part(self, constraint, rhs, activity_bound, direction, made_change):
        for variable, coefficient in self.iter_active_row(constraint):
            new_direction = _times(1 - direction, coefficient)
            used = self.variable_bound(variable, _times(direction, coefficient))
            residual = activity_bound - coefficient * used
            new_value = (rhs - residual) / coefficient
            change = self.update_activity_variable_bound(variable, new_direction, new_value)
            if change[0] == "new":
                self.after_bound_change(variable, new_direction, None)
                made_change[0] = MEANINGFUL
            elif change[0] == "shift":
                self.after_bound_change(variable, new_direction, change[1])
                if made_change[0] != MEANINGFUL:
                    made_change[0] = NOT_MEANINGFUL

    def for_entire_constraint(self, constraint, direction):
        made_change = [NO_CHANGE]
        activity_bound = self.compute_activity_bound_if_needed(constraint, direction)
        remove, apply_variable_part = self.constraint_part(constraint, activity_bound, direction, made_change)
        if apply_variable_part:
            rhs = self.can_variable_rule_be_applied(constraint, direction)
            if rhs is not None:
                self.variable_part(constraint, rhs, activity_bound, direction, made_change)
        if remove:
            self.remove_constraint_values(constraint)
            self.constraints_marked_removed.append(constraint)
        return made_change[0]

    def create_variable_bound(self, constraint, direction):
        rhs = self.can_variable_rule_be_applied(constraint, direction)
        if rhs is None:
            return NO_CHANGE
        total = Fraction(0)
        target = None
        for variable, coefficient in self.iter_active_row(constraint):
            bound = self.variable_bound(variable, _times(direction, coefficient))
            if bound is None:
                if target is None:
                    target = (variable, coefficient)
            else:
                total += coefficient * bound
        target_column, target_coefficient = target
        value = (rhs - total) / target_coefficient
        bound_direction = _times(1 - direction, target_coefficient)
        change = self.update_activity_variable_bound(target_column, bound_direction, value)
        if change[0] == "new":
            self.after_bound_change(target_column, bound_direction, None)
            return MEANINGFUL
        if change[0] == "shift":
            self.after_bound_change(target_column, bound_direction, change[1])
            return NOT_MEANINGFUL
        return NO_CHANGE


def compute_presolve_changes(gf):
    """general_form/mod.rs:378-402"""
    index = Index(gf)
    without_meaningful_change = 0
    while not index.queues_empty() and \
            without_meaningful_change < index.nr_variables_remaining() + index.nr_constraints_remaining():
        change = index.presolve_step()
        if change == MEANINGFUL:
            without_meaningful_change = 0
        elif change == NOT_MEANINGFUL:
            without_meaningful_change += 1
    return index.into_changes()
