"""CPU ORACLE (test infrastructure only -- never imported by the product path).

Exact restatement, in Python `fractions.Fraction` arithmetic, of relp's two-phase simplex hot
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module; `relp_b200/` must not.

Every function cites the reference file:line it follows (paths relative to the reference's
`src/algorithm/two_phase/` unless noted).  The arithmetic crate `relp-num 0.1.13` (RationalBig) is
not vendored in the reference tree; its semantics relied upon here are "exact, always normalised
rationals with exact ordering", which `fractions.Fraction` provides.

Parity pinning: checked against the reference's own golden fixtures in
`tests/test_oracle_golden.py` -- staged carries / bases / optima of `src/tests/problem_{1,2}.rs`,
`tableau/mod.rs:491-566`, `strategy/pivot_rule.rs:314-344`, `two_phase/test.rs`, the max-flow and
shortest-path examples.  Basis *sequences* and iteration counts are pinned by no reference test
("parity unpinned" for traces); they follow from the tie-break rules restated below.

Index conventions are the reference's: in phase one, columns [0, n_a) are the virtual artificial
columns and provider column j sits at n_a + j (`tableau/kind/artificial/partially.rs:52-80`).
"""
from __future__ import annotations

from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

ZERO = Fraction(0)
ONE = Fraction(1)

SparseCol = List[Tuple[int, Fraction]]


# --------------------------------------------------------------------------------------------
# Matrix providers (matrix_provider/mod.rs:37-134)
# --------------------------------------------------------------------------------------------
class MatrixProvider:
    """`MatrixProvider` trait (matrix_provider/mod.rs:37-134)."""

    #: `PartialInitialBasis` implemented? (phase_one.rs:66-79)
    has_partial_initial_basis = False
    #: `FullInitialBasis` implemented? (phase_one.rs:101-110)
    has_full_initial_basis = False

    def nr_rows(self) -> int:
        raise NotImplementedError

    def nr_columns(self) -> int:
        raise NotImplementedError

    def column(self, j: int) -> SparseCol:
        """Sparse column, sorted by row index."""
        raise NotImplementedError

    def cost_value(self, j: int) -> Fraction:
        raise NotImplementedError

    def right_hand_side(self) -> List[Fraction]:
        raise NotImplementedError

    def pivot_element_indices(self) -> List[Tuple[int, int]]:
        raise NotImplementedError


class ExplicitProvider(MatrixProvider):
    """A provider given by explicit columns; used for synthetic LPs and lazy-provider snapshots."""

    def __init__(self, m, columns, costs, rhs, pivots=None, full_basis=False):
        self.m = m
        self.columns = [[(int(i), Fraction(v)) for i, v in col] for col in columns]
        self.costs = [Fraction(c) for c in costs]
        self.rhs = [Fraction(v) for v in rhs]
        self.pivots = None if pivots is None else [(int(i), int(j)) for i, j in pivots]
        self.has_partial_initial_basis = pivots is not None
        self.has_full_initial_basis = bool(full_basis)

    def nr_rows(self):
        return self.m

    def nr_columns(self):
        return len(self.columns)

    def column(self, j):
        return self.columns[j]

    def cost_value(self, j):
        return self.costs[j]

    def right_hand_side(self):
        return list(self.rhs)

    def pivot_element_indices(self):
        return list(self.pivots)


class Variable:
    """`general_form::Variable` fields the provider uses (matrix_data.rs:331-353)."""

    def __init__(self, cost, upper_bound=None):
        self.cost = Fraction(cost)
        self.upper_bound = None if upper_bound is None else Fraction(upper_bound)


class MatrixData(MatrixProvider):
    """`MatrixData` (matrix_provider/matrix_data.rs:63-450).

    Row order: equality, range, upper (<=), lower (>=), variable-bound rows, range-slack-bound
    rows (matrix_data.rs:46-61,139-157).  Column order: structural, range slacks, <= slacks,
    >= slacks (coefficient -1), variable-bound slacks, range-bound slacks (:159-176,291-329).
    """

    has_partial_initial_basis = True

    def __init__(self, constraint_columns: Sequence[SparseCol], b, ranges, nr_eq, nr_range,
                 nr_upper, nr_lower, variables: Sequence[Variable]):
        self.constraint_columns = [[(int(i), Fraction(v)) for i, v in c] for c in constraint_columns]
        self.b = [Fraction(v) for v in b]
        self.ranges = [Fraction(v) for v in ranges]
        assert len(self.ranges) == nr_range
        self.variables = list(variables)
        assert len(self.variables) == len(self.constraint_columns)
        self.bound_to_var = [j for j, v in enumerate(self.variables) if v.upper_bound is not None]
        self.var_to_bound = {j: k for k, j in enumerate(self.bound_to_var)}
        nr_bounds = len(self.bound_to_var)
        counts_r = [nr_eq, nr_range, nr_upper, nr_lower, nr_bounds, nr_range]
        counts_c = [len(self.variables), nr_range, nr_upper, nr_lower, nr_bounds, nr_range]
        self.row_end = _cumsum(counts_r)
        self.col_end = _cumsum(counts_c)
        assert self.row_end[3] == len(self.b)
        self.nr_upper = nr_upper
        self.nr_range = nr_range

    def nr_rows(self):
        return self.row_end[5]

    def nr_columns(self):
        return self.col_end[5]

    def _column_type(self, j):
        # matrix_data.rs:186-207
        for t in range(6):
            if j < self.col_end[t]:
                return t, j - (self.col_end[t - 1] if t else 0)
        raise IndexError(j)

    def column(self, j):
        # matrix_data.rs:291-329
        t, k = self._column_type(j)
        if t == 0:
            col = list(self.constraint_columns[k])
            if k in self.var_to_bound:
                col.append((self.row_end[3] + self.var_to_bound[k], ONE))
            return col
        if t == 1:
            return [(self.row_end[0] + k, ONE), (self.row_end[4] + k, ONE)]
        if t == 2:
            return [(self.row_end[1] + k, ONE)]
        if t == 3:
            return [(self.row_end[2] + k, -ONE)]
        if t == 4:
            return [(self.row_end[3] + k, ONE)]
        return [(self.row_end[4] + k, ONE)]

    def cost_value(self, j):
        # matrix_data.rs:331-339 (None for slack columns == zero cost)
        t, k = self._column_type(j)
        return self.variables[k].cost if t == 0 else ZERO

    def right_hand_side(self):
        # matrix_data.rs:341-353
        return (list(self.b) + [self.variables[j].upper_bound for j in self.bound_to_var]
                + list(self.ranges))

    def pivot_element_indices(self):
        # matrix_data.rs:419-445
        out = [(self.row_end[1] + j, self.col_end[1] + j) for j in range(self.nr_upper)]
        out += [(self.row_end[3] + j, self.col_end[3] + j) for j in range(len(self.bound_to_var))]
        out += [(self.row_end[4] + j, self.col_end[4] + j) for j in range(self.nr_range)]
        return out

    def reconstruct_solution(self, values: Dict[int, Fraction]) -> Dict[int, Fraction]:
        # matrix_data.rs:402-411
        return {j: v for j, v in values.items() if j < len(self.variables)}


def _cumsum(xs):
    out, s = [], 0
    for x in xs:
        s += x
        out.append(s)
    return out


class RemoveRows(MatrixProvider):
    """`RemoveRows` wrapper (matrix_provider/filter/generic_wrapper.rs:52-285)."""

    def __init__(self, provider: MatrixProvider, rows_to_skip: List[int]):
        self.provider = provider
        self.rows_to_skip = list(rows_to_skip)
        assert self.rows_to_skip == sorted(set(self.rows_to_skip))
        self.has_partial_initial_basis = provider.has_partial_initial_basis
        skip = set(self.rows_to_skip)
        self._new_index = {}
        k = 0
        for i in range(provider.nr_rows()):
            if i in skip:
                continue
            self._new_index[i] = k
            k += 1

    def nr_rows(self):
        return self.provider.nr_rows() - len(self.rows_to_skip)

    def nr_columns(self):
        return self.provider.nr_columns()

    def column(self, j):
        # generic_wrapper.rs:229-233 -> remove_sparse_indices
        return [(self._new_index[i], v) for i, v in self.provider.column(j) if i in self._new_index]

    def cost_value(self, j):
        return self.provider.cost_value(j)

    def right_hand_side(self):
        rhs = self.provider.right_hand_side()
        return [v for i, v in enumerate(rhs) if i in self._new_index]

    def pivot_element_indices(self):
        return [(self._new_index[i], j) for i, j in self.provider.pivot_element_indices()
                if i in self._new_index]


# --------------------------------------------------------------------------------------------
# Carry with explicit B^-1 rows (carry/mod.rs:46-66, carry/basis_inverse_rows.rs)
# --------------------------------------------------------------------------------------------
def sparse_dot_dense(dense: Sequence[Fraction], col: SparseCol) -> Fraction:
    """`DenseVector::sparse_inner_product` (data/linear_algebra/vector/dense.rs:101-112)."""
    s = ZERO
    for i, v in col:
        d = dense[i]
        if d:
            s += d * v
    return s


def sparse_dot_row(row: Dict[int, Fraction], col: SparseCol) -> Fraction:
    """`SparseVector::sparse_inner_product` (data/linear_algebra/vector/sparse.rs:105-128)."""
    s = ZERO
    for i, v in col:
        r = row.get(i)
        if r is not None:
            s += r * v
    return s


class Carry:
    """`Carry<F, BasisInverseRows<F>>` (carry/mod.rs:46-66): -obj, -pi, b, basis_indices, B^-1.

    B^-1 is kept as `m` sparse rows (dict column -> value), like `BasisInverseRows`
    (basis_inverse_rows.rs:23-30).  The LU representation of the reference
    (carry/lower_upper/) yields the same exact values and is not restated.
    """

    def __init__(self, minus_objective, minus_pi, b, basis_indices, rows):
        self.minus_objective = Fraction(minus_objective)
        self.minus_pi = [Fraction(v) for v in minus_pi]
        self.b = [Fraction(v) for v in b]
        self.basis_indices = list(basis_indices)
        self.rows: List[Dict[int, Fraction]] = rows

    @property
    def m(self):
        return len(self.b)

    # ---- constructors -------------------------------------------------------------------
    @staticmethod
    def identity_rows(m):
        return [{i: ONE} for i in range(m)]

    @classmethod
    def create_for_fully_artificial(cls, b):
        # carry/mod.rs:374-395
        m = len(b)
        return cls(-sum(b, ZERO), [-ONE] * m, b, list(range(m)), cls.identity_rows(m))

    @classmethod
    def create_for_partially_artificial(cls, artificial_rows, free_basis_values, b, basis_indices):
        # carry/mod.rs:397-442
        m = len(b)
        assert len(artificial_rows) + len(free_basis_values) == m
        objective = sum((b[i] for i in artificial_rows), ZERO)
        art = set(artificial_rows)
        minus_pi = [-ONE if i in art else ZERO for i in range(m)]
        return cls(-objective, minus_pi, b, basis_indices, cls.identity_rows(m))

    @staticmethod
    def _minus_pi_from_basis(rows, provider, basis):
        # create_minus_pi_from_artificial, carry/mod.rs:226-260: pi_j = sum_i B^-1[i][j] c_B(i)
        m = len(rows)
        pi = [ZERO] * m
        for i, row in enumerate(rows):
            c = provider.cost_value(basis[i])
            if c:
                for j, v in row.items():
                    pi[j] += v * c
        return [-v for v in pi]

    @staticmethod
    def _minus_obj_from_basis(provider, basis, b):
        # create_minus_obj_from_artificial, carry/mod.rs:270-283
        return -sum((b[i] * provider.cost_value(basis[i]) for i in range(len(b))), ZERO)

    @classmethod
    def from_artificial(cls, artificial: "Carry", provider, nr_artificial):
        # carry/mod.rs:499-525
        basis = [j - nr_artificial for j in artificial.basis_indices]
        assert all(j >= 0 for j in basis)
        minus_pi = cls._minus_pi_from_basis(artificial.rows, provider, basis)
        minus_obj = cls._minus_obj_from_basis(provider, basis, artificial.b)
        return cls(minus_obj, minus_pi, artificial.b, basis, artificial.rows)

    @classmethod
    def from_artificial_remove_rows(cls, artificial: "Carry", rows_removed: RemoveRows, nr_artificial):
        # carry/mod.rs:673-712 with BasisInverseRows::remove_basis_part
        # (basis_inverse_rows.rs:212-229): delete rows and the same-index columns.
        skip = set(rows_removed.rows_to_skip)
        keep = [i for i in range(artificial.m) if i not in skip]
        new_index = {i: k for k, i in enumerate(keep)}
        basis = [artificial.basis_indices[i] - nr_artificial for i in keep]
        assert all(j >= 0 for j in basis)
        rows = [{new_index[k]: v for k, v in artificial.rows[i].items() if k in new_index}
                for i in keep]
        b = [artificial.b[i] for i in keep]
        minus_pi = cls._minus_pi_from_basis(rows, rows_removed, basis)
        minus_obj = cls._minus_obj_from_basis(rows_removed, basis, b)
        return cls(minus_obj, minus_pi, b, basis, rows)

    @classmethod
    def from_basis(cls, basis, provider):
        """carry/mod.rs:444-478: `basis[i]` is the column basic in row i.  `BI::invert(columns)`
        (basis_inverse_rows.rs:104-129: every identity column through the inverted basis) restated as an exact
        Gauss-Jordan inversion; b = B^-1 rhs (:456-466), then -obj and -pi from the basic costs."""
        m = provider.nr_rows()
        assert len(basis) == m
        # augmented [B | I], B[:, i] = column basis[i]
        M = [[ZERO] * (2 * m) for _ in range(m)]
        for i, j in enumerate(basis):
            for r, v in provider.column(j):
                M[r][i] = Fraction(v)
        for r in range(m):
            M[r][m + r] = ONE
        for c in range(m):
            piv = next((r for r in range(c, m) if M[r][c] != 0), None)
            assert piv is not None, "singular basis"
            M[c], M[piv] = M[piv], M[c]
            inv = ONE / M[c][c]
            M[c] = [v * inv for v in M[c]]
            for r in range(m):
                if r != c and M[r][c] != 0:
                    f = M[r][c]
                    M[r] = [a - f * x for a, x in zip(M[r], M[c])]
        rows = [{k: M[i][m + k] for k in range(m) if M[i][m + k] != 0} for i in range(m)]
        rhs = provider.right_hand_side()
        b = [sum((v * rhs[k] for k, v in rows[i].items()), ZERO) for i in range(m)]
        minus_obj = cls._minus_obj_from_basis(provider, basis, b)
        minus_pi = cls._minus_pi_from_basis(rows, provider, basis)
        return cls(minus_obj, minus_pi, b, list(basis), rows)

    @classmethod
    def from_basis_pivots(cls, pivots, provider):
        # carry/mod.rs:480-497: sort by row, then from_basis
        elements = sorted(pivots, key=lambda rc: rc[0])
        return cls.from_basis([c for _, c in elements], provider)

    # ---- queries ------------------------------------------------------------------------
    def cost_difference(self, column: SparseCol) -> Fraction:
        # carry/mod.rs:606-611
        return sparse_dot_dense(self.minus_pi, column)

    def generate_column(self, column: SparseCol) -> Dict[int, Fraction]:
        # carry/mod.rs:613-621 -> left_multiply_by_basis_inverse (basis_inverse_rows.rs:147-160)
        out = {}
        for i, row in enumerate(self.rows):
            v = sparse_dot_row(row, column)
            if v:
                out[i] = v
        return out

    def generate_element(self, i: int, column: SparseCol) -> Optional[Fraction]:
        # basis_inverse_rows.rs:179-195
        v = sparse_dot_row(self.rows[i], column)
        return v if v else None

    def current_bfs(self) -> List[Tuple[int, Fraction]]:
        # carry/mod.rs:636-645
        out = [(self.basis_indices[i], v) for i, v in enumerate(self.b) if v]
        out.sort(key=lambda t: t[0])
        return out

    def get_objective_function_value(self) -> Fraction:
        return -self.minus_objective

    # ---- pivot --------------------------------------------------------------------------
    def change_basis(self, p: int, q: int, column: Dict[int, Fraction], relative_cost: Fraction):
        """`Carry::change_basis` (carry/mod.rs:561-604).  Returns BasisChangeComputationInfo."""
        # work_vector = column^T B^-1 (right_multiply_by_basis_inverse, basis_inverse_rows.rs:162-177)
        work: Dict[int, Fraction] = {}
        for i, a in column.items():
            for k, v in self.rows[i].items():
                work[k] = work.get(k, ZERO) + a * v
        work = {k: v for k, v in work.items() if v}

        # update_b (carry/mod.rs:295-325)
        pivot_value = column[p]
        self.b[p] /= pivot_value
        bp = self.b[p]
        for i, a in column.items():
            if i != p:
                self.b[i] -= a * bp

        leaving = self.basis_indices[p]
        self.basis_indices[p] = q

        # BasisInverseRows::change_basis (basis_inverse_rows.rs:91-99): normalise + row-reduce
        rowp = self.rows[p]
        if pivot_value != 1:
            for k in rowp:
                rowp[k] /= pivot_value
        for i, a in column.items():
            if i == p:
                continue
            row = self.rows[i]
            for k, v in rowp.items():
                nv = row.get(k, ZERO) - a * v
                if nv:
                    row[k] = nv
                else:
                    row.pop(k, None)

        # update_minus_pi_and_obj (carry/mod.rs:338-349)
        for k, v in rowp.items():
            self.minus_pi[k] -= relative_cost * v
        self.minus_objective -= relative_cost * self.b[p]

        return dict(pivot_row_index=p, pivot_column_index=q, leaving_column_index=leaving,
                    column_before_change=column, work_vector=work, basis_inverse_row=dict(rowp))


# --------------------------------------------------------------------------------------------
# Tableau + kinds (tableau/mod.rs, tableau/kind/)
# --------------------------------------------------------------------------------------------
class Tableau:
    """`Tableau<IM, K>` (tableau/mod.rs:25-39) for K in {Fully, Partially, NonArtificial}.

    `column_to_row` is the artificial-column -> row map of `Partially`
    (partially.rs:17-21); `None` means `NonArtificial` (non_artificial.rs:18-71).
    """

    def __init__(self, provider, carry: Carry, basis_columns, column_to_row=None):
        self.provider = provider
        self.im = carry
        self.basis_columns = set(basis_columns)
        self.column_to_row = column_to_row

    # ---- kind ---------------------------------------------------------------------------
    def nr_artificial_variables(self):
        return 0 if self.column_to_row is None else len(self.column_to_row)

    def start_index(self):
        # pivot_rule.rs:57-80
        return self.nr_artificial_variables()

    def nr_rows(self):
        return self.provider.nr_rows()

    def nr_columns(self):
        return self.nr_artificial_variables() + self.provider.nr_columns()

    def initial_cost_value(self, j):
        # fully.rs:29 / partially.rs:52 / non_artificial.rs:42
        if self.column_to_row is None:
            return self.provider.cost_value(j)
        return ONE if j < len(self.column_to_row) else ZERO

    def original_column(self, j) -> SparseCol:
        # fully.rs:37 / partially.rs:62 / non_artificial.rs:48
        na = self.nr_artificial_variables()
        if j < na:
            return [(self.column_to_row[j], ONE)]
        return self.provider.column(j - na)

    # ---- constructors -------------------------------------------------------------------
    @classmethod
    def new_fully(cls, provider):
        # fully.rs:82-97
        m = provider.nr_rows()
        carry = Carry.create_for_fully_artificial(provider.right_hand_side())
        return cls(provider, carry, range(m), list(range(m)))

    @classmethod
    def new_partially(cls, provider):
        # partially.rs:125-205
        m = provider.nr_rows()
        real = provider.pivot_element_indices()
        assert real == sorted(real, key=lambda rc: rc[0])
        real_rows = {r for r, _ in real}
        artificial = [i for i in range(m) if i not in real_rows]
        nr_artificial = len(artificial)
        art_col_of_row = {r: k for k, r in enumerate(artificial)}
        real_col_of_row = {r: c for r, c in real}
        basis_indices = [art_col_of_row[i] if i in art_col_of_row else nr_artificial + real_col_of_row[i]
                         for i in range(m)]
        carry = Carry.create_for_partially_artificial(artificial, real, provider.right_hand_side(),
                                                      basis_indices)
        return cls(provider, carry, basis_indices, artificial)

    @classmethod
    def from_artificial(cls, carry, nr_artificial, basis, provider):
        # non_artificial.rs:151-172
        im = Carry.from_artificial(carry, provider, nr_artificial)
        return cls(provider, im, {j - nr_artificial for j in basis}, None)

    @classmethod
    def from_artificial_removing_rows(cls, carry, nr_artificial, basis, rows_removed: RemoveRows):
        # non_artificial.rs:191-226
        basis = set(basis)
        for row in rows_removed.rows_to_skip:
            basis.remove(carry.basis_indices[row])
        im = Carry.from_artificial_remove_rows(carry, rows_removed, nr_artificial)
        return cls(rows_removed, im, {j - nr_artificial for j in basis}, None)

    # ---- operations ---------------------------------------------------------------------
    def is_in_basis(self, j):
        return j in self.basis_columns

    def relative_cost(self, j) -> Fraction:
        # tableau/mod.rs:106-112
        return self.im.cost_difference(self.original_column(j)) + self.initial_cost_value(j)

    def generate_column(self, j):
        # tableau/mod.rs:126-130
        return self.im.generate_column(self.original_column(j))

    def generate_element(self, i, j):
        # tableau/mod.rs:133-138
        return self.im.generate_element(i, self.original_column(j))

    def variable_value(self, column):
        # tableau/mod.rs:163-175
        if column in self.basis_columns:
            row = self.im.basis_indices.index(column)
            return self.im.b[row]
        return ZERO

    def objective_function_value(self):
        return self.im.get_objective_function_value()

    def current_bfs(self):
        return self.im.current_bfs()

    def select_primal_pivot_row(self, column: Dict[int, Fraction]) -> Optional[int]:
        """Ratio test with Bland tie-break on the leaving column (tableau/mod.rs:287-313)."""
        best = None  # (row, ratio, leaving_column)
        for row in sorted(column):
            xij = column[row]
            if xij > 0:
                ratio = self.im.b[row] / xij
                leaving = self.im.basis_indices[row]
                if best is None:
                    best = (row, ratio, leaving)
                elif ratio == best[1] and leaving < best[2]:
                    best = (row, best[1], leaving)
                elif ratio < best[1]:
                    best = (row, ratio, leaving)
        return None if best is None else best[0]

    def bring_into_basis(self, q, p, column, cost):
        # tableau/mod.rs:48-64
        info = self.im.change_basis(p, q, column, cost)
        self.basis_columns.remove(info["leaving_column_index"])
        self.basis_columns.add(q)
        return info

    # ---- artificial-only ----------------------------------------------------------------
    def has_artificial_in_basis(self):
        na = self.nr_artificial_variables()
        return any(c < na for c in self.basis_columns)

    def artificial_basis_columns(self):
        na = self.nr_artificial_variables()
        return [(i, j) for i, j in enumerate(self.im.basis_indices) if j < na]


# --------------------------------------------------------------------------------------------
# Pivot rules (strategy/pivot_rule.rs)
# --------------------------------------------------------------------------------------------
class FirstProfitable:
    """pivot_rule.rs:86-109."""
    name = "first_profitable"

    def __init__(self, tableau):
        pass

    def select_primal_pivot_column(self, t: Tableau):
        for j in range(t.start_index(), t.nr_columns()):
            if not t.is_in_basis(j):
                c = t.relative_cost(j)
                if c < 0:
                    return j, c
        return None

    def after_basis_update(self, info, t):
        pass


class FirstProfitableWithMemory:
    """pivot_rule.rs:113-150."""
    name = "first_profitable_with_memory"

    def __init__(self, tableau):
        self.last_selected = None

    def select_primal_pivot_column(self, t: Tableau):
        def find(lo, hi):
            for j in range(lo, hi):
                if not t.is_in_basis(j):
                    c = t.relative_cost(j)
                    if c < 0:
                        return j, c
            return None

        if self.last_selected is None:
            potential = find(t.start_index(), t.nr_columns())
        else:
            potential = find(self.last_selected + 1, t.nr_columns()) or find(t.start_index(),
                                                                             self.last_selected)
        self.last_selected = None if potential is None else potential[0]
        return potential

    def after_basis_update(self, info, t):
        pass


class SteepestDescentAlongVariable:
    """Dantzig: most negative relative cost, strict `<` so the lowest index wins ties
    (pivot_rule.rs:153-187)."""
    name = "dantzig"

    def __init__(self, tableau):
        pass

    def select_primal_pivot_column(self, t: Tableau):
        smallest = None
        for j in range(t.start_index(), t.nr_columns()):
            if t.is_in_basis(j):
                continue
            c = t.relative_cost(j)
            if c < 0 and (smallest is None or c < smallest[1]):
                smallest = (j, c)
        return smallest

    def after_basis_update(self, info, t):
        pass


class SteepestDescentAlongObjective:
    """Goldfarb-Reid steepest edge (pivot_rule.rs:190-305).

    `max_by_key` returns the LAST maximal element, so the highest index wins ties
    (pivot_rule.rs:233-240).
    """
    name = "steepest_edge"

    def __init__(self, t: Tableau, check=False):
        # pivot_rule.rs:202-219
        self.check = check
        self.gamma: List[Optional[Fraction]] = [
            initial_gamma(j, t) if (j >= t.start_index() and not t.is_in_basis(j)) else None
            for j in range(t.nr_columns())
        ]

    def select_primal_pivot_column(self, t: Tableau):
        best = None
        best_key = None
        for j in range(t.start_index(), t.nr_columns()):
            if t.is_in_basis(j):
                continue
            c = t.relative_cost(j)
            if c < 0:
                key = c * c / self.gamma[j]
                if best is None or key >= best_key:
                    best, best_key = (j, c), key
        return best

    def after_basis_update(self, info, t: Tableau):
        # pivot_rule.rs:243-296
        q = info["pivot_column_index"]
        self.gamma[q] = None
        col = info["column_before_change"]
        gamma_q = ONE + sum((v * v for v in col.values()), ZERO)
        row_p = info["basis_inverse_row"]
        work = info["work_vector"]
        for j in range(t.start_index(), len(self.gamma)):
            g = self.gamma[j]
            if g is None:
                continue
            original = t.original_column(j)
            alpha_j_bar = sparse_dot_row(row_p, original)
            if alpha_j_bar:
                sq = alpha_j_bar * alpha_j_bar
                inner = sparse_dot_row(work, original)
                if inner:
                    g -= 2 * alpha_j_bar * inner
                g += sq * gamma_q
                alternative = ONE + sq
            else:
                alternative = ONE
            if g < alternative:
                g = alternative
            self.gamma[j] = g
            if self.check:
                assert g == initial_gamma(j, t), (j, g, initial_gamma(j, t))
        w_p = col[info["pivot_row_index"]]
        self.gamma[info["leaving_column_index"]] = gamma_q / (w_p * w_p)


def initial_gamma(j, t: Tableau) -> Fraction:
    # pivot_rule.rs:299-305
    return ONE + sum((v * v for v in t.generate_column(j).values()), ZERO)


PIVOT_RULES = {
    "first_profitable": FirstProfitable,
    "first_profitable_with_memory": FirstProfitableWithMemory,
    "dantzig": SteepestDescentAlongVariable,
    "steepest_edge": SteepestDescentAlongObjective,
}


# --------------------------------------------------------------------------------------------
# The loops (phase_one.rs, phase_two.rs, two_phase/mod.rs)
# --------------------------------------------------------------------------------------------
class Trace:
    """Per-pivot record used for GPU parity: (phase, entering, row, leaving, objective).

    Column indices are in the reference's index space of the respective phase.  `phase` is
    1, 2 or 0 for zero-level pivots of `remove_artificial_basis_variables`.
    """

    def __init__(self, limit=None):
        self.pivots: List[Tuple[int, int, int, int, Fraction]] = []
        self.limit = limit

    def record(self, phase, q, p, leaving, objective):
        self.pivots.append((phase, q, p, leaving, objective))
        if self.limit is not None and len(self.pivots) >= self.limit:
            raise PivotLimit()


class PivotLimit(Exception):
    pass


def phase_one_primal(tableau: Tableau, rule_cls, trace: Optional[Trace] = None):
    """`phase_one::primal` (phase_one.rs:123-179).

    Returns ("feasible", rank_rows_to_remove, nr_artificial, carry, basis_set) or ("infeasible",).
    """
    rule = rule_cls(tableau)
    while True:
        sel = rule.select_primal_pivot_column(tableau)
        if sel is None:
            break
        q, cost = sel
        column = tableau.generate_column(q)
        p = tableau.select_primal_pivot_row(column)
        if p is None:
            raise RuntimeError("Artificial cost can not be unbounded.")  # phase_one.rs:151
        info = tableau.bring_into_basis(q, p, column, cost)
        rule.after_basis_update(info, tableau)
        if trace is not None:
            trace.record(1, q, p, info["leaving_column_index"], tableau.objective_function_value())

    if tableau.objective_function_value() != 0:
        return ("infeasible",)
    rows_to_remove: List[int] = []
    if tableau.has_artificial_in_basis():
        rows_to_remove = remove_artificial_basis_variables(tableau, trace)
    return ("feasible", rows_to_remove, tableau.nr_artificial_variables(), tableau.im,
            set(tableau.basis_columns))


def remove_artificial_basis_variables(tableau: Tableau, trace: Optional[Trace] = None) -> List[int]:
    """phase_one.rs:232-278."""
    rows_to_remove = []
    for pivot_row, artificial in tableau.artificial_basis_columns():
        constraint_value = tableau.variable_value(artificial)
        found = None
        for j in range(tableau.nr_artificial_variables(), tableau.nr_columns()):
            if tableau.is_in_basis(j):
                continue
            cost = tableau.relative_cost(j)
            if constraint_value:
                if cost != 0:
                    continue
                e = tableau.generate_element(pivot_row, j)
                if e is not None and e > 0:
                    found = (j, cost)
                    break
            else:
                e = tableau.generate_element(pivot_row, j)
                if e is not None:
                    found = (j, cost)
                    break
        if found is not None:
            q, cost = found
            column = tableau.generate_column(q)
            info = tableau.bring_into_basis(q, pivot_row, column, cost)
            if trace is not None:
                trace.record(0, q, pivot_row, info["leaving_column_index"],
                             tableau.objective_function_value())
        else:
            rows_to_remove.append(pivot_row)
    return rows_to_remove


def phase_two_primal(tableau: Tableau, rule_cls, trace: Optional[Trace] = None):
    """`phase_two::primal` (phase_two.rs:22-58).  Returns ("optimal", bfs) or ("unbounded",)."""
    rule = rule_cls(tableau)
    while True:
        sel = rule.select_primal_pivot_column(tableau)
        if sel is None:
            return ("optimal", tableau.current_bfs())
        q, cost = sel
        column = tableau.generate_column(q)
        p = tableau.select_primal_pivot_row(column)
        if p is None:
            return ("unbounded",)
        info = tableau.bring_into_basis(q, p, column, cost)
        rule.after_basis_update(info, tableau)
        if trace is not None:
            trace.record(2, q, p, info["leaving_column_index"], tableau.objective_function_value())


class Result:
    def __init__(self, status, bfs=None, objective=None, tableau=None, nr_artificial=0,
                 rows_removed=()):
        self.status = status            # "optimal" | "unbounded" | "infeasible"
        self.bfs = bfs                  # sorted [(column, value)] with value != 0
        self.objective = objective
        self.tableau = tableau
        self.nr_artificial = nr_artificial
        self.rows_removed = list(rows_removed)


def solve_relaxation(provider: MatrixProvider, rule="steepest_edge",
                     trace: Optional[Trace] = None) -> Result:
    """`SolveRelaxation::solve_relaxation` (algorithm/mod.rs:17-36, two_phase/mod.rs:25-109).

    The reference hard-codes `SteepestDescentAlongObjective` (two_phase/mod.rs:57,68,107); `rule`
    lets the tests drive the other `PivotRule`s through the same loops.
    """
    rule_cls = PIVOT_RULES[rule] if isinstance(rule, str) else rule
    if provider.has_full_initial_basis:
        # two_phase/mod.rs:80-109
        pivots = provider.pivot_element_indices()
        im = Carry.from_basis_pivots(pivots, provider)
        tableau = Tableau(provider, im, [c for _, c in pivots], None)
        nr_a, removed = 0, []
    else:
        # phase_one.rs:41-60 (Fully) / :82-100 (Partially)
        art = (Tableau.new_partially(provider) if provider.has_partial_initial_basis
               else Tableau.new_fully(provider))
        res = phase_one_primal(art, rule_cls, trace)
        if res[0] == "infeasible":
            return Result("infeasible")
        _, removed, nr_a, carry, basis = res
        if removed:
            # two_phase/mod.rs:47-58
            tableau = Tableau.from_artificial_removing_rows(carry, nr_a, basis,
                                                            RemoveRows(provider, removed))
        else:
            tableau = Tableau.from_artificial(carry, nr_a, basis, provider)
    out = phase_two_primal(tableau, rule_cls, trace)
    if out[0] == "unbounded":
        return Result("unbounded", tableau=tableau, nr_artificial=nr_a, rows_removed=removed)
    return Result("optimal", bfs=out[1], objective=tableau.objective_function_value(),
                  tableau=tableau, nr_artificial=nr_a, rows_removed=removed)


# --------------------------------------------------------------------------------------------
# Network providers of the reference's examples
# --------------------------------------------------------------------------------------------
def incidence_matrix(adjacency_columns: Sequence[Sequence[Tuple[int, Fraction]]], removed):
    """`IncidenceMatrix::new` (data/linear_program/network/representation.rs:34-84).

    `adjacency_columns[from]` lists `(to, value)`.  Returns (edge columns, edge values);
    Incoming = +1, Outgoing = -1 (representation.rs:159-175).
    """
    removed = sorted(removed)
    nr_vertices = len(adjacency_columns)
    shift = {}
    k = 0
    for v in range(nr_vertices):
        if v in removed:
            k += 1
            continue
        shift[v] = v - k
    edges, values = [], []
    for frm, outgoing in enumerate(adjacency_columns):
        for to, value in outgoing:
            assert to != frm
            col = []
            if frm in shift:
                col.append((shift[frm], -ONE))
            if to in shift:
                col.append((shift[to], ONE))
            col.sort(key=lambda t: t[0])
            edges.append(col)
            values.append(Fraction(value))
    return edges, values


class MaxFlowPrimal(MatrixProvider):
    """`Primal<F>` of examples/max_flow.rs:31-223 (MatrixProvider + PartialInitialBasis)."""

    has_partial_initial_basis = True

    def __init__(self, adjacency_columns, s, t):
        self.nr_vertices = len(adjacency_columns)
        before = sum(len(adjacency_columns[v]) for v in range(s))
        self.s_arc_range = range(before, before + len(adjacency_columns[s]))
        self.edges, self.capacity = incidence_matrix(adjacency_columns, [s, t])

    def nr_edges(self):
        return len(self.edges)

    def nr_constraints(self):
        return self.nr_vertices - 2

    def nr_rows(self):
        return self.nr_constraints() + self.nr_edges()

    def nr_columns(self):
        return 2 * self.nr_edges()

    def column(self, j):
        # max_flow.rs:149-163
        e = self.nr_edges()
        if j < e:
            return list(self.edges[j]) + [(self.nr_constraints() + j, ONE)]
        return [(self.nr_constraints() + j - e, ONE)]

    def cost_value(self, j):
        return -ONE if j in self.s_arc_range else ZERO

    def right_hand_side(self):
        return [ZERO] * self.nr_constraints() + list(self.capacity)

    def pivot_element_indices(self):
        # max_flow.rs:214-216
        return [(j + self.nr_constraints(), self.nr_edges() + j) for j in range(self.nr_edges())]


class ShortestPathPrimal(MatrixProvider):
    """`Primal<F>` of examples/shortest_path.rs:35-119 (MatrixProvider only => `Fully`)."""

    def __init__(self, adjacency_columns, s, t):
        self.nr_vertices = len(adjacency_columns)
        self.s, self.t = s, t
        self.edges, self.cost = incidence_matrix(adjacency_columns, [s])

    def nr_rows(self):
        return self.nr_vertices - 1

    def nr_columns(self):
        return len(self.edges)

    def column(self, j):
        return list(self.edges[j])

    def cost_value(self, j):
        return self.cost[j]

    def right_hand_side(self):
        b = [ZERO] * self.nr_rows()
        b[self.t if self.t < self.s else self.t - 1] = ONE
        return b


def adjacency_from_rows(rows: Sequence[Sequence[int]]):
    """`ColumnMajor::from_test_data` on an adjacency matrix written as rows ("from is top, to is
    on the right", examples/max_flow.rs:266-273): returns adjacency_columns[from] = [(to, v)]."""
    n = len(rows[0])
    cols = [[] for _ in range(n)]
    for to, row in enumerate(rows):
        for frm, v in enumerate(row):
            if v != 0:
                cols[frm].append((to, Fraction(v)))
    return cols


def columns_from_rows(rows: Sequence[Sequence], nr_columns: int) -> List[SparseCol]:
    """`ColumnMajor::from_test_data(rows, nr_columns)` (data/linear_algebra/matrix.rs): dense rows
    -> sparse columns."""
    cols = [[] for _ in range(nr_columns)]
    for i, row in enumerate(rows):
        assert len(row) == nr_columns
        for j, v in enumerate(row):
            if v != 0:
                cols[j].append((i, Fraction(v)))
    return cols
