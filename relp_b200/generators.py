"""Synthetic workloads of BASELINE.json (recipes frozen in SURVEY.md section 8d / DESIGN.md).

All generators are deterministic in `seed` (numpy PCG64) and return an `IntegerProblem` whose
provider layout is that of the reference's `MatrixData` for an all-`<=` problem: structural columns
first, then one +1 slack per row (matrix_data.rs:291-329), slack basis = `pivot_element_indices`.
"""
import numpy as np

from .solver import IntegerProblem


def bounded_lp(m, n_struct, k_bounding=90, nnz_per_col=8, dense=False, seed=0, b_big=None,
               coupling=1, full_initial_basis=True, dense_block=None):
    """Bounded integer LP  min c x, A x <= b, x >= 0  with b > 0 (slack basis feasible).

    Rows 0..K-1 are 'bounding rows': structural column j carries an entry in {1..100} in its group
    row j mod K and in `coupling` further random bounding rows; b in [1000, 10000].  Every bounding
    row can bind, so up to K structural columns enter the basis and |det B| grows by ~6 bits per
    entered column -- this is what sweeps the limb widths.  Rows >= K can never bind
    (b = b_big > nnz * 100 * 10000), their slacks stay basic, which bounds |det B| (SURVEY.md hard
    part 3) and keeps the exact solve well defined at 16 limbs.  Costs uniform in {-100..-1}.

    dense=False: rows >= K hold exactly `nnz_per_col` non-zeros per column in {-100..100}\\{0}
    (config 4: m=4096, n_struct=8192); dense=True: every entry of rows >= K is drawn from
    {-100..100} (config 5: m=16384, n_struct=32768).
    """
    rng = np.random.default_rng(seed)
    K = min(k_bounding, m)
    rest = m - K
    if b_big is None:
        b_big = 10 ** 11 if dense else 10 ** 7
    if dense_block is None:
        dense_block = dense and m * n_struct > 2 ** 22    # large dense problems: 1 byte per coefficient
    # bounding block: (1 + coupling) candidate entries per column, duplicates collapse
    grp = (np.arange(n_struct, dtype=np.int64) % K)[:, None]
    others = rng.integers(0, K, size=(n_struct, coupling), dtype=np.int64)
    top_rows = np.concatenate([grp, others], axis=1)
    top_vals = rng.integers(1, 101, size=top_rows.shape, dtype=np.int64)
    order = np.argsort(top_rows, axis=1, kind="stable")
    top_rows = np.take_along_axis(top_rows, order, axis=1)
    top_vals = np.take_along_axis(top_vals, order, axis=1)
    dup = np.zeros_like(top_rows, dtype=bool)
    dup[:, 1:] = top_rows[:, 1:] == top_rows[:, :-1]
    if dense and dense_block:
        # structural columns as a dense int8 block (column-major), slacks + costs + rhs as usual
        block = np.zeros((n_struct, m), dtype=np.int8)
        block[:, K:] = rng.integers(-100, 101, size=(n_struct, rest), dtype=np.int8)
        rows_i = np.arange(n_struct)[:, None]
        # later duplicates overwrite earlier ones exactly like the CSC path keeps the first: write in
        # reverse column order of the sorted pairs so the first occurrence wins
        for c in range(top_rows.shape[1] - 1, -1, -1):
            sel = ~dup[:, c]
            block[rows_i[sel, 0], top_rows[sel, c]] = top_vals[sel, c].astype(np.int8)
        n = n_struct + m
        colptr = np.concatenate([np.zeros(n_struct + 1, dtype=np.int64), np.arange(1, m + 1, dtype=np.int64)])
        rowidx = np.arange(m, dtype=np.int32)
        vals = np.ones(m, dtype=np.int64)
        cost = np.concatenate([-rng.integers(1, 101, size=n_struct, dtype=np.int64),
                               np.zeros(m, dtype=np.int64)])
        rhs = np.concatenate([rng.integers(1000, 10001, size=K, dtype=np.int64),
                              np.full(rest, b_big, dtype=np.int64)])
        pivots = [(i, n_struct + i) for i in range(m)]
        prob = IntegerProblem(m, n, colptr, rowidx, vals, cost, rhs, pivots, full_initial_basis)
        prob.set_dense_block(block)
        return prob
    if dense:
        low_rows = np.tile(np.arange(K, m, dtype=np.int64), (n_struct, 1))
        low_vals = rng.integers(-100, 101, size=(n_struct, rest), dtype=np.int8).astype(np.int64)
        low_keep = low_vals != 0
    else:
        k = min(nnz_per_col, rest)
        # k distinct rows per column: argpartition of random keys
        keys = rng.random((n_struct, rest))
        low_rows = np.sort(np.argpartition(keys, k - 1, axis=1)[:, :k], axis=1).astype(np.int64) + K
        low_vals = (rng.integers(1, 101, size=(n_struct, k), dtype=np.int64)
                    * (rng.integers(0, 2, size=(n_struct, k), dtype=np.int64) * 2 - 1))
        low_keep = np.ones_like(low_vals, dtype=bool)
    rows_all = np.concatenate([top_rows, low_rows], axis=1)
    vals_all = np.concatenate([top_vals, low_vals], axis=1)
    keep = np.concatenate([~dup, low_keep], axis=1)
    counts = keep.sum(axis=1)
    colptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rowidx = rows_all[keep].astype(np.int32)
    vals = vals_all[keep]
    # slack columns
    n = n_struct + m
    colptr = np.concatenate([colptr, colptr[-1] + np.arange(1, m + 1, dtype=np.int64)])
    rowidx = np.concatenate([rowidx, np.arange(m, dtype=np.int32)])
    vals = np.concatenate([vals, np.ones(m, dtype=np.int64)])
    cost = np.concatenate([-rng.integers(1, 101, size=n_struct, dtype=np.int64),
                           np.zeros(m, dtype=np.int64)])
    rhs = np.concatenate([rng.integers(1000, 10001, size=K, dtype=np.int64),
                          np.full(rest, b_big, dtype=np.int64)])
    pivots = [(i, n_struct + i) for i in range(m)]
    return IntegerProblem(m, n, colptr, rowidx, vals, cost, rhs, pivots, full_initial_basis)


def max_flow(n_vertices=2000, out_degree=4, seed=0, max_capacity=100):
    """Config 3: the `Primal` provider of the reference's examples/max_flow.rs:141-223 on a random
    digraph: s = 0, t = V-1, no self arcs, no arcs into s or out of t; integer capacities.
    Rows: V-2 conservation rows then one capacity row per arc; columns: arcs then capacity slacks."""
    rng = np.random.default_rng(seed)
    V = n_vertices
    s, t = 0, V - 1
    arcs = []
    for frm in range(V):
        if frm == t:
            continue
        targets = set()
        while len(targets) < min(out_degree, V - 2):
            to = int(rng.integers(1, V))
            if to != frm:
                targets.add(to)
        for to in sorted(targets):
            arcs.append((frm, to, int(rng.integers(1, max_capacity + 1))))
    arcs.sort(key=lambda a: (a[0], a[1]))          # adjacency columns: by `from`, then `to`
    E = len(arcs)
    nc = V - 2
    shift = lambda v: v - 1                        # s = 0 and t = V-1 removed
    columns, cost = [], []
    for j, (frm, to, cap) in enumerate(arcs):
        col = []
        if frm not in (s, t):
            col.append((shift(frm), -1))
        if to not in (s, t):
            col.append((shift(to), 1))
        col.sort()
        col.append((nc + j, 1))
        columns.append(col)
        cost.append(-1 if frm == s else 0)
    for j in range(E):
        columns.append([(nc + j, 1)])
        cost.append(0)
    rhs = [0] * nc + [a[2] for a in arcs]
    pivots = [(nc + j, E + j) for j in range(E)]
    return IntegerProblem.from_columns(nc + E, columns, cost, rhs, pivots, False)
