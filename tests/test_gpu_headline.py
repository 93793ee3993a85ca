"""Parity at the HEADLINE sizes of BASELINE.json (configs 3, 4, 5) and across every limb width / exact-division
width the kernels instantiate.  The CPU side is the C++ oracle (oracle/fast_oracle.cpp, pinned to the Python
oracle which is pinned to the reference's golden fixtures) running on all host threads.

  config 4  sparse 4096 x 8192, K = 90, seed 0, steepest edge: FULL trace, objective and solution
  config 5  dense 16384 x 32768, K = 160, seed 0, steepest edge: trace / objective / solution of a PREFIX
            (the oracle's rational arithmetic makes later pivots cost minutes each), plus an exact
            optimality certificate of the full solve
  config 3  max-flow provider on a 2000-vertex digraph: full trace, objective and solution
"""
import math
import os
from fractions import Fraction as F

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve_both(prob, rule="steepest_edge", max_pivots=0, **kw):
    import relp_b200
    from oracle import fast_oracle as fo
    fo.set_threads(0)
    ref = fo.solve_problem(prob, rule, max_pivots=max_pivots)
    g = relp_b200.solve_relaxation(prob, rule=rule, max_pivots=max_pivots, **kw)
    return g, ref


def _assert_same(g, ref):
    assert g.status == ref.status
    assert g.pivots == len(ref.trace)
    assert g.trace == ref.trace
    if ref.objective is not None:
        assert g.objective == ref.objective
        assert g.bfs == ref.bfs


def test_config4_full_size_trace_objective_solution():
    from relp_b200.generators import bounded_lp
    prob = bounded_lp(4096, 8192, k_bounding=90, nnz_per_col=8, seed=0)
    g, ref = _solve_both(prob)
    assert ref.status == "optimal" and len(ref.trace) > 200
    _assert_same(g, ref)
    assert g.stats["limbs"] >= 8 and g.stats["promotions"] >= 2


def test_config4_full_size_dense_carry_and_no_graph(monkeypatch):
    """same LP through the dense carry kernels and through eager launches: same pivots"""
    import relp_b200
    from relp_b200.generators import bounded_lp
    prob = bounded_lp(4096, 8192, k_bounding=90, nnz_per_col=8, seed=0)
    base = relp_b200.solve_relaxation(prob, rule="steepest_edge")
    dense = relp_b200.solve_relaxation(prob, rule="steepest_edge", dense_carry=True)
    monkeypatch.setenv("RG_NO_GRAPH", "1")
    eager = relp_b200.solve_relaxation(prob, rule="steepest_edge")
    for other in (dense, eager):
        assert (other.status, other.trace, other.objective, other.bfs) == (base.status, base.trace, base.objective,
                                                                          base.bfs)


def _dense16k():
    from relp_b200.generators import bounded_lp
    return bounded_lp(16384, 32768, k_bounding=160, dense=True, seed=0)


def test_config5_prefix_trace_against_oracle():
    """first P pivots of config 5 at its real size (crosses the 2 -> 4 -> 8 limb promotions)"""
    P = int(os.environ.get("RG_TEST_C5_PREFIX", "44"))
    prob = _dense16k()
    g, ref = _solve_both(prob, max_pivots=P)
    assert ref.status == "pivot_limit" and len(ref.trace) == P
    _assert_same(g, ref)
    assert g.stats["limbs"] >= 8


def exact_optimality_certificate(prob, g):
    """Size-independent property: the final basis is primal and dual feasible in EXACT arithmetic, so the
    reported objective is the optimum of the LP -- checked with Python integers, independent of the engine.
    Works for the all-slack-start synthetic LPs (every non-structural basic column is the slack of its row)."""
    m, n = prob.m, prob.n
    ns = n - m                                         # structural columns
    basis = g.basis
    struct_rows = [i for i in range(m) if basis[i] < ns]             # rows whose basic variable is structural
    tight_rows = sorted(set(range(m)) - {basis[i] - ns for i in range(m) if basis[i] >= ns})
    cols = [basis[i] for i in struct_rows]
    k = len(cols)
    assert len(tight_rows) == k
    if prob.dense_block is not None:
        colmat = prob.dense_block[cols].astype(np.int64)             # k x m
    else:
        colmat = np.zeros((k, m), dtype=np.int64)
        for a, j in enumerate(cols):
            for i, v in prob.column(j):
                colmat[a, i] = v
    # solve S x = b_T, S = rows tight_rows x cols (k x k) exactly (fraction-free Gauss-Jordan on Python ints)
    S = [[int(colmat[a, i]) for a in range(k)] + [int(prob.rhs[i])] for i in tight_rows]
    ST = [[int(colmat[a, i]) for i in tight_rows] + [int(prob.cost[cols[a]])] for a in range(k)]   # S^T y = c_B

    def solve(M):
        nrow = len(M)
        M = [[F(v) for v in row] for row in M]
        for c in range(nrow):
            piv = next(r for r in range(c, nrow) if M[r][c] != 0)
            M[c], M[piv] = M[piv], M[c]
            inv = 1 / M[c][c]
            M[c] = [v * inv for v in M[c]]
            for r in range(nrow):
                if r != c and M[r][c] != 0:
                    f = M[r][c]
                    M[r] = [a - f * b for a, b in zip(M[r], M[c])]
        return [row[-1] for row in M]

    x = solve(S)                                       # values of the structural basics
    y = solve(ST)                                      # duals of the tight rows (others are 0)
    assert all(v >= 0 for v in x)
    obj = sum(F(int(prob.cost[j])) * v for j, v in zip(cols, x))
    assert obj == g.objective
    # primal feasibility of every row (slack values) with exact integers over a common denominator
    den = 1
    for v in x:
        den = den * v.denominator // math.gcd(den, v.denominator)
    xn = np.array([int(v * den) for v in x], dtype=object)
    ax = colmat.astype(object).T.dot(xn)               # m entries, numerators over den
    rhs = np.array([int(b) * den for b in prob.rhs], dtype=object)
    assert all(ax[i] <= rhs[i] for i in range(m))
    got = dict(g.bfs)
    for j, v in zip(cols, x):
        assert got.get(j, F(0)) == v
    for i in range(m):
        if basis[i] >= ns:
            assert got.get(basis[i], F(0)) == F(int(rhs[basis[i] - ns] - ax[basis[i] - ns]), den)
    # dual feasibility: reduced cost of every structural column c_j - y . a_j[tight] >= 0, slacks: -y_i >= 0
    assert all(v <= 0 for v in y)
    dy = 1
    for v in y:
        dy = dy * v.denominator // math.gcd(dy, v.denominator)
    yn = np.array([int(v * dy) for v in y], dtype=object)
    if prob.dense_block is not None:
        T = prob.dense_block[:, tight_rows].astype(object)           # ns x k
    else:
        T = np.zeros((ns, k), dtype=object)
        pos = {r: a for a, r in enumerate(tight_rows)}
        for j in range(ns):
            for i, v in prob.column(j):
                if i in pos:
                    T[j, pos[i]] = v
    red = np.array([int(c) * dy for c in prob.cost[:ns]], dtype=object) - T.dot(yn)
    assert all(r >= 0 for r in red)
    return obj


def test_config5_full_solve_exact_optimality_certificate():
    """config 5 solved to optimality at its real size (2 -> 16 limbs); the result is certified optimal by an
    exact primal/dual feasibility check in Python integers"""
    import relp_b200
    prob = _dense16k()
    g = relp_b200.solve_relaxation(prob, rule="steepest_edge")
    assert g.status == "optimal"
    # 783-bit numerators at the peak: 14 limbs of the ladder, narrowed again as they shrink towards the optimum
    assert g.stats["max_bits"] <= 64 * g.stats["limbs"] and g.stats["promotions"] >= 5 and g.stats["demotions"] >= 1
    assert sum(g.stats["pivots_at_limbs"]) == g.pivots and g.stats["pivots_at_limbs"][-1] == 0
    exact_optimality_certificate(prob, g)


def test_config4_exact_optimality_certificate_helper_agrees_with_oracle():
    """the certificate helper itself, on an LP whose optimum the oracle also delivers"""
    from relp_b200.generators import bounded_lp
    prob = bounded_lp(300, 600, k_bounding=40, nnz_per_col=6, seed=4)
    g, ref = _solve_both(prob)
    _assert_same(g, ref)
    assert exact_optimality_certificate(prob, g) == ref.objective


def test_config3_max_flow_2000_full_trace():
    from relp_b200.generators import max_flow
    prob = max_flow(2000, 4, 0)
    g, ref = _solve_both(prob)
    assert ref.status == "optimal" and len(ref.trace) > 1500
    _assert_same(g, ref)
    assert g.denominator == 1 and g.stats["limbs"] == 2      # totally unimodular


def scaled_structural(prob, factor):
    """multiplies every structural column (and its cost) by `factor`: the determinant of a basis with k
    structural columns gains k * log2(factor) trailing zero bits, which drives the exact-division width E"""
    from relp_b200.solver import IntegerProblem
    ns = prob.n - prob.m
    vals = prob.vals.copy()
    vals[: int(prob.colptr[ns])] *= factor
    cost = prob.cost.copy()
    cost[:ns] *= factor
    return IntegerProblem(prob.m, prob.n, prob.colptr, prob.rowidx, vals, cost, prob.rhs, prob.pivots,
                          prob.full_initial_basis)


def _ctz(v):
    return (v & -v).bit_length() - 1


@pytest.mark.parametrize("factor,kb,min_ctz", [(1, 160, 129), (8, 100, 321), (64, 78, 513)])
@pytest.mark.parametrize("no_graph", [False, True])
@pytest.mark.parametrize("ladder", ["pow2", "full"])
def test_limb_and_division_width_sweep(factor, kb, min_ctz, no_graph, ladder, monkeypatch):
    """Runs that end at 16 limbs with ctz(D) beyond 128 / 256 / 512 bits: the K1 variants E in {3,4}, {6,8}
    and the run-time-width kernels (k_update_generic, k_gamma_update) all execute, with graph replay and
    with eager launches; full trace / objective / solution against the oracle."""
    from relp_b200.generators import bounded_lp
    if no_graph:
        monkeypatch.setenv("RG_NO_GRAPH", "1")
    else:
        monkeypatch.delenv("RG_NO_GRAPH", raising=False)
    # the power-of-two ladder reaches the 16-limb kernels (and their E in {6, 8} variants); the full ladder walks
    # 8 -> 10 -> 12 -> 14 (-> 16) and narrows again (demotions down to one limb are allowed here)
    if ladder == "pow2":
        monkeypatch.setenv("RG_WIDTH_LADDER", "pow2")
        monkeypatch.delenv("RG_DEMOTE_FLOOR", raising=False)
    else:
        monkeypatch.delenv("RG_WIDTH_LADDER", raising=False)
        monkeypatch.setenv("RG_DEMOTE_FLOOR", "1")
    base = bounded_lp(256, 512, k_bounding=kb, dense=True, seed=2, dense_block=False)
    prob = scaled_structural(base, factor) if factor > 1 else base
    g, ref = _solve_both(prob, initial_limbs=1)
    assert ref.status == "optimal"
    _assert_same(g, ref)
    if ladder == "pow2":
        assert g.stats["limbs"] == 16, g.stats
    else:
        assert g.stats["limbs"] >= 10 and sum(g.stats["pivots_at_limbs"][4:7]) > 0, g.stats
    assert _ctz(g.denominator) >= min_ctz, (_ctz(g.denominator), g.denominator.bit_length())


def test_limb_sweep_dense_block_wide(monkeypatch):
    """same sweep through the dense int8 block (tensor-core dots at every width up to 16 limbs)"""
    from relp_b200.generators import bounded_lp
    prob = bounded_lp(256, 512, k_bounding=160, dense=True, seed=2, dense_block=True)
    monkeypatch.setenv("RG_WIDTH_LADDER", "pow2")
    g, ref = _solve_both(prob, initial_limbs=1)
    _assert_same(g, ref)
    assert g.stats["limbs"] == 16
    monkeypatch.delenv("RG_WIDTH_LADDER")
    monkeypatch.setenv("RG_DEMOTE_FLOOR", "1")
    g, ref = _solve_both(prob, initial_limbs=1)
    _assert_same(g, ref)
    assert sum(g.stats["pivots_at_limbs"][4:7]) > 0, g.stats      # 10 / 12 / 14 limbs were used


@pytest.mark.parametrize("limbs", [1, 12, 16])
def test_tcgen05_dense_dots_match_mma_sync_and_oracle(limbs, monkeypatch):
    """The dense dots run as tcgen05.mma.kind::i8 (TMA-fed, TMEM accumulators; dense_umma.cuh) once the block has
    a few 128 x 128 tiles; RG_NO_UMMA=1 keeps the mma.sync kernel.  Both must walk the oracle's pivots -- from one
    limb (promotions) and from 16 limbs (the widest slice layouts: 38 tiles over two TMEM layers)."""
    import relp_b200
    from oracle import fast_oracle as fo
    from relp_b200.generators import bounded_lp
    prob = bounded_lp(1100, 640, k_bounding=24, dense=True, seed=9, dense_block=True)
    fo.set_threads(0)
    ref = fo.solve_problem(prob, "steepest_edge")
    assert ref.status == "optimal"
    runs = []
    for no_umma in (False, True):
        if no_umma:
            monkeypatch.setenv("RG_NO_UMMA", "1")
        else:
            monkeypatch.delenv("RG_NO_UMMA", raising=False)
        g = relp_b200.solve_relaxation(prob, rule="steepest_edge", initial_limbs=limbs)
        _assert_same(g, ref)
        runs.append(g.trace)
    assert runs[0] == runs[1]
