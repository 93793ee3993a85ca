"""ctypes face of oracle/fast_oracle.cpp (CPU ORACLE -- test infrastructure and timed CPU baseline only;
the product path never imports this)."""
import ctypes as C
import os
import subprocess
import time
from fractions import Fraction

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfast_oracle.so")
RULES = {"first_profitable": 0, "first_profitable_with_memory": 1, "dantzig": 2, "steepest_edge": 3}
_lib = None


def available():
    try:
        load()
        return True
    except Exception:
        return False


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "fast_oracle.cpp")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    lib = C.CDLL(LIB)
    P = C.c_void_p
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    lib.fo_solve.restype = C.c_int
    lib.fo_solve.argtypes = [C.c_int, C.c_int, i64p, i32p, i64p, i64p, i64p, i64p, i64p, i64p, C.c_int,
                             i32p, i32p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.POINTER(C.c_int8),
                             C.POINTER(P)]
    lib.fo_set_threads.restype = C.c_int
    lib.fo_set_threads.argtypes = [C.c_int]
    lib.fo_set_time_limit.restype = None
    lib.fo_set_time_limit.argtypes = [C.c_double]
    for name, res in [("fo_status", C.c_int), ("fo_trace_len", C.c_longlong), ("fo_trace", i32p),
                      ("fo_rows_removed_len", C.c_int), ("fo_rows_removed", i32p), ("fo_bfs_len", C.c_int),
                      ("fo_bfs_cols", i32p), ("fo_bfs_values", C.c_char_p), ("fo_objective", C.c_char_p),
                      ("fo_nr_artificial", C.c_int), ("fo_seconds", C.c_double)]:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = [P]
    lib.fo_free.restype = None
    lib.fo_free.argtypes = [P]
    lib.fo_arith.restype = C.c_char_p
    lib.fo_arith.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int]
    _lib = lib
    return lib


def parse_rational(s):
    s = s.decode() if isinstance(s, bytes) else s
    num, den = s.split("/")
    neg = num.startswith("-")
    v = Fraction(int(num.lstrip("-"), 16), int(den, 16))
    return -v if neg else v


class FastResult:
    STATUS = {0: "optimal", 1: "unbounded", 2: "infeasible", -1: "pivot_limit", -2: "error"}

    def __init__(self):
        self.status = None
        self.trace = []
        self.objective = None
        self.bfs = []
        self.rows_removed = []
        self.nr_artificial = 0
        self.seconds = 0.0


def _split(values):
    """list of Fractions/ints -> (num int64 array, den int64 array)"""
    num = np.array([Fraction(v).numerator for v in values], dtype=np.int64)
    den = np.array([Fraction(v).denominator for v in values], dtype=np.int64)
    return num, den


def set_threads(n=0):
    """OpenMP team size of the column-parallel loops (0 = all hardware threads); returns the size in use.
    Results do not depend on it (exact arithmetic, index-ordered reductions)."""
    return load().fo_set_threads(int(n))


def solve(m, n, colptr, rowidx, vals, cost, rhs, pivots, full_basis, rule="steepest_edge", max_pivots=0,
          vals_den=None, cost_den=None, rhs_den=None, dense_block=None):
    """Integer (or num/den) CSC problem -> FastResult.  Trace rows are in the original row space.
    dense_block: optional int8 array (n_dense, m): provider columns [0, n_dense), CSC ranges empty."""
    lib = load()
    p64 = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))
    p32 = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.int64) if a is not None else None
            for a in (vals, vals_den, cost, cost_den, rhs, rhs_den)]
    if pivots is None:
        npv, pr, pc = -1, None, None
    else:
        npv = len(pivots)
        pr = np.array([r for r, _ in pivots], dtype=np.int32)
        pc = np.array([c for _, c in pivots], dtype=np.int32)
    h = C.c_void_p()
    nd, dptr = 0, None
    if dense_block is not None:
        dense_block = np.ascontiguousarray(dense_block, dtype=np.int8)
        assert dense_block.ndim == 2 and dense_block.shape[1] == m
        nd, dptr = dense_block.shape[0], dense_block.ctypes.data_as(C.POINTER(C.c_int8))
    rc = lib.fo_solve(m, n, p64(colptr), p32(rowidx), p64(arrs[0]), p64(arrs[1]), p64(arrs[2]), p64(arrs[3]),
                      p64(arrs[4]), p64(arrs[5]), npv, p32(pr), p32(pc), 1 if full_basis else 0,
                      RULES[rule], max_pivots, nd, dptr, C.byref(h))
    assert rc == 0
    try:
        res = FastResult()
        res.status = FastResult.STATUS[lib.fo_status(h)]
        k = lib.fo_trace_len(h)
        t = lib.fo_trace(h)
        res.trace = [(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]) for i in range(k)]
        rr = lib.fo_rows_removed(h)
        res.rows_removed = [rr[i] for i in range(lib.fo_rows_removed_len(h))]
        res.nr_artificial = lib.fo_nr_artificial(h)
        res.seconds = lib.fo_seconds(h)
        if res.status == "optimal" or (res.status == "pivot_limit" and lib.fo_objective(h)):
            # (pivot limit inside phase two: the objective and basic solution of the basis reached)
            res.objective = parse_rational(lib.fo_objective(h))
            cols = lib.fo_bfs_cols(h)
            nb = lib.fo_bfs_len(h)
            vals_s = lib.fo_bfs_values(h).decode()
            vs = [parse_rational(x) for x in vals_s.split(";")] if nb else []
            res.bfs = [(cols[i], vs[i]) for i in range(nb)]
        return res
    finally:
        lib.fo_free(h)


def solve_problem(problem, rule="steepest_edge", max_pivots=0):
    """relp_b200.IntegerProblem (duck-typed: m, n, colptr, rowidx, vals, cost, rhs, pivots,
    full_initial_basis) -> FastResult"""
    return solve(problem.m, problem.n, problem.colptr, problem.rowidx, problem.vals, problem.cost,
                 problem.rhs, problem.pivots, problem.full_initial_basis, rule, max_pivots,
                 dense_block=getattr(problem, "dense_block", None))


def solve_provider(provider, rule="steepest_edge", max_pivots=0):
    """oracle.relp_oracle provider with rational data (numerators/denominators below 2^62)."""
    m, n = provider.nr_rows(), provider.nr_columns()
    colptr = [0]
    rowidx, vals = [], []
    for j in range(n):
        for i, v in provider.column(j):
            rowidx.append(i)
            vals.append(v)
        colptr.append(len(rowidx))
    vn, vd = _split(vals)
    cn, cd = _split([provider.cost_value(j) for j in range(n)])
    bn, bd = _split(provider.right_hand_side())
    pivots = provider.pivot_element_indices() if provider.has_partial_initial_basis else None
    return solve(m, n, colptr, rowidx, vn, cn, bn, pivots, provider.has_full_initial_basis, rule,
                 max_pivots, vals_den=vd, cost_den=cd, rhs_den=bd)


def set_time_limit(seconds=0.0):
    """A solve stops like at a pivot limit once it has run this long (0: no limit)."""
    load().fo_set_time_limit(float(seconds))


def timed_sample(problem, rule, budget_s=15.0):
    """bench.py cpu_baseline: the pivots of the same LP and trace done within ~budget_s seconds (the rule
    initialisation is part of the sample, as it is part of the GPU's timed loop)."""
    set_time_limit(budget_s)
    try:
        r = solve_problem(problem, rule)
    finally:
        set_time_limit(0)
    n = len(r.trace)
    whole = r.status != "pivot_limit"
    return {"value": n / max(r.seconds, 1e-9), "unit": "pivots/s", "cores": 1, "kind": "port",
            "sample": (f"{'all' if whole else 'first'} {n} pivots of the same LP and trace (incl. rule "
                       f"initialisation), C++ big-rational restatement of Carry<RationalBig, "
                       f"BasisInverseRows> (oracle/fast_oracle.cpp), {r.seconds:.1f} s; a prefix overstates the "
                       f"CPU (early pivots have the smallest numbers)")}
