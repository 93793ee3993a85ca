"""Python face of the host driver: builds an integer `MatrixProvider` image, calls
`rh_solve_relaxation` (include/relp_host.h) and converts the limb results to exact fractions.

Mirrors `SolveRelaxation::solve_relaxation` (reference src/algorithm/mod.rs:17-36): the result is
`FiniteOptimum(bfs)`, `Unbounded` or `Infeasible` (src/algorithm/mod.rs:43-47).
"""
import ctypes as C
from fractions import Fraction

import numpy as np

from . import _lib

RULES = {"first_profitable": 0, "first_profitable_with_memory": 1, "dantzig": 2, "steepest_edge": 3}
STATUS = {0: "optimal", 1: "unbounded", 2: "infeasible", -1: "pivot_limit"}


class IntegerProblem:
    """A materialised integer MatrixProvider (reference matrix_provider/mod.rs:37-134).

    columns: CSC (colptr int64[n+1], rowidx int32[nnz] ascending per column, vals int64[nnz]);
    cost int64[n]; rhs int64[m] (>= 0); pivots: list of (row, column) positive unit slack pivots
    (`PartialInitialBasis::pivot_element_indices`) or None; full_initial_basis: `FullInitialBasis`.
    """

    def __init__(self, m, n, colptr, rowidx, vals, cost, rhs, pivots=None, full_initial_basis=False):
        self.m, self.n = int(m), int(n)
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        self.rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
        self.vals = np.ascontiguousarray(vals, dtype=np.int64)
        self.cost = np.ascontiguousarray(cost, dtype=np.int64)
        self.rhs = np.ascontiguousarray(rhs, dtype=np.int64)
        assert self.colptr.shape == (self.n + 1,) and self.cost.shape == (self.n,)
        assert self.rhs.shape == (self.m,) and self.rowidx.shape == self.vals.shape
        assert int(self.colptr[-1]) == self.vals.shape[0]
        if (self.rhs < 0).any():
            raise ValueError("right-hand side must be non-negative")
        self.pivots = None if pivots is None else [(int(r), int(c)) for r, c in pivots]
        self.full_initial_basis = bool(full_initial_basis)
        if self.full_initial_basis:
            assert self.pivots is not None and len(self.pivots) == self.m
        # weights of a prescaled rational problem (relp_b200.frontend.prescale); None = integer problem
        self.col_weight = None     # w_j
        self.row_scale = None      # r_i
        self.W = 1                 # lcm of all weights
        self.cost_scale = 1        # integer costs = cost_scale * c_j / w_j
        # optional dense int8 block: columns [0, n_dense) stored column-major (n_dense, m); their CSC
        # ranges are empty (config 5: implicit row indices, 1 byte per coefficient)
        self.dense_block = None

    @classmethod
    def from_columns(cls, m, columns, cost, rhs, pivots=None, full_initial_basis=False):
        """columns: list of [(row, int value)] sorted by row."""
        colptr = np.zeros(len(columns) + 1, dtype=np.int64)
        rowidx, vals = [], []
        for j, col in enumerate(columns):
            for i, v in col:
                rowidx.append(int(i))
                vals.append(int(v))
            colptr[j + 1] = len(rowidx)
        return cls(m, len(columns), colptr, np.array(rowidx, dtype=np.int32),
                   np.array(vals, dtype=np.int64), cost, rhs, pivots, full_initial_basis)

    def set_dense_block(self, block):
        block = np.ascontiguousarray(block, dtype=np.int8)
        assert block.ndim == 2 and block.shape[1] == self.m and block.shape[0] <= self.n
        assert int(self.colptr[block.shape[0]]) == 0, "CSC ranges of the dense columns must be empty"
        self.dense_block = block

    def column(self, j):
        if self.dense_block is not None and j < self.dense_block.shape[0]:
            col = self.dense_block[j]
            return [(int(i), int(col[i])) for i in np.nonzero(col)[0]]
        a, b = int(self.colptr[j]), int(self.colptr[j + 1])
        return [(int(self.rowidx[k]), int(self.vals[k])) for k in range(a, b)]

    def _as_c(self):
        p = _lib.rh_problem()
        p.m, p.n = self.m, self.n
        p.colptr = self.colptr.ctypes.data_as(C.POINTER(C.c_int64))
        p.rowidx = self.rowidx.ctypes.data_as(C.POINTER(C.c_int32))
        p.vals = self.vals.ctypes.data_as(C.POINTER(C.c_int64))
        p.cost = self.cost.ctypes.data_as(C.POINTER(C.c_int64))
        p.rhs = self.rhs.ctypes.data_as(C.POINTER(C.c_int64))
        keep = []
        if self.pivots is None:
            p.n_pivots = -1
            p.pivot_rows = None
            p.pivot_cols = None
        else:
            rows = np.array([r for r, _ in self.pivots], dtype=np.int32)
            cols = np.array([c for _, c in self.pivots], dtype=np.int32)
            keep = [rows, cols]
            p.n_pivots = len(self.pivots)
            p.pivot_rows = rows.ctypes.data_as(C.POINTER(C.c_int32))
            p.pivot_cols = cols.ctypes.data_as(C.POINTER(C.c_int32))
        p.full_initial_basis = 1 if self.full_initial_basis else 0
        if self.dense_block is not None:
            keep.append(self.dense_block)
            p.n_dense = self.dense_block.shape[0]
            p.dense = self.dense_block.ctypes.data_as(C.POINTER(C.c_int8))
        if self.col_weight is not None:
            from math import gcd
            w = [int(x) for x in self.col_weight]
            r = [int(x) for x in self.row_scale]
            W = int(self.W)
            real_rows = set(rc[0] for rc in self.pivots) if self.pivots is not None else set()
            w1 = 1
            for i in range(self.m):
                if i not in real_rows:
                    w1 = w1 * r[i] // gcd(w1, r[i])
            colfac = np.array([W // x for x in w], dtype=np.int64)
            artfac = np.array([W // x for x in r], dtype=np.int64)
            colw = np.array(w, dtype=np.int64)
            artcost = np.array([w1 // x for x in r], dtype=np.int64)
            keep += [colfac, artfac, colw, artcost]
            p.colfac = colfac.ctypes.data_as(C.POINTER(C.c_int64))
            p.artfac = artfac.ctypes.data_as(C.POINTER(C.c_int64))
            p.colw = colw.ctypes.data_as(C.POINTER(C.c_int64))
            p.artcost = artcost.ctypes.data_as(C.POINTER(C.c_int64))
        return p, keep


def limbs_to_int(words):
    """two's complement little-endian 64-bit words -> Python int"""
    n = len(words)
    v = 0
    for k in range(n - 1, -1, -1):
        v = (v << 64) | int(words[k])
    if n and int(words[n - 1]) >> 63:
        v -= 1 << (64 * n)
    return v


class GpuResult:
    def __init__(self):
        self.status = None
        self.trace = []            # [(phase, entering, row, leaving)] reference index space
        self.pivots = 0
        self.objective = None      # Fraction (of the integer problem handed in)
        self.denominator = None    # int, |det B|
        self.bfs = []              # sorted [(provider column, Fraction)] non-zero basic values
        self.basis = []            # engine column id per row (negative: inert artificial)
        self.nr_artificial = 0
        self.rows_removed = []
        self.stats = {}
        self.seconds = 0.0
        self.seconds_total = 0.0
        self.device_ms = 0.0


def nccl_unique_id():
    """A fresh 128-byte NCCL id (call on rank 0, broadcast to the other ranks)."""
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    rc = lib.rg_nccl_unique_id(buf, 128)
    if rc < 0:
        raise RuntimeError("rg_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw


def solve_relaxation(problem, rule="steepest_edge", fused=True, initial_limbs=0, device=0,
                     max_pivots=0, profile=False, rank=0, world=1, nccl_id=None, dense_carry=False):
    """Solves `problem`.  world > 1: row-sharded over `world` GPUs (one process per GPU; every rank
    passes the same problem and the same `nccl_id` and receives the same result)."""
    lib = _lib.load()
    cprob, keep = problem._as_c()
    idbuf = None
    if world > 1:
        assert nccl_id is not None and len(nccl_id) == 128
        idbuf = C.create_string_buffer(bytes(nccl_id), 128)
    opts = _lib.rh_options(device=device, initial_limbs=initial_limbs, rule=RULES[rule],
                           fused=1 if fused else 0, max_pivots=max_pivots,
                           profile=int(profile), rank=rank, world=world,
                           dense_carry=1 if dense_carry else 0,
                           nccl_unique_id=C.cast(idbuf, C.c_void_p) if idbuf is not None else None)
    handle = C.c_void_p()
    import os, time
    _t0 = time.perf_counter()
    rc = lib.rh_solve_relaxation(C.byref(cprob), C.byref(opts), C.byref(handle))
    if os.environ.get("RG_HOSTPROF"):
        print(f"[hostprof] rh_solve_relaxation call {time.perf_counter() - _t0:.3f} s", flush=True)
    try:
        if rc != 0:
            msg = lib.rh_result_error(handle).decode() if handle else ""
            raise RuntimeError(f"rh_solve_relaxation failed ({rc}): {msg}")
        res = GpuResult()
        res.status = STATUS[lib.rh_result_status(handle)]
        res.pivots = lib.rh_result_pivots(handle)
        n = lib.rh_result_trace_len(handle)
        tr = lib.rh_result_trace(handle)
        res.trace = [(tr[k].phase, tr[k].entering, tr[k].row, tr[k].leaving) for k in range(n)]
        L = lib.rh_result_limbs(handle)
        mo = lib.rh_result_minus_objective(handle)
        dn = lib.rh_result_denominator(handle)
        D = limbs_to_int([dn[k] for k in range(L)])
        res.denominator = D
        res.objective = Fraction(-limbs_to_int([mo[k] for k in range(L)]), D) / int(problem.cost_scale)
        bs = lib.rh_result_basis(handle)
        bw = lib.rh_result_b(handle)
        res.basis = [bs[i] for i in range(problem.m)]
        bfs = []
        for i in range(problem.m):
            v = limbs_to_int([bw[i * L + k] for k in range(L)])
            if v != 0 and res.basis[i] >= 0:
                wj = 1 if problem.col_weight is None else int(problem.col_weight[res.basis[i]])
                bfs.append((res.basis[i], Fraction(v, D) / wj))
        bfs.sort(key=lambda t: t[0])
        res.bfs = bfs
        res.nr_artificial = lib.rh_result_nr_artificial(handle)
        rr = lib.rh_result_rows_removed(handle)
        res.rows_removed = [rr[k] for k in range(lib.rh_result_rows_removed_len(handle))]
        st = _lib.rg_stats()
        lib.rh_result_stats(handle, C.byref(st))
        res.stats = dict(pivots=st.pivots, promotions=st.promotions, limbs=st.limbs,
                         max_bits=st.max_bits, denominator_bits=st.denominator_bits,
                         kernel_launches=st.kernel_launches, demotions=st.demotions,
                         limb_widths=[st.limb_widths[k] for k in range(_lib.RG_NWIDTHS)],
                         pivots_at_limbs=[st.pivots_at_limbs[k] for k in range(_lib.RG_NWIDTHS)],
                         k1_launches_at_limbs=[st.k1_launches_at_limbs[k] for k in range(_lib.RG_NWIDTHS)],
                         k1_ms_at_limbs=[st.k1_ms_at_limbs[k] for k in range(_lib.RG_NWIDTHS)],
                         k1_bytes_at_limbs=[st.k1_bytes_at_limbs[k] for k in range(_lib.RG_NWIDTHS)],
                         k1_imads_at_limbs=[st.k1_imads_at_limbs[k] for k in range(_lib.RG_NWIDTHS)],
                         phase_ms=[st.phase_ms[k] for k in range(8)],
                         active_columns=st.reserved)
        res.device_ms = lib.rh_result_device_ms(handle)
        res.seconds = lib.rh_result_seconds(handle)
        res.seconds_total = lib.rh_result_seconds_total(handle)
        return res
    finally:
        if handle:
            lib.rh_result_free(handle)
        del keep
