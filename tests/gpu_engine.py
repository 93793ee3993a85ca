"""ctypes face of the trait-shaped C ABI (include/relp_gpu.h) for the GPU tests: one method per entry
point, numbers converted to exact `fractions.Fraction`s over the engine's common denominator."""
import ctypes as C
from fractions import Fraction as F

import numpy as np

from relp_b200 import _lib
from relp_b200.solver import RULES, limbs_to_int


class Engine:
    def __init__(self, problem, initial_limbs=0, dense_carry=False):
        self.lib = _lib.load()
        self.p = problem
        self.m, self.n = problem.m, problem.n
        opts = _lib.rg_options(device=0, initial_limbs=initial_limbs, rank=0, world=1,
                               dense_carry=1 if dense_carry else 0, reserved=0, nccl_unique_id=None)
        self.ctx = C.c_void_p()
        self._ck(self.lib.rg_create(C.byref(opts), C.byref(self.ctx)), "rg_create")
        i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        self._ck(self.lib.rg_load_csc(self.ctx, self.m, self.n, problem.colptr.ctypes.data_as(i64p),
                                      problem.rowidx.ctypes.data_as(i32p), problem.vals.ctypes.data_as(i64p)),
                 "rg_load_csc")
        if problem.dense_block is not None:
            self._ck(self.lib.rg_load_dense_i8(self.ctx, problem.dense_block.shape[0],
                                               problem.dense_block.ctypes.data_as(C.POINTER(C.c_int8))),
                     "rg_load_dense_i8")
        self._ck(self.lib.rg_set_rhs(self.ctx, problem.rhs.ctypes.data_as(i64p)), "rg_set_rhs")

    def close(self):
        if self.ctx:
            self.lib.rg_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, where):
        if rc != 0:
            msg = self.lib.rg_last_error(self.ctx).decode() if self.ctx else ""
            raise RuntimeError(f"{where} failed ({rc}): {msg}")

    # ---- constructors / rule -------------------------------------------------------------------------
    def init_identity_basis(self, basis_ids, cost=None):
        b = np.ascontiguousarray(basis_ids, dtype=np.int32)
        c = None if cost is None else np.ascontiguousarray(cost, dtype=np.int64)
        self._ck(self.lib.rg_init_identity_basis(self.ctx, b.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 None if c is None else c.ctypes.data_as(C.POINTER(C.c_int64))),
                 "rg_init_identity_basis")

    def init_basis(self, basis_cols, cost):
        b = np.ascontiguousarray(basis_cols, dtype=np.int32)
        c = np.ascontiguousarray(cost, dtype=np.int64)
        self._ck(self.lib.rg_init_basis(self.ctx, b.ctypes.data_as(C.POINTER(C.c_int32)),
                                        c.ctypes.data_as(C.POINTER(C.c_int64))), "rg_init_basis")

    def phase_switch(self, cost):
        c = np.ascontiguousarray(cost, dtype=np.int64)
        self._ck(self.lib.rg_phase_switch(self.ctx, c.ctypes.data_as(C.POINTER(C.c_int64))), "rg_phase_switch")

    def rule_new(self, rule):
        self._ck(self.lib.rg_rule_new(self.ctx, RULES[rule]), "rg_rule_new")

    # ---- iteration -----------------------------------------------------------------------------------
    def select_column(self):
        st, q = C.c_int32(), C.c_int32()
        self._ck(self.lib.rg_select_primal_pivot_column(self.ctx, C.byref(st), C.byref(q)), "select column")
        return None if st.value == 1 else q.value

    def generate_column(self, q):
        self._ck(self.lib.rg_generate_column(self.ctx, q), "rg_generate_column")

    def select_row(self):
        st, r = C.c_int32(), C.c_int32()
        self._ck(self.lib.rg_select_primal_pivot_row(self.ctx, C.byref(st), C.byref(r)), "select row")
        return None if st.value == 2 else r.value

    def bring_into_basis(self, q, row, update_rule=True):
        info = _lib.rg_pivot_info()
        self._ck(self.lib.rg_bring_into_basis(self.ctx, q, row, 1 if update_rule else 0, C.byref(info)),
                 "rg_bring_into_basis")
        return info.entering, info.row, info.leaving

    def remove_artificial_row(self, row):
        info = _lib.rg_pivot_info()
        self._ck(self.lib.rg_remove_artificial_row(self.ctx, row, C.byref(info)), "rg_remove_artificial_row")
        return (info.status == 0), info.entering, info.row, info.leaving

    # ---- exports -------------------------------------------------------------------------------------
    def limbs(self):
        L = C.c_int32()
        self._ck(self.lib.rg_get_limbs(self.ctx, C.byref(L)), "rg_get_limbs")
        return L.value

    def _ints(self, buf, count, nl):
        return [limbs_to_int([buf[i * nl + k] for k in range(nl)]) for i in range(count)]

    def denominator(self):
        L = self.limbs()
        buf = (C.c_uint64 * L)()
        self._ck(self.lib.rg_get_denominator(self.ctx, buf), "rg_get_denominator")
        return limbs_to_int(list(buf))

    def _vec(self, fn, count, nl, *args):
        buf = (C.c_uint64 * (count * nl))()
        self._ck(fn(self.ctx, *args, buf), fn.__name__)
        return self._ints(buf, count, nl)

    def b(self):
        D = self.denominator()
        return [F(v, D) for v in self._vec(self.lib.rg_get_b, self.m, self.limbs())]

    def minus_objective(self):
        D = self.denominator()
        return F(self._vec(self.lib.rg_get_minus_objective, 1, self.limbs())[0], D)

    def minus_pi(self):
        D = self.denominator()
        return [F(v, D) for v in self._vec(self.lib.rg_get_minus_pi, self.m, self.limbs())]

    def basis_inverse_row(self, row):
        D = self.denominator()
        return [F(v, D) for v in self._vec(self.lib.rg_get_basis_inverse_row, self.m, self.limbs(), row)]

    def pivot_column(self):
        D = self.denominator()
        return [F(v, D) for v in self._vec(self.lib.rg_get_pivot_column, self.m, self.limbs() + 2)]

    def relative_costs(self):
        D = self.denominator()
        return [F(v, D) for v in self._vec(self.lib.rg_get_relative_costs, self.n, self.limbs() + 2)]

    def gamma(self):
        D = self.denominator()
        return [F(v, D * D) for v in self._vec(self.lib.rg_get_gamma, self.n, 2 * self.limbs() + 6)]

    def element(self, row, j):
        D = self.denominator()
        return F(self._vec(self.lib.rg_get_element, 1, self.limbs() + 2, row, j)[0], D)

    def basis(self):
        buf = (C.c_int32 * self.m)()
        self._ck(self.lib.rg_get_basis(self.ctx, buf), "rg_get_basis")
        return list(buf)

    def basis_change_info(self, want_work=True):
        """(column_before_change, work_vector or None, basis_inverse_row) as Fractions"""
        L = self.limbs()
        col = (C.c_uint64 * (self.m * (L + 2)))()
        work = (C.c_uint64 * (self.m * (2 * L + 5)))() if want_work else None
        row = (C.c_uint64 * (self.m * L))()
        dold = (C.c_uint64 * L)()
        self._ck(self.lib.rg_get_basis_change_info(self.ctx, col, work, row, dold), "rg_get_basis_change_info")
        D0 = limbs_to_int(list(dold))
        D1 = self.denominator()
        column = [F(v, D0) for v in self._ints(col, self.m, L + 2)]
        w = [F(v, D0 * D0) for v in self._ints(work, self.m, 2 * L + 5)] if want_work else None
        r = [F(v, D1) for v in self._ints(row, self.m, L)]
        return column, w, r
