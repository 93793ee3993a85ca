"""Step-by-step GPU vs oracle state dump (debug aid, not a test)."""
import ctypes as C, sys
from fractions import Fraction as F
sys.path.insert(0, '.')
from relp_b200 import _lib
from relp_b200.solver import limbs_to_int
from tests.common import problem_from_provider
from tests.test_oracle_golden import problem_2
from oracle import relp_oracle as ro

lib = _lib.load()
prov = problem_2()
prob = problem_from_provider(prov)
ctx = C.c_void_p()
opts = _lib.rg_options(0, int(sys.argv[1]) if len(sys.argv) > 1 else 2, 0, 1)
assert lib.rg_create(C.byref(opts), C.byref(ctx)) == 0
p32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
assert lib.rg_load_csc(ctx, prob.m, prob.n, p64(prob.colptr), p32(prob.rowidx), p64(prob.vals)) == 0
assert lib.rg_set_rhs(ctx, p64(prob.rhs)) == 0
import numpy as np
basis = np.array([-3, -2, -1], dtype=np.int32)
assert lib.rg_init_identity_basis(ctx, p32(basis), None) == 0

def getv(fn, count, per, *args):
    L = C.c_int32(); lib.rg_get_limbs(ctx, C.byref(L)); L = L.value
    w = per(L)
    buf = (C.c_uint64 * (count * w))()
    rc = fn(ctx, *args, buf)
    assert rc == 0, (rc, lib.rg_last_error(ctx))
    return [limbs_to_int(buf[i * w:(i + 1) * w]) for i in range(count)]

def dump(tag):
    D = getv(lib.rg_get_denominator, 1, lambda L: L)[0]
    print(tag, "D", D, "obj", getv(lib.rg_get_minus_objective, 1, lambda L: L), "pi", getv(lib.rg_get_minus_pi, prob.m, lambda L: L),
          "b", getv(lib.rg_get_b, prob.m, lambda L: L))
    for i in range(prob.m):
        print("   row", i, getv(lib.rg_get_basis_inverse_row, prob.m, lambda L: L, i))
    print("   kappa", getv(lib.rg_get_relative_costs, prob.n, lambda L: L + 2))

dump("init")
assert lib.rg_rule_new(ctx, 0) == 0
for it in range(4):
    st = C.c_int32(); q = C.c_int32(); row = C.c_int32()
    assert lib.rg_select_primal_pivot_column(ctx, C.byref(st), C.byref(q)) == 0
    print("select", st.value, q.value)
    if st.value != 0: break
    assert lib.rg_generate_column(ctx, q.value) == 0
    print("   column", getv(lib.rg_get_pivot_column, prob.m, lambda L: L + 2))
    assert lib.rg_select_primal_pivot_row(ctx, C.byref(st), C.byref(row)) == 0
    print("row", st.value, row.value)
    info = _lib.rg_pivot_info()
    rc = lib.rg_bring_into_basis(ctx, q.value, row.value, 1, C.byref(info))
    print("pivot rc", rc, lib.rg_last_error(ctx), info.entering, info.row, info.leaving)
    raw = (C.c_int32 * 20)(); lib.rg_debug_scalars(ctx, raw, 80); print("   scalars", list(raw))
    big = (C.c_uint64 * 200)(); lib.rg_debug_scalars(ctx, big, 1600); print("   D,Dnew,Dinv,A,up", [hex(big[10+k]) for k in (0,1,16,17,32,33,34,66,67,68,100,101,102)])
    print("   u", getv(lib.rg_debug_vector, prob.m+1, lambda L: L+2, 0)); print("   rowp", getv(lib.rg_debug_vector, prob.m+1, lambda L: L, 1))
    dump("after %d" % it)
st = _lib.rg_stats(); lib.rg_get_stats(ctx, C.byref(st)); print("limbs", st.limbs, "maxbits", st.max_bits, "launches", st.kernel_launches)
