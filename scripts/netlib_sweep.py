"""One-off sweep (build container only: reads /root/reference): every non-ignored test of the reference's
tests/netlib/test.rs through the restated pipeline (fixed-format reader, presolve, standardize, MatrixData, C++ oracle
with a 40 s limit, reconstruction) against the objective and tolerance that file asserts.
Output of the round-2 run: profiles/r2_netlib_sweep_cpu.txt."""
import re, sys, os, time, signal
sys.path.insert(0,'/root/repo')
from fractions import Fraction as F
from relp_b200 import frontend, presolve, mps as reader
from oracle import relp_oracle as ro, fast_oracle as fo
src=open('/root/reference/tests/netlib/test.rs').read()
tests=[]
for blk in re.split(r"#\[test\]", src)[1:]:
    ign = "#[ignore" in blk
    m1 = re.search(r'fn test_(\w+)\(\)', blk); m2 = re.search(r'solve\("([^"]+)"\)', blk)
    m3 = re.search(r'let expected = ([-+0-9.eE_]+)', blk); m4 = re.search(r'< RB!\(([-+0-9.eE]+)\)', blk)
    if m1 and m2 and m3 and m4:
        tests.append((ign, m1.group(1), m2.group(1), m3.group(1).replace("_",""), m4.group(1)))
print(len(tests), "tests")
fo.set_threads(0)
for ign, name, fname, expected, tol in tests:
    if ign: continue
    path=f'/root/reference/tests/netlib/problem_files/{fname}.SIF'
    size=os.path.getsize(path)
    t0=time.time()
    try:
        text=open(path).read()
        m=reader.parse_fixed(text); gf=m.to_general_form()
        d=frontend.parse_mps.__wrapped__ if hasattr(frontend.parse_mps,'__wrapped__') else None
        # build dict like parse_mps but forcing fixed mode
        mp=frontend.parse_mps(text) if True else None
        mp['general_form']=reader.parse_fixed(text).to_general_form()
        try:
            lp=frontend.canonicalize(mp)
        except presolve.FiniteOptimum as e:
            print(f"{fname:10s} size {size:7d} solved by presolve obj {float(e.objective):.8f} expected {expected} OK={abs(float(e.objective)-float(eval(expected)))<float(tol)}"); continue
        variables=[ro.Variable(c,u) for c,u in zip(lp.costs, lp.upper)]
        md=ro.MatrixData(lp.constraint_columns, lp.b, lp.ranges, *lp.counts, variables)
        fo.set_time_limit(40.0)
        ref=fo.solve_provider(md,"steepest_edge")
        dt=time.time()-t0
        if ref.status!="optimal":
            print(f"{fname:10s} size {size:7d} rows {len(lp.b)} cols {len(lp.costs)} status {ref.status} after {dt:.1f}s pivots {len(ref.trace)}"); continue
        sol=frontend.recover(lp, ref.bfs, ref.objective)
        ok=abs(float(sol.objective_value)-float(eval(expected)))<float(tol)
        print(f"{fname:10s} size {size:7d} rows {len(lp.b)} cols {len(lp.costs)} pivots {len(ref.trace)} {dt:.1f}s obj {float(sol.objective_value):.9f} expected {expected} OK={ok}", flush=True)
    except Exception as e:
        print(f"{fname:10s} size {size:7d} ERROR {type(e).__name__}: {str(e)[:100]}", flush=True)
