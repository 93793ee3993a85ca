import sys, time
sys.path.insert(0, '.')
import relp_b200
from relp_b200.generators import bounded_lp, max_flow
which = sys.argv[1]
rule = sys.argv[2] if len(sys.argv) > 2 else "steepest_edge"
t = time.time()
if which == "c4":
    prob = bounded_lp(4096, 8192, k_bounding=90, nnz_per_col=8, seed=0)
elif which == "c4s":
    prob = bounded_lp(1024, 2048, k_bounding=60, nnz_per_col=8, seed=0)
elif which == "c5":
    prob = bounded_lp(16384, 32768, k_bounding=90, dense=True, seed=0)
elif which == "c5s":
    prob = bounded_lp(2048, 4096, k_bounding=90, dense=True, seed=0)
elif which == "mf":
    prob = max_flow(int(sys.argv[3]) if len(sys.argv) > 3 else 2000, 4, 0)
print("generated", which, prob.m, prob.n, prob.vals.shape, round(time.time() - t, 2), flush=True)
maxp = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t = time.time()
dense_carry = len(sys.argv) > 5 and sys.argv[5] == "dense_carry"
g = relp_b200.solve_relaxation(prob, rule=rule, max_pivots=maxp, dense_carry=dense_carry)
print("status", g.status, "pivots", g.pivots, "loop s", round(g.seconds, 4), "total s", round(g.seconds_total, 3),
      "wall", round(time.time() - t, 3))
print("pivots/s", round(g.pivots / g.seconds, 1), "stats", g.stats)
print("objective", float(g.objective), "den bits", g.denominator.bit_length())
