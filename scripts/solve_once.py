"""One solve of a bench workload (used under ncu): python scripts/solve_once.py dense16k [max_pivots]"""
import sys
sys.path.insert(0, '.')
import relp_b200, bench
name = sys.argv[1] if len(sys.argv) > 1 else "dense16k"
mp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
prob = bench.make_problem(name, 0)
g = relp_b200.solve_relaxation(prob, rule="steepest_edge", max_pivots=mp)
print(name, g.status, g.pivots, "pivots", f"{g.device_ms:.1f} ms", g.stats["pivots_at_limbs"])
