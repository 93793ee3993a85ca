"""dense16k: repeatability of the loop time at each profiling level (one process, alternating levels)."""
import os, sys
sys.path.insert(0, '.')
import relp_b200, bench
prob = bench.make_problem(sys.argv[1] if len(sys.argv) > 1 else "dense16k", 0)
g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=0)   # warm-up
for prof in (0, 0, 1, 1, 0, 1, 0, 1, 2, 2, 0):
    g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=prof)
    print(f"profile={prof} pivots={g.pivots} device {g.device_ms:.1f} ms loop {g.seconds*1e3:.1f} ms -> {g.pivots/(g.device_ms/1e3):.1f} pivots/s", flush=True)
