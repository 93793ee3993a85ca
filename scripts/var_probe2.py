"""dense16k: repeatability of the loop time, profile 0, six solves (env switches are read at rg_create)."""
import os, sys
sys.path.insert(0, '.')
import relp_b200, bench
prob = bench.make_problem(sys.argv[1] if len(sys.argv) > 1 else "dense16k", 0)
g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=0)   # warm-up
ts = []
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=0)
    ts.append(g.device_ms)
print(" ".join(f"{t:.0f}" for t in ts), "ms;", {k: os.environ[k] for k in os.environ if k.startswith("RG_")}, flush=True)
