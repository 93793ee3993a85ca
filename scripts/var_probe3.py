"""dense16k: per-pivot host time of several solves (RG_TRACE_BITS=1 on stderr) to locate sporadic slow solves."""
import os, sys
sys.path.insert(0, '.')
import relp_b200, bench
prob = bench.make_problem("dense16k", 0)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=0)
    print(f"SOLVE {k} device_ms {g.device_ms:.0f}", file=sys.stderr, flush=True)
