"""dense16k (config 5): loop wall time with / without CUDA graphs and profiling levels."""
import os, sys, time
sys.path.insert(0, '.')
import relp_b200, bench
prob = bench.make_problem("dense16k", 0)
g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=0)   # warm-up
for prof, nog in ((0, False), (0, True), (2, False), (2, True)):
    if nog: os.environ["RG_NO_GRAPH"] = "1"
    else: os.environ.pop("RG_NO_GRAPH", None)
    g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=prof)
    print(f"profile={prof} graphs={not nog} pivots={g.pivots} loop {g.seconds:.3f} s device {g.device_ms/1e3:.3f} s -> {g.pivots/g.seconds:.1f} pivots/s",
          "phases", [round(x) for x in g.stats["phase_ms"]], flush=True)
    print("   limbs histogram", g.stats["pivots_at_limbs"], "promotions", g.stats["promotions"], "launches", g.stats["kernel_launches"], flush=True)
