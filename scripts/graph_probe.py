"""Wall time of the fused loop with / without CUDA graphs and with / without profiling events."""
import os, sys, time
sys.path.insert(0, '.')
import relp_b200
from relp_b200.generators import bounded_lp
prob = bounded_lp(4096, 8192, k_bounding=90, nnz_per_col=8, seed=0)
for prof in (False, True):
    for nog in (False, True):
        if nog: os.environ["RG_NO_GRAPH"] = "1"
        else: os.environ.pop("RG_NO_GRAPH", None)
        best = 1e9
        allt = []
        for rep in range(4):
            g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=prof)
            best = min(best, g.seconds)
            allt.append(round(g.seconds * 1e3, 1))
        print(f"profile={prof} graphs={not nog} pivots={g.pivots} loop best {best*1e3:.1f} ms -> {g.pivots/best:.0f} pivots/s", "limbs histogram", g.stats["pivots_at_limbs"], "promotions", g.stats["promotions"], allt, flush=True)
