"""Config 3 at full size (max flow, V = 2000): the LP optimum of the engine against an independent max-flow
value (scipy.sparse.csgraph.maximum_flow on the same arcs), plus timing."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import relp_b200
from relp_b200.generators import max_flow

V = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rule = sys.argv[2] if len(sys.argv) > 2 else "steepest_edge"
prob = max_flow(V, 4, 0)
# rebuild the arc list exactly as the generator does
rng = np.random.default_rng(0)
arcs = []
for frm in range(V):
    if frm == V - 1:
        continue
    targets = set()
    while len(targets) < min(4, V - 2):
        to = int(rng.integers(1, V))
        if to != frm:
            targets.add(to)
    for to in sorted(targets):
        arcs.append((frm, to, int(rng.integers(1, 101))))
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import maximum_flow
cap = csr_matrix(([a[2] for a in arcs], ([a[0] for a in arcs], [a[1] for a in arcs])), shape=(V, V), dtype=np.int32)
ref = maximum_flow(cap, 0, V - 1).flow_value
t = time.time()
g = relp_b200.solve_relaxation(prob, rule=rule)
print(f"config 3: V={V} m={prob.m} n={prob.n} rule={rule} status={g.status} pivots={g.pivots} "
      f"loop {g.seconds:.3f} s -> {g.pivots / g.seconds:.0f} pivots/s, wall {time.time() - t:.2f} s")
print("LP optimum", -g.objective, "scipy max flow", ref, "MATCH" if -g.objective == ref else "MISMATCH",
      "| denominator", g.denominator, "limbs", g.stats["limbs"], "active columns", g.stats.get("active_columns"))
