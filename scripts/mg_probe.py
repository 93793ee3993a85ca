"""Multi-GPU timing probe (torchrun): loop time vs the sum of the per-phase CUDA-event times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import relp_b200, bench
from relp_b200.solver import nccl_unique_id

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("gloo")

def share_id():
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8)
    dist.broadcast(t, 0)
    return bytes(t.tolist())

for name in sys.argv[1:] or ["sparse4k"]:
    prob = bench.make_problem(name, 0)
    for prof in (0, 0, 2):
        nid = share_id()
        dist.barrier()
        g = relp_b200.solve_relaxation(prob, rule="steepest_edge", device=local, profile=prof, rank=rank, world=world, nccl_id=nid)
        if rank == 0:
            print(f"{name} world={world} profile={prof} pivots={g.pivots} loop {g.seconds*1e3:.1f} ms -> {g.pivots/g.seconds:.1f} pivots/s",
                  "phase sum", round(sum(g.stats["phase_ms"][:6]), 1), [round(x, 1) for x in g.stats["phase_ms"]], flush=True)
dist.destroy_process_group()
