"""Row-sharded parity check: run under torchrun (one process per GPU).  Every rank solves the same
problems with world = WORLD_SIZE and compares trace / objective / solution with the CPU oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import relp_b200
from relp_b200.generators import bounded_lp, max_flow
from relp_b200.solver import nccl_unique_id
from relp_b200.sharding import share_unique_id


def share_id(rank):
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8)
    dist.broadcast(t, 0)
    return bytes(t.tolist())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    from oracle import fast_oracle as fo
    cases = [
        ("bounded 40x60", bounded_lp(40, 60, k_bounding=12, nnz_per_col=4, seed=0), ["steepest_edge", "dantzig"], 1),
        ("bounded dense 30x40", bounded_lp(30, 40, k_bounding=10, dense=True, seed=1), ["steepest_edge"], 1),
        ("max flow 24", max_flow(24, 3, 3, 9), ["steepest_edge", "first_profitable"], 2),
        ("bounded 300x600", bounded_lp(300, 600, k_bounding=40, nnz_per_col=6, seed=2), ["steepest_edge"], 2),
        ("dense block 200x120", bounded_lp(200, 120, k_bounding=20, dense=True, seed=3, dense_block=True),
         ["steepest_edge"], 1),
        ("limb sweep 256x512", bounded_lp(256, 512, k_bounding=160, dense=True, seed=2, dense_block=True),
         ["steepest_edge"], 1),
    ]
    try:
        from tests.test_gpu_parity import random_matrix_data
        from tests.common import problem_from_provider
        import numpy as np
        for seed in range(4):
            rng = np.random.default_rng(3000 + seed)
            md = random_matrix_data(rng, 6, (2, 1, 2, 1))
            cases.append((f"random md {seed}", problem_from_provider(md), ["steepest_edge", "dantzig"], 2))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            print("random cases skipped:", e)
    ok = True
    for name, prob, rules, limbs in cases:
        for rule in rules:
            ref = fo.solve_problem(prob, rule)
            for fused, dense_carry in ((True, False), (True, True), (False, False)):
                nid = share_id(rank)
                g = relp_b200.solve_relaxation(prob, rule=rule, fused=fused, initial_limbs=limbs, device=local,
                                               rank=rank, world=world, nccl_id=nid, dense_carry=dense_carry)
                good = (g.status == ref.status and g.trace == ref.trace and
                        (ref.status != "optimal" or (g.objective == ref.objective and g.bfs == ref.bfs)))
                ok = ok and good
                if rank == 0:
                    print(f"{name:22s} {rule:18s} fused={fused} dense={dense_carry} world={world} pivots={g.pivots} "
                          f"limbs={g.stats['limbs']} {'OK' if good else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if flag.item() != 1:
        sys.exit(1)
    if rank == 0:
        print("MGPU PARITY OK")


if __name__ == "__main__":
    main()
