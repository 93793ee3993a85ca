"""World-size-2 gloo tests (CPU) of the host-side logic of the row-sharded path: the block row
partition the device side uses, and the NCCL-id plumbing over torch.distributed."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relp_b200.sharding import owner_of_row, row_block, share_unique_id


def test_row_blocks_partition_every_row():
    for m in (1, 2, 3, 7, 27, 205, 4096, 16384, 9998):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, n = row_block(m, world, r)
                assert n >= 0
                seen += list(range(lo, lo + n))
                for row in range(lo, lo + n):
                    assert owner_of_row(m, world, row) == r
            assert seen == list(range(m))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw = share_unique_id(dist, lambda: bytes((7 * i + 3) % 256 for i in range(128)))
    # every rank assembles the distributed vector from the blocks, as rg_get_b does after its all-gather
    m = 11
    lo, n = row_block(m, world, rank)
    q = -(-m // world)
    mine = torch.zeros(q, dtype=torch.int64)
    mine[:n] = torch.arange(lo, lo + n) * 10
    gathered = [torch.zeros(q, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)
    full = []
    for r in range(world):
        _, nr = row_block(m, world, r)
        full += gathered[r][:nr].tolist()
    ok = raw == bytes((7 * i + 3) % 256 for i in range(128)) and full == [10 * i for i in range(m)]
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(t.item()))
    dist.destroy_process_group()


def test_gloo_world2_id_broadcast_and_block_assembly():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29731, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
