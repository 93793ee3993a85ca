"""Host-side helpers of the row-sharded engine (one process per GPU).

The carry rows are block-distributed exactly as `rg_load_csc` does it on the device side
(relp_b200/csrc/relp_gpu.cu): rank r owns constraint rows [r*q, min(m, (r+1)*q)), q = ceil(m / world).
"""


def row_block(m, world, rank):
    """(first row, number of rows) of `rank`'s block: asks the engine library itself (`rg_shard_block`, the
    function `rg_load_csc` partitions with), so a device-side partition bug cannot hide behind a restatement."""
    import ctypes as C
    from . import _lib
    first, number = C.c_int32(), C.c_int32()
    rc = _lib.load().rg_shard_block(m, world, rank, C.byref(first), C.byref(number))
    if rc != 0:
        raise ValueError(f"rg_shard_block({m}, {world}, {rank}) failed: {rc}")
    return first.value, number.value


def owner_of_row(m, world, row):
    for rank in range(world):
        lo, n = row_block(m, world, rank)
        if lo <= row < lo + n:
            return rank
    raise ValueError("row out of range")


def share_unique_id(dist, make_id, device=None):
    """Rank 0 creates the 128-byte NCCL id with `make_id()`; everyone receives it through the
    torch.distributed process group (`gloo` on CPU, `nccl` on GPU)."""
    import torch
    rank = dist.get_rank()
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())
