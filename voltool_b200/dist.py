"""Multi-GPU plumbing for the one sharded path: fit's pose search.

Rank r evaluates candidates r, r+world, ... of every rotate() table; the per-stage tables are merged
with ONE small all_gather (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests) and
every rank replays the reference's sequential skip rule on the identical merged table, so all ranks
commit the same pose.  Nothing else in the library communicates.
"""
import ctypes as C

import numpy as np

from . import api


def table_merge_allgather(table, rank, world, device=None):
    """interleaved shards -> complete table on every rank (in place).  table: float32 numpy array."""
    import torch
    import torch.distributed as dist
    n = table.size
    per = (n + world - 1) // world
    mine = np.full(per, np.nan, np.float32)
    own = table[rank::world]
    mine[:own.size] = own
    t = torch.from_numpy(mine)
    if device is not None:
        t = t.to(device)
    out = torch.empty(per * world, dtype=torch.float32, device=t.device)
    dist.all_gather_into_tensor(out, t)
    full = out.cpu().numpy().reshape(world, per)
    for r in range(world):
        cnt = table[r::world].size
        table[r::world] = full[r, :cnt]
    return table


_keep = []


def init_nccl_from_torch(device=None):
    """bring up the LIBRARY's NCCL communicator (vt_nccl_init) for the ranks of the current torch.distributed job:
    rank 0 makes the unique id, torch carries it (one 128-byte broadcast), every rank joins.  After this vt_fit
    shards by itself and merges its per-stage tables with one ncclAllGather on device buffers."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    backend = dist.get_backend()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else "cpu")
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8).clone()
    t = t.to(dev)
    dist.broadcast(t, 0)
    api.nccl_init(rank, world, bytes(t.cpu().numpy().tobytes()))
    return rank, world



def install_shard(rank=None, world=None, device=None):
    """register this process as shard `rank` of `world` with the C library (vt_fit_set_shard)"""
    import torch.distributed as dist
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world

    def _merge(ptr, n, r, w, user):
        try:
            table = np.ctypeslib.as_array(ptr, shape=(n,))
            table_merge_allgather(table, r, w, device)
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return 9
    cb = api.MERGE_FN(_merge)
    _keep.append(cb)
    rc = api.lib().vt_fit_set_shard(rank, world, cb, None)
    if rc:
        raise api.VoltoolGpuError("vt_fit_set_shard failed")
    return rank, world


def clear_shard():
    api.lib().vt_fit_set_shard(0, 1, api.MERGE_FN(0), None)
