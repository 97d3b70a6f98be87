"""Multi-GPU plumbing for the sharded argmax: one process per GPU (torchrun), weights
replicated, start points sharded, ONE max all-reduce on a packed int64 key to agree on the
winner, plus one small broadcast to fetch its record.

NCCL has ncclMax but no MAXLOC, so value and location travel in one integer
(``bore_select_best``): key = (orderable(-fun) << 31) | (0x7fffffff - global_index); the
largest key is the smallest ``fun``, ties resolved to the lowest global index -- the reference's
first-minimum rule (bore/mixins.py:86).  Backend-agnostic (``nccl`` on GPUs, ``gloo`` in the CPU
tests); torch.distributed is plumbing only.
"""
import os

import numpy as np

KEY_INDEX_MASK = 0x7FFFFFFF


def env_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world == 1:
        return rank, local_rank, world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def decode_key(key):
    """key -> global start index, or None when no start qualified anywhere."""
    key = int(key)
    if key == 0:
        return None
    return KEY_INDEX_MASK - (key & KEY_INDEX_MASK)


def shard_bounds(total, rank, world):
    """Contiguous shard [lo, hi) of ``total`` items for ``rank`` (remainder to the low ranks)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(global_index, total, world):
    for r in range(world):
        lo, hi = shard_bounds(total, r, world)
        if lo <= global_index < hi:
            return r, global_index - lo
    raise IndexError(global_index)


def global_winner(key, record_fn, total, record_len, group=None):
    """Agree on the global winner.

    key        int64 tensor [1] holding this rank's packed key (bore_select_best with
               idx_offset = this rank's shard start); reduced IN PLACE with one MAX all-reduce.
    record_fn  local_index -> float64 tensor [record_len] describing that start (x, fun, ...),
               on the same device as ``key``; called on the owning rank only.
    Returns (global_index or None, record tensor or None) on every rank.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)  # the single max(loc) reduction
    gidx = decode_key(key.item())
    if gidx is None:
        return None, None
    owner, local = owner_of(gidx, total, world)
    if rank == owner:
        rec = record_fn(local).to(torch.float64).reshape(record_len).contiguous()
    else:
        rec = torch.empty(record_len, dtype=torch.float64, device=key.device)
    if world > 1:
        # `owner` is a rank WITHIN `group`; broadcast wants a global rank
        src = dist.get_global_rank(group, owner) if group is not None else owner
        dist.broadcast(rec, src=src, group=group)
    return gidx, rec


def pack_key_numpy(fun, status, idx_offset=0, keep=None):
    """Host mirror of bore_select_best's key (for the CPU/gloo tests of this module only; the
    product computes keys on the device)."""
    best = 0
    for i, (f, st) in enumerate(zip(fun, status)):
        if st not in (0, 1) or (keep is not None and not keep[i]):
            continue
        v = np.float32(-np.float32(f))
        if np.isnan(v):
            continue
        if v == 0:
            v = np.float32(0.0)
        u = int(np.array(v, np.float32).view(np.uint32))
        o = (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
        key = (o << 31) | (KEY_INDEX_MASK - (i + idx_offset))
        best = max(best, key)
    return best
