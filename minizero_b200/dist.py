"""Multi-GPU plumbing of the self-play path: games are independent, so they shard across ranks with NO data-path
collective; the only exchange is one broadcast of the packed weight blob per model load (the reference
re-reads the .pt on every worker instead, actor/actor_group.cpp:227-232) and the max-over-ranks of the timings.

One process per GPU (torchrun); backend nccl on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_games(total_games, world, rank):
    """Games of rank `rank`: actor i belongs to GPU i % world, as actors are striped over networks in
    ActorGroup::createActors (actor/actor_group.cpp:184-186). Returns the global game indices."""
    return list(range(rank, total_games, world))


def broadcast_blob(dist, blob, src=0):
    """Broadcast the packed weights in place. `blob` is a torch uint8 tensor (a view of the engine's device blob under NCCL)."""
    dist.broadcast(blob, src=src)
    return blob


def broadcast_object(dist, obj, src=0):
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def max_over_ranks(dist, values, device="cpu"):
    """Element-wise max of a small list of floats over all ranks (device timings are reported as the max over ranks)."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(dist, values, device="cpu"):
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


class DeviceBlob:
    """`__cuda_array_interface__` view of a raw device pointer so that torch can wrap the engine's weight blob without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
