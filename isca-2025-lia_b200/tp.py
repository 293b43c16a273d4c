"""Tensor-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch) for bootstrap and the all-reduce that follows each row-parallel projection.

Replaces the reference's GPU GEMM -> .to('cpu') -> deepspeed_comm.all_reduce (oneCCL) ->
.to('cuda') bounce (decoder.py:60-77) with an in-place device all-reduce on the compute
stream.  Sharding rule: weights.shard_layer (tensor_parallel.py:30-141).
"""
import os

import torch
import torch.distributed as dist

_group = None


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's env (RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world).  A world of 1 needs no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def set_group(group):
    global _group
    _group = group


def world_size():
    return dist.get_world_size(_group) if dist.is_initialized() else 1


def all_reduce(t):
    """In-place sum over the tensor-parallel group (bf16, like the reference's message dtype)."""
    if dist.is_initialized() and dist.get_world_size(_group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)
    return t


def barrier():
    if dist.is_initialized():
        dist.barrier(group=_group)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing rule: device time, max over ranks)."""
    if not dist.is_initialized() or dist.get_world_size(_group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_group)
    return float(t.item())
