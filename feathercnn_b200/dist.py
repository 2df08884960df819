"""Multi-GPU plumbing: one process per GPU, batch sharding, one weight broadcast at load, no collective in Forward.

The reference has no distributed code at all; CNN inference over a batch is embarrassingly
parallel (§8e), so the only exchange is getting rank 0's weights to every GPU once.  ``torch.distributed`` (NCCL
over NVLink/NVSwitch on the GPU box, gloo in CPU tests) is used purely as that transport:

    rank 0:  Net.LoadParam + Net.LoadWeights (file -> host staging -> device weight arena)
    rank r:  Net.LoadParamFromText(param text broadcast from rank 0) + Net.PrepareWeightArena()
    all:     dist.broadcast(arena)  ->  rank r: Net.AttachWeights()
    Forward: each rank runs its contiguous shard of the batch; outputs stay per-GPU or are gathered by the caller.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of `total` images owned by `rank`; the first total % world ranks get one extra."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_bytes(data: bytes | None, src: int = 0, device: str | torch.device = "cpu") -> bytes:
    """Broadcast a byte string from `src` (other ranks pass None)."""
    rank = dist.get_rank()
    n = torch.tensor([len(data) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    if rank == src:
        buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


class _DeviceBuffer:
    """CUDA array interface view of a raw device pointer so torch can wrap the weight arena without a copy."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def arena_tensor(net, device: int) -> torch.Tensor:
    ptr, n = net.WeightArena()
    if n == 0:
        return torch.empty(0, device=f"cuda:{device}")
    return torch.as_tensor(_DeviceBuffer(ptr, n), device=f"cuda:{device}")


def load_net_distributed(param_path, bin_path, device: int, **net_kwargs):
    """Builds the same Net on every rank with ONE broadcast of the weights (NCCL when the backend is nccl)."""
    from .net import Net

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    net = Net(**net_kwargs)
    if world == 1:
        net.LoadParam(param_path)
        net.LoadWeights(bin_path)
        return net
    dev = torch.device(f"cuda:{device}")
    text = broadcast_bytes(Path(param_path).read_bytes() if rank == 0 else None, 0, dev)
    if rank == 0:
        net.LoadParam(param_path)
        net.LoadWeights(bin_path)
    else:
        net.LoadParamFromText(text)
        net.PrepareWeightArena()
    arena = arena_tensor(net, device)
    dist.broadcast(arena, 0)  # the single collective of the whole job
    torch.cuda.synchronize(dev)
    if rank != 0:
        net.AttachWeights()
    return net


def gather_outputs(local: np.ndarray, total: int) -> np.ndarray | None:
    """Concatenate per-rank outputs in batch order on rank 0 (host side; Forward itself has no collective)."""
    world = dist.get_world_size()
    parts: list = [None] * world
    dist.all_gather_object(parts, local)
    if dist.get_rank() != 0:
        return None
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == total
    return out
