"""Host-side logic of the multi-GPU path on CPU with gloo, world_size 2: shard arithmetic,
the single broadcast that carries rank 0's parameters/weights, and batch-order gathering of per-rank outputs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from feathercnn_b200 import dist as fdist
        text = fdist.broadcast_bytes(b"7767517\n1 1\nInput data 0 1 data\n" if rank == 0 else None, 0)
        # the weight "arena": rank 0 owns the values, everyone allocates the same size and receives them
        arena = torch.arange(1000, dtype=torch.float32) if rank == 0 else torch.zeros(1000)
        dist.broadcast(arena, 0)
        total = 7
        b, e = fdist.shard_range(total, rank, world)
        local = np.arange(b, e, dtype=np.float32)[:, None] * np.ones((1, 3), np.float32)
        out = fdist.gather_outputs(local, total)
        q.put((rank, text, float(arena.sum()), (b, e), None if out is None else out[:, 0].tolist()))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from feathercnn_b200.dist import shard_range
    for total in (1, 7, 64, 1024, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(1024, 3, 8) == (384, 512)  # BASELINE.json configs[4]: 1024 images over 8 GPUs


def test_net_group_shards_like_the_process_per_gpu_path():
    """feather::NetGroup (one process, N devices) splits a batch exactly like dist.shard_range (one process per device)."""
    import ctypes
    from feathercnn_b200._lib import feather
    from feathercnn_b200.dist import shard_range
    lib = feather()
    lo, hi = ctypes.c_int(), ctypes.c_int()
    for total in (0, 1, 7, 64, 1000, 1024):
        for world in (1, 2, 3, 8):
            for r in range(world):
                assert lib.fgroup_shard_range(total, world, r, ctypes.byref(lo), ctypes.byref(hi)) == 0
                assert (lo.value, hi.value) == shard_range(total, r, world)
    assert lib.fgroup_shard_range(8, 2, 2, ctypes.byref(lo), ctypes.byref(hi)) == -1   # member out of range
    g = lib.fgroup_create()
    try:
        assert lib.fgroup_size(g) == 0 and lib.fgroup_broadcast_transport(g) == b""
        assert lib.fgroup_forward_batch(g, None, 1, None, None) == -1                  # not initialised
        import torch
        if not torch.cuda.is_available():   # no device: a clean error code, not a crash (the product has no CPU path)
            assert lib.fgroup_init_from_path(g, b"/nonexistent/model", None, 0) == -700
            assert lib.fgroup_size(g) == 0
    finally:
        lib.fgroup_destroy(g)


def test_two_rank_broadcast_and_gather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, t0, s0, span0, out0), (r1, t1, s1, span1, out1) = res
    assert t0 == t1 and t0.startswith(b"7767517")
    assert s0 == s1 == float(sum(range(1000)))
    assert span0 == (0, 4) and span1 == (4, 7)
    assert out0 == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0] and out1 is None
