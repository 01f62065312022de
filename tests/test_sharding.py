"""CPU tests of the multi-GPU host logic: partitioning, and the gather of variable-size bitstreams over
torch.distributed with the gloo backend, world_size 2 (the same code runs over NCCL on the GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_balances_and_covers():
    from vgaudio_b200.sharding import imbalance, partition

    rng = np.random.default_rng(1)
    lengths = rng.integers(48000, 480000, 1000).tolist()   # config C5: 1 s .. 10 s files
    for world in (1, 2, 4, 8):
        shards = partition(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(1000))
        assert imbalance(lengths, shards) < 1.01
        for s in shards:  # longest first inside a rank
            assert all(lengths[a] >= lengths[b] for a, b in zip(s, s[1:]))
    assert partition([], 4) == [[], [], [], []]
    assert partition([5], 2) == [[0], []]


def _worker(rank: int, world: int, port: int, ret):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from vgaudio_b200.sharding import gather_bitstreams, partition

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        lengths = rng.integers(10, 5000, 37).tolist()
        shards = partition(lengths, world)
        # stand-in for the encoder (the CUDA codec needs a GPU): a deterministic byte stream per item
        local = {i: (np.arange(lengths[i] // 3 + 1, dtype=np.int64) * (i + 1) % 251).astype(np.uint8) for i in shards[rank]}
        out = gather_bitstreams(local, len(lengths), dst=0)
        if rank == 0:
            ok = all(np.array_equal(out[i], (np.arange(lengths[i] // 3 + 1, dtype=np.int64) * (i + 1) % 251).astype(np.uint8))
                     for i in range(len(lengths)))
            ret.put(("ok" if ok else "mismatch", len(out)))
        else:
            assert out is None
            ret.put(("none", 0))
    finally:
        dist.destroy_process_group()


def test_gather_bitstreams_gloo_world2():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    results = [ret.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ("ok", 37) in results and ("none", 0) in results
