#!/usr/bin/env python
"""tools/c5_sharded_bench.py — BASELINE config C5 in miniature: a batch of mono "files" of mixed length (uniform in
[1 s, 10 s], seeded), even index -> GC-ADPCM (coefficients + encode), odd index -> CRI ADX (Linear, v4, 18-byte frames),
sharded across the GPUs of one box and gathered back to rank 0.

  python tools/c5_sharded_bench.py --files-per-gpu 1024                                     (one GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/c5_sharded_bench.py --files-per-gpu 1024                                      (N GPUs, one rank each)

What it shows: the path shards with NO data-path collective - vgaudio_b200.sharding.partition() balances
the ragged lengths (greedy longest-first), every rank encodes its own files through the host C ABI - and the only
collective is the gather of the variable-size bitstreams to rank 0 over NCCL (sharding.gather_bitstreams).  Rank 0
checks that every file came back and compares a sample of them with the CPU oracle.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vgaudio_b200 as vg  # noqa: E402
from vgaudio_b200 import sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files-per-gpu", type=int, default=1024)
    ap.add_argument("--min-seconds", type=float, default=1.0)
    ap.add_argument("--max-seconds", type=float, default=10.0)
    ap.add_argument("--check", type=int, default=16, help="files rank 0 compares with the oracle")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    vg._native.check(vg.lib.vgb_init(local, 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_files = a.files_per_gpu * world
    rng = np.random.default_rng(0x5647)
    lengths = rng.integers(int(a.min_seconds * 48000), int(a.max_seconds * 48000) + 1, n_files)
    shards = sharding.partition(lengths, world)
    mine = shards[rank]
    base = synth.batch(16, int(a.max_seconds * 48000), degenerate=False, first_index=20)
    pcm = {i: base[i % 16][: lengths[i]] for i in mine}
    gc_idx = [i for i in mine if i % 2 == 0]
    adx_idx = [i for i in mine if i % 2 == 1]
    cfg = vg.criadx.CriAdxParameters()

    def encode_all():
        out = {}
        if gc_idx:
            coefs, adpcm = vg.gcadpcm.encode_batch([pcm[i] for i in gc_idx])
            for k, i in enumerate(gc_idx):  # a .dsp-like payload: 32 coefficient bytes + the ADPCM stream
                out[i] = np.concatenate([coefs[k].view(np.uint8), adpcm[k]])
        if adx_idx:
            adx, _ = vg.criadx.encode_batch([pcm[i] for i in adx_idx], cfg)
            for k, i in enumerate(adx_idx):
                out[i] = adx[k]
        return out

    encode_all()  # warm-up (allocations, first-use costs)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    items = encode_all()
    torch.cuda.synchronize()
    enc_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([enc_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        enc_ms = float(t.item())
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:
        gathered = sharding.gather_bitstreams(items, n_files, dst=0)
    else:
        gathered = [items[i] for i in range(n_files)]
    torch.cuda.synchronize()
    gather_ms = (time.perf_counter() - t0) * 1e3

    if rank == 0:
        from oracle import pyoracle as oracle
        step = max(1, n_files // max(a.check, 1))
        ok = True
        for i in range(0, n_files, step):
            x = base[i % 16][: lengths[i]]
            if i % 2 == 0:
                co = oracle.calculate_coefficients(x)
                want = np.concatenate([co.view(np.uint8), oracle.encode(x, co)])
            else:
                want, _ = oracle.adx_encode(x, 48000, 18, 4, 0, 3, 0)
            ok = ok and np.array_equal(gathered[i], want)
        total = int(lengths.sum())
        print(json.dumps({
            "workload": "C5 miniature: mixed-length mono files, even -> GC-ADPCM, odd -> CRI ADX (Linear v4)",
            "n_gpus": world, "files": n_files, "samples": total, "bytes_out": int(sum(g.size for g in gathered)),
            "shard_imbalance": round(sharding.imbalance(lengths, shards), 4),
            "encode_ms_max_over_ranks": round(enc_ms, 1), "Msamples_per_s": round(total / enc_ms / 1e3, 1),
            "gather_ms": round(gather_ms, 1), "collective": "NCCL gather of padded shards (sharding.gather_bitstreams)" if world > 1 else None,
            "api": "host C ABI, pageable numpy buffers (H2D/D2H inside the encode time)",
            "oracle_checked_files": len(range(0, n_files, step)), "parity": bool(ok)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
