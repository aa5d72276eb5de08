"""world_size-2 `gloo` test of the multi-GPU path's host logic, on CPU.

The data path has no collective: ranks only partition (row bands of one image, or images of a
batch), render independently, and the launcher reduces the elapsed time with MAX.  Here the CPU
oracle stands in for the kernel so the partitioning / halo / gather logic is exercised without a
GPU: every rank renders its band from a source copy in which all rows OUTSIDE the halo reported by
smol_cuda_band_source_rows are destroyed, and the gathered result must equal the whole-image output.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

JOBS = [
    # type_in, w_in, h_in, type_out, w_out, h_out, srgb
    (1, 96, 540, 5, 48, 270, 0),      # 2:1 bilinear, premul -> unassoc
    (0, 64, 1300, 0, 20, 90, 1),      # box x box, linear light
    (8, 40, 70, 8, 130, 301, 0),      # magnify, 24bpp
    (2, 128, 512, 2, 16, 64, 0),      # two halvings
]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    import torch
    import torch.distributed as dist
    import cases
    import oracle
    import smolscale_b200 as sb
    from smolscale_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chk = oracle.restatement()
    ok = True
    for ti, wi, hi, to, wo, ho, srgb in JOBS:
        si, so = wi * cases.bpp(ti), wo * cases.bpp(to)
        src = cases.make_image(ti, wi, hi, si, "random", seed=42)
        whole = chk.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
        # --- row bands: only the halo rows survive in this rank's copy of the source
        first, n = sharding.row_band(ho, rank, world)
        ctx = sb.ScaleCtx(src, ti, wi, hi, si, None, to, wo, ho, so, srgb)     # host planning only
        r0, nr = ctx.band_source_rows(first, n)
        ctx.destroy()
        mine = np.full_like(src, 0xA5)
        mine[r0 * si:(r0 + nr) * si] = src[r0 * si:(r0 + nr) * si]
        band = chk.scale_rows(mine, ti, wi, hi, si, to, wo, ho, first, n, so, srgb)
        gathered = [torch.zeros(sharding.row_band(ho, r, world)[1] * so, dtype=torch.uint8) for r in range(world)]
        # ragged all_gather via per-rank broadcast (no data-path collective exists in the product;
        # this is test plumbing only)
        for r in range(world):
            t = torch.from_numpy(band.copy()) if r == rank else gathered[r]
            dist.broadcast(t, src=r)
            gathered[r] = t
        full = np.concatenate([g.numpy() for g in gathered])
        ok = ok and np.array_equal(full, whole)
        # --- image shards cover the batch exactly once
        cover = torch.zeros(37, dtype=torch.int32)
        f, c = sharding.image_shard(37, rank, world)
        cover[f:f + c] += 1
        dist.all_reduce(cover)
        ok = ok and bool((cover == 1).all())
    # --- timing reduction used by bench.py: MAX over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()
    results[rank] = ok


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    results = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(results.get(r) for r in range(world)), dict(results)


def test_shard_helpers():
    from smolscale_b200 import sharding
    for n in (1, 7, 64, 4096):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                f, c = sharding.image_shard(n, r, world)
                seen += list(range(f, f + c))
            assert seen == list(range(n))
            rows = []
            for r in range(world):
                f, c = sharding.row_band(n, r, world)
                rows += list(range(f, f + c))
            assert rows == list(range(n))
