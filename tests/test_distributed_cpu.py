"""Multi-rank host logic on CPU: gloo, world_size 2 (sharding + the single stats all-gather)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphik_b200.distributed import STAT_FIELDS, gather_stats, reduce_stats, shard_bounds, summary_stats


def test_shard_bounds_partition():
    for B in (0, 1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [h - l for l, h in cuts]
            assert max(sizes) - min(sizes) <= 1


def _fake_output(lo, hi):
    idx = torch.arange(lo, hi)
    return {"iterations": (50 + idx % 7).to(torch.int32), "status": (idx % 10 == 0).to(torch.int32),
            "f(x)": 1e-15 * (1 + idx.double()), "n_inner": (1000 + idx).to(torch.int32)}


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(B, world, rank)
    per_rank, totals = gather_stats(summary_stats(_fake_output(lo, hi), device_ms=10.0 + rank))
    q.put((rank, per_rank.tolist(), totals))
    dist.destroy_process_group()


def test_stats_all_gather_gloo_world2():
    B, world, port = 101, 2, 29541
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = reduce_stats(summary_stats(_fake_output(0, B), device_ms=11.0).numpy()[None])
    for rank, per_rank, totals in got:
        assert len(per_rank) == world and len(per_rank[0]) == len(STAT_FIELDS)
        for key in ("count", "converged", "sum_outer", "sum_inner", "max_outer"):
            assert totals[key] == single[key], key
        assert abs(totals["sum_f"] - single["sum_f"]) <= 1e-25 and totals["max_f"] == single["max_f"]
        assert totals["device_ms"] == 11.0   # max over ranks
