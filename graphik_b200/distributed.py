"""Multi-GPU execution: goal poses shard over ranks, no exchange during the solve.

Every IK problem is independent, so the batch is cut into contiguous slices, one
per rank (one process per GPU, `torch.distributed`); the only collective is ONE
all-gather of a fixed-size summary-statistics vector per rank at the end (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  Per-problem outputs stay on the
owning GPU unless `gather_outputs` is asked for.  The reference has no
distributed code at all; this is new.
"""
import numpy as np

STAT_FIELDS = ("count", "converged", "sum_outer", "sum_inner", "max_outer", "sum_f", "max_f", "device_ms")


def shard_bounds(B, world_size, rank):
    """Contiguous slice [lo, hi) of a batch of B goals owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def summary_stats(out, device_ms=0.0):
    """Fixed-size stats vector (STAT_FIELDS) of one rank's solve output (device tensors)."""
    import torch
    it = out["iterations"]
    st = out["status"]
    f = out["f(x)"]
    n = it.numel()
    vals = [float(n), float((st == 0).sum()) if n else 0.0, float(it.sum()) if n else 0.0,
            float(out["n_inner"].sum()) if n else 0.0, float(it.max()) if n else 0.0,
            float(f.sum()) if n else 0.0, float(f.max()) if n else 0.0, float(device_ms)]
    return torch.tensor(vals, dtype=torch.float64, device=f.device)


def gather_stats(local_stats, group=None):
    """The single collective of the path: all-gather of the per-rank stats vectors.
    Returns (per_rank[world, len(STAT_FIELDS)], totals dict)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        buf = [torch.empty_like(local_stats) for _ in range(world)]
        dist.all_gather(buf, local_stats, group=group)
        per_rank = torch.stack(buf).cpu().numpy()
    else:
        per_rank = local_stats.detach().cpu().numpy()[None]
    return per_rank, reduce_stats(per_rank)


def reduce_stats(per_rank):
    per_rank = np.asarray(per_rank, dtype=float)
    col = {k: per_rank[:, i] for i, k in enumerate(STAT_FIELDS)}
    return {
        "count": int(col["count"].sum()), "converged": int(col["converged"].sum()),
        "sum_outer": float(col["sum_outer"].sum()), "sum_inner": float(col["sum_inner"].sum()),
        "max_outer": int(col["max_outer"].max()), "sum_f": float(col["sum_f"].sum()),
        "max_f": float(col["max_f"].max()), "device_ms": float(col["device_ms"].max()),
    }


class ShardedBatchIK:
    """Runs BatchIK on this rank's slice of a global goal batch."""

    def __init__(self, graph, params=None, group=None):
        import torch
        import torch.distributed as dist
        from graphik_b200.engine import BatchIK
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.engine = BatchIK(graph, params, device=torch.cuda.current_device())

    def solve(self, T_goals_global, gather_outputs=False):
        import torch
        import torch.distributed as dist
        lo, hi = shard_bounds(len(T_goals_global), self.world, self.rank)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        T_local = self.engine._f64(T_goals_global[lo:hi])
        start.record()
        out = self.engine.solve(T_local, check=False)
        stop.record()
        stop.synchronize()
        per_rank, totals = gather_stats(summary_stats(out, start.elapsed_time(stop)), self.group)
        out["slice"] = (lo, hi)
        out["stats"], out["stats_per_rank"] = totals, per_rank
        if gather_outputs and self.world > 1:
            sizes = [shard_bounds(len(T_goals_global), self.world, r) for r in range(self.world)]
            for key in ("q", "f(x)", "iterations", "status"):
                parts = [torch.empty((h - l,) + tuple(out[key].shape[1:]), dtype=out[key].dtype,
                                     device=out[key].device) for l, h in sizes]
                dist.all_gather(parts, out[key].contiguous(), group=self.group)
                out[key + "_global"] = torch.cat(parts)
        return out
