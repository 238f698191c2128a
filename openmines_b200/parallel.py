"""Multi-GPU plumbing: instances are independent, so they are sharded by global instance id with no
data-path collective; the only exchange is one all-reduce of the per-dispatcher production statistics
(SURVEY.md section 8e).  One process per GPU (torchrun), ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests)."""
import numpy as np

STATS_FIELDS = ["count", "sum_tons", "sum_tons_sq", "sum_events", "sum_decisions", "sum_jams", "sum_road_repairs",
                "sum_truck_repairs", "sum_unrepairable", "n_err", "sum_hops", "sum_dumped_tons"]
N_DISPATCHERS, STATS_WIDTH = 6, 16


def shard_range(n_total, world_size, rank):
    """Global instance ids [lo, hi) owned by `rank`: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stats_from_summaries(summary, dispatcher_ids):
    """Host mirror of ``om_stats_kernel`` (csrc/om_engine.cu): [6, 16] float64 partial sums."""
    out = np.zeros((N_DISPATCHERS, STATS_WIDTH), dtype=np.float64)
    for d in range(N_DISPATCHERS):
        s = summary[np.asarray(dispatcher_ids) == d]
        if len(s) == 0:
            continue
        tons = s["reported_tons"].astype(np.float64)
        out[d, 0] = len(s)
        out[d, 1] = tons.sum()
        out[d, 2] = (tons * tons).sum()
        out[d, 3] = s["n_truck_events"].sum()
        out[d, 4] = s["n_decisions"].sum()
        out[d, 5] = s["n_jams"].sum()
        out[d, 6] = s["n_road_repairs"].sum()
        out[d, 7] = s["n_truck_repairs"].sum()
        out[d, 8] = s["n_unrepairable"].sum()
        out[d, 9] = ((s["err"] != 0) & (s["err"] != 5)).sum()
        out[d, 10] = s["n_hops"].sum()
        out[d, 11] = s["dumped_tons"].sum()
    return out


def allreduce_stats(stats, group=None):
    """Sum the [6,16] statistics tensor over all ranks (the path's only collective)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def describe(stats):
    """Per-dispatcher mean / std of reported tons etc. from (all-reduced) statistics."""
    stats = np.asarray(stats, dtype=np.float64)
    rows = {}
    names = ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher", "SPTFDispatcher",
             "SQDispatcher"]
    for d, name in enumerate(names):
        n = stats[d, 0]
        if n == 0:
            continue
        mean = stats[d, 1] / n
        var = max(stats[d, 2] / n - mean * mean, 0.0)
        rows[name] = {"instances": int(n), "mean_tons": mean, "std_tons": var ** 0.5,
                      "mean_events": stats[d, 3] / n, "mean_decisions": stats[d, 4] / n, "mean_jams": stats[d, 5] / n,
                      "mean_road_repairs": stats[d, 6] / n, "mean_truck_repairs": stats[d, 7] / n,
                      "mean_unrepairable": stats[d, 8] / n, "errors": int(stats[d, 9])}
    return rows
