"""Multi-GPU plumbing: instances are independent, so they are sharded by global instance id with no
data-path collective; the only exchange is one all-reduce of the per-dispatcher production statistics
(SURVEY.md section 8e).  One process per GPU (torchrun), ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests)."""
import numpy as np

STATS_FIELDS = ["count", "tons_hi", "tons_lo", "sum_events", "sum_decisions", "sum_jams", "sum_road_repairs",
                "sum_truck_repairs", "sum_unrepairable", "n_err", "sum_hops", "dumped_hi", "dumped_lo", "tons_sq_hi", "tons_sq_lo"]
N_DISPATCHERS, STATS_WIDTH = 7, 16
DISPATCHER_NAMES = ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher", "SPTFDispatcher",
                    "SQDispatcher", "ShortestTripDispatcher"]


def shard_range(n_total, world_size, rank):
    """Global instance ids [lo, hi) owned by `rank`: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _fixed_point(tons, shift):
    """llrint(tons * 2^shift) as uint64, clamped the way om_stats_kernel clamps (csrc/om_engine.cu)."""
    x = np.asarray(tons, dtype=np.float64) * float(1 << shift)
    x = np.where(x > 0.0, x, 0.0)                       # also NaN -> 0
    x = np.minimum(x, 1.0e18 if shift == 24 else 4.0e9)
    return np.rint(x).astype(np.uint64)


def stats_from_summaries(summary, dispatcher_ids):
    """Host mirror of ``om_stats_kernel`` (csrc/om_engine.cu): [7, 16] float64, every entry an integer.  Tonnages are
    accumulated in fixed point (2^-24 t, split into two words at bit 25; squares of tons rounded to 2^-8 t, split at bit 32)
    so that the sums are exact: the same instances give the same statistics in any order and on any number of GPUs."""
    out = np.zeros((N_DISPATCHERS, STATS_WIDTH), dtype=np.float64)
    ids = np.asarray(dispatcher_ids)
    for d in range(N_DISPATCHERS):
        s = summary[ids == d]
        if len(s) == 0:
            continue
        q = _fixed_point(s["reported_tons"], 24)
        qd = _fixed_point(s["dumped_tons"], 24)
        q8 = _fixed_point(s["reported_tons"], 8)
        sq = q8 * q8
        cols = [len(s), (q >> np.uint64(25)).sum(), (q & np.uint64(0x1FFFFFF)).sum(), s["n_truck_events"].sum(dtype=np.uint64),
                s["n_decisions"].sum(dtype=np.uint64), s["n_jams"].sum(dtype=np.uint64), s["n_road_repairs"].sum(dtype=np.uint64),
                s["n_truck_repairs"].sum(dtype=np.uint64), s["n_unrepairable"].sum(dtype=np.uint64),
                ((s["err"] != 0) & (s["err"] != 5)).sum(), s["n_hops"].sum(dtype=np.uint64),
                (qd >> np.uint64(25)).sum(), (qd & np.uint64(0x1FFFFFF)).sum(),
                (sq >> np.uint64(32)).sum(), (sq & np.uint64(0xFFFFFFFF)).sum()]
        out[d, :len(cols)] = [float(int(c)) for c in cols]
    return out


def allreduce_stats(stats, group=None):
    """Sum the [7,16] statistics tensor over all ranks (the path's only collective).  Every entry is an integer below 2^53,
    so the fp64 sum is exact whatever order the collective adds the ranks in."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def decode_tons(hi, lo):
    """Exact fixed-point total (hi * 2^25 + lo) * 2^-24 tons -> float."""
    return float((int(hi) << 25) + int(lo)) / float(1 << 24)


def describe(stats):
    """Per-dispatcher mean / std of reported tons etc. from (all-reduced) statistics."""
    stats = np.asarray(stats, dtype=np.float64)
    rows = {}
    for d, name in enumerate(DISPATCHER_NAMES):
        n = stats[d, 0]
        if n == 0:
            continue
        total = decode_tons(stats[d, 1], stats[d, 2])
        total_sq = float((int(stats[d, 13]) << 32) + int(stats[d, 14])) / 65536.0
        mean = total / n
        var = max(total_sq / n - mean * mean, 0.0)
        rows[name] = {"instances": int(n), "mean_tons": mean, "std_tons": var ** 0.5, "sum_tons": total,
                      "sum_dumped_tons": decode_tons(stats[d, 11], stats[d, 12]),
                      "mean_events": stats[d, 3] / n, "mean_decisions": stats[d, 4] / n, "mean_jams": stats[d, 5] / n,
                      "mean_road_repairs": stats[d, 6] / n, "mean_truck_repairs": stats[d, 7] / n,
                      "mean_unrepairable": stats[d, 8] / n, "errors": int(stats[d, 9])}
    return rows
