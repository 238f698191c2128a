"""world_size-2 gloo test (CPU) of the multi-GPU host logic: disjoint shards by global instance id and
the single statistics all-reduce give the same totals as one process."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openmines_b200 import parallel
from openmines_b200.trace import SUMMARY_DTYPE


def _fake_summaries(lo, hi):
    ids = np.arange(lo, hi)
    s = np.zeros(len(ids), dtype=SUMMARY_DTYPE)
    # tonnages with full 52-bit mantissas (sums of jittered loads look like this): floating-point partial sums of them
    # depend on the order of addition, the fixed-point statistics must not
    rng = np.random.default_rng(12345)
    tons = 8000.0 + 600.0 * rng.standard_normal(1 << 12)
    s["reported_tons"] = tons[ids % len(tons)] + 1e-7 * ids
    s["dumped_tons"] = s["reported_tons"] + 37.0 * rng.random(1 << 12)[ids % len(tons)]
    s["n_truck_events"] = 600 + ids
    s["n_decisions"] = 150 + ids % 7
    s["n_jams"] = ids % 5
    s["n_hops"] = 2000 + ids
    s["err"] = (ids % 11 == 0).astype(np.uint32)
    return s, (ids % 6).astype(np.int32)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_total, world, rank)
    s, d = _fake_summaries(lo, hi)
    st = torch.from_numpy(parallel.stats_from_summaries(s, d))
    parallel.allreduce_stats(st)
    q.put((rank, lo, hi, st.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    n_total, world = 1001, 2
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == n_total      # disjoint cover
    s, d = _fake_summaries(0, n_total)
    want = parallel.stats_from_summaries(s, d)
    for _, _, _, st in got:
        assert np.array_equal(st, want)            # all-reduced 2-rank totals == single-process totals, bit for bit
    assert (want == np.rint(want)).all()           # integer-valued: exact in any order of addition
    rows = parallel.describe(want)
    assert rows["NaiveDispatcher"]["instances"] == len(range(0, n_total, 6))
    t = s["reported_tons"][d == 5]
    assert abs(rows["SQDispatcher"]["mean_tons"] - t.mean()) < 1e-6 and abs(rows["SQDispatcher"]["std_tons"] - t.std()) < 0.05


def test_shard_range_properties():
    for n in (1, 7, 16384, 131072, 100003):
        for w in (1, 2, 3, 4, 8):
            r = [parallel.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
