"""Bit-level parity of the PHILOX-mode kernels -- the variant bench.py times.

In Philox mode the engine draws from counter-based streams, merges draws (one call serves an order's road-repair gap and
the next move's unrepairable gap; one call serves three trucks' breakdown gaps; one call serves 64 roads' jam coins) and
samples the jam position by rejection, so its trajectories cannot be compared with a reference run seeded the usual way.
Instead the TRACE variants export the DRAW TAPE: every value the reference would have drawn, post-transform, in the
reference's call order (OmOutputs.draw_values / draw_sites).  The CPU restatement then REPLAYS that tape
(oracle.des_oracle.run_oracle(..., "replay")) and must reproduce the device trajectory bit for bit -- event trace, tick
arrays, tonnage, order lists, final per-truck / per-unit values -- and consume the tape to its last entry.  This pins the
control flow of the Philox kernels (which draw is made when, in what order, and what it does); only the marginal law of
each draw is left to the statistical tests (tests/test_gpu_rng_marginals.py).

Reference behaviour replayed: truck.py:108-251,281-457,522-547, road.py:137-331, mine.py:50-155 (through the oracle,
which is pinned to the unmodified reference by tests/golden).
"""
import numpy as np
import pytest

from helpers import HostResult, engine_vs_ref, oracle_replays

pytestmark = pytest.mark.gpu


def _check(cfg, names, seeds, trace_capacity, order_capacity, min_draw_kinds, **mine_kw):
    import openmines_b200 as om
    n = len(names)
    mine = om.BatchedMine(cfg, n, **mine_kw)
    mine.reset(names, seeds=seeds, rng_mode="philox")
    res = mine.run(trace_capacity=trace_capacity, site_ticks=True, order_capacity=order_capacity,
                   draw_capacity=mine.draw_tape_slots())
    host = HostResult(mine, res)
    s = host.summary_np()
    assert (s["err"] == 0).all(), np.unique(s["err"], return_counts=True)
    tapes = [host.draw_tape(i) for i in range(n)]
    oracles = oracle_replays([(cfg, names[i], tapes[i], {}) for i in range(n)])
    seen = set()
    for i in range(n):
        o = oracles[i]
        assert o.error is None, (i, o.error)
        diffs = engine_vs_ref(o, mine, host, i, sptf=(names[i] == "SPTFDispatcher"))
        assert not diffs, "instance %d (%s, seed %d): %s" % (i, names[i], seeds[i], diffs[:5])
        # the replay consumed the whole tape, entry for entry
        assert np.array_equal(o.tape[0], tapes[i][0]) and np.array_equal(o.tape[1], tapes[i][1]), i
        seen |= set(np.unique(tapes[i][0]).tolist())
    assert seen >= set(min_draw_kinds), "draw sites exercised: %s" % sorted(seen)
    return s


def test_philox_draw_tape_replay_north_pit(north_pit):
    """256 seeds x {SQ, SPTF} (the benchmarked workload) + 32 seeds of each other dispatcher on north_pit_mine."""
    names = ["SQDispatcher", "SPTFDispatcher"] * 256
    seeds = np.repeat(np.arange(256, dtype=np.int64) + 500, 2)
    others = ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher"] * 32
    names += others
    seeds = np.concatenate([seeds, np.repeat(np.arange(32, dtype=np.int64) + 9000, 4)])
    # sites 0,1,3,4,5,6,9,10,11 fire in every episode; 12 = RandomDispatcher's randint
    _check(north_pit, names, seeds, 8192, 2048, {0, 1, 3, 4, 5, 6, 9, 10, 11, 12})


def test_philox_draw_tape_replay_ragged_mine(ragged_mine):
    """256 seeds x {SQ, SPTF} on the ragged mine (L != D, 1..9 units per site, 130 trucks, frequent road repairs and
    detours: draw sites 7 and 8 fire dozens of times per episode)."""
    names = ["SQDispatcher", "SPTFDispatcher"] * 256
    seeds = np.repeat(np.arange(256, dtype=np.int64) + 77, 2)
    _check(ragged_mine, names, seeds, 8192, 4096, {0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11})


def test_philox_draw_tape_replay_scaled_mine():
    """256 seeds x {SQ, SPTF} on the 500-truck mine of BASELINE configs[4] (SPLIT kernels: truck records and jam pool in
    HBM, cached slot minima, shared-memory helper arrays)."""
    from openmines_b200.synthetic import scaled_mine
    cfg = scaled_mine()
    names = ["SQDispatcher", "SPTFDispatcher"] * 256
    seeds = np.repeat(np.arange(256, dtype=np.int64) + 3, 2)
    _check(cfg, names, seeds, 32768, 16384, {0, 1, 3, 4, 5, 6, 9, 10, 11}, jam_capacity=512)


def test_draw_tape_replays_on_the_device_too(north_pit):
    """The exported tape fed back through OM_RNG_TAPE gives the same trajectory as the Philox run it came from (the two RNG
    front ends of the engine agree), and a run split in two appends to the same tape."""
    import openmines_b200 as om
    names = ["SQDispatcher", "SPTFDispatcher", "RandomDispatcher", "FixedGroupDispatcher"]
    seeds = np.arange(4) + 31
    mine = om.BatchedMine(north_pit, 4)
    mine.reset(names, seeds=seeds, rng_mode="philox")
    res = mine.run(trace_capacity=8192, site_ticks=True, draw_capacity=mine.draw_tape_slots())
    tapes = [res.draw_tape(i) for i in range(4)]
    a = res.summary_np().copy()
    traces = [res.trace_np(i)[0].copy() for i in range(4)]
    # resumed run: same tape
    mine.reset(names, seeds=seeds, rng_mode="philox")
    r2 = mine.run(until=100.5, trace_capacity=8192, draw_capacity=mine.draw_tape_slots())
    r2 = mine.run(result=r2)
    for i in range(4):
        s2, v2 = r2.draw_tape(i)
        assert np.array_equal(s2, tapes[i][0]) and np.array_equal(v2, tapes[i][1])
    # tape mode
    mine.reset(names, rng_mode="tape", tapes=tapes)
    r3 = mine.run(trace_capacity=8192, site_ticks=True)
    b = r3.summary_np()
    for f in ("reported_tons", "dumped_tons", "n_truck_events", "n_decisions", "n_jams", "n_road_repairs",
              "n_truck_repairs", "n_unrepairable", "err"):
        assert np.array_equal(a[f], b[f]), f
    for i in range(4):
        assert np.array_equal(traces[i], r3.trace_np(i)[0])
    # asking for the tape outside Philox mode is refused
    with pytest.raises(RuntimeError, match="OM_RNG_PHILOX"):
        mine.run(draw_capacity=1024)
