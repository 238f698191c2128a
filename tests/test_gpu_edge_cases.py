"""GPU parity on the edge-case mines whose goldens come from the unmodified reference
(tests/golden/make_golden_edge.py), plus the engine's own error paths."""
import numpy as np
import pytest

from helpers import DISPATCHERS, Golden, engine_vs_ref

pytestmark = pytest.mark.gpu


def _mine(cfg, n, **kw):
    import openmines_b200 as om
    return om.BatchedMine(cfg, n, **kw)


def test_ragged_mine_det_vs_golden(ragged_mine):
    """L = 3 != D = 2, 1 / 4 / 7 shovels and 1 / 9 dumpers per site, 130 trucks (5 scan slots, ragged last one), four
    truck types incl. a fractional capacity.  RandomDispatcher hits the reference's IndexError (haul orders are drawn
    over len(load_sites)): the engine must stop at the same record with OM_IERR_BAD_ORDER."""
    mine = _mine(ragged_mine, len(DISPATCHERS))
    mine.reset(DISPATCHERS, seeds=[0] * len(DISPATCHERS), rng_mode="det")
    res = mine.run(trace_capacity=4096, site_ticks=True, order_capacity=2048)
    summ = res.summary_np()
    for i, d in enumerate(DISPATCHERS):
        g = Golden("det_ragged_mine_%s_s0" % d)
        diffs = engine_vs_ref(g, mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
        assert int(summ[i]["err"]) == (2 if g.error else 0), (d, summ[i]["err"], g.error)


def test_ragged_mine_tape_vs_golden(ragged_mine):
    names = ["SQDispatcher", "SPTFDispatcher", "RandomDispatcher"]
    goldens = [Golden("live_ragged_mine_%s_s5" % d) for d in names]
    mine = _mine(ragged_mine, len(names))
    mine.reset(names, rng_mode="tape", tapes=[g.tape for g in goldens])
    res = mine.run(trace_capacity=4096, site_ticks=True, order_capacity=2048)
    for i, d in enumerate(names):
        diffs = engine_vs_ref(goldens[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
        assert int(res.summary_np()[i]["err"]) == (2 if goldens[i].error else 0)


def test_one_truck_mine_vs_golden(one_truck_mine):
    """The smallest legal mine: one truck, one shovel, one dumper (2 + 8 event records per cycle)."""
    g = Golden("det_one_truck_mine_NaiveDispatcher_s0")
    mine = _mine(one_truck_mine, 1)
    mine.reset("NaiveDispatcher", rng_mode="det")
    res = mine.run(trace_capacity=256, site_ticks=True, order_capacity=64)
    assert not engine_vs_ref(g, mine, res, 0)
    g = Golden("live_one_truck_mine_NaiveDispatcher_s2")
    mine = _mine(one_truck_mine, 1)
    mine.reset("NaiveDispatcher", rng_mode="tape", tapes=[g.tape])
    res = mine.run(trace_capacity=256, site_ticks=True, order_capacity=64)
    assert not engine_vs_ref(g, mine, res, 0)


def test_ragged_mine_live_runs_clean(ragged_mine):
    """Philox mode on the ragged mine: no engine errors for the dispatchers that are valid on it, plausible output."""
    names = ["NaiveDispatcher", "NearestDispatcher", "FixedGroupDispatcher", "SPTFDispatcher", "SQDispatcher"]
    n = 64
    ids = np.repeat(np.arange(len(names)), n)
    mine = _mine(ragged_mine, len(ids))
    mine.reset([names[k] for k in ids], seeds=np.tile(np.arange(n), len(names)), rng_mode="philox")
    s = mine.run().summary_np()
    assert (s["err"] == 0).all()
    tons = s["reported_tons"].reshape(len(names), n).mean(axis=1)
    assert tons[0] < tons[1:].min()                    # Naive sends everything to one single-shovel site
    assert (s["n_truck_events"] > 100).all()


def test_engine_error_paths(north_pit):
    """Truncated tape -> OM_IERR_TAPE; a trace buffer that is too small -> OM_IERR_TRACE_OVERFLOW with the simulation
    itself unaffected; a jam pool that is too small -> OM_IERR_JAM_OVERFLOW (never silent corruption)."""
    g = Golden("live_north_pit_mine_SQDispatcher_s11")
    sites, values = g.tape
    mine = _mine(north_pit, 2)
    mine.reset(["SQDispatcher", "SQDispatcher"], rng_mode="tape", tapes=[(sites, values), (sites[:5000], values[:5000])])
    res = mine.run(trace_capacity=64)
    s = res.summary_np()
    assert s["err"][0] == 5 and s["err"][1] == 4       # trace overflow only | tape exhausted
    assert float(s["reported_tons"][0]) == g.reported_tons and int(s["n_truck_events"][0]) == g.n_truck_events
    assert s["now"][1] < s["now"][0]
    small = _mine(north_pit, 32, jam_capacity=2)
    small.reset("SQDispatcher", seeds=np.arange(32), rng_mode="philox")
    e = small.run().summary_np()["err"]
    assert set(np.unique(e)).issubset({0, 3}) and (e == 3).any()


def test_integer_mine_tie_stress_vs_golden(integer_mine):
    """Whole-minute durations everywhere: truck hops, Release events and the monitor block tie at integer timestamps
    all the time, so every pop exercises the (time, priority, seq) tie-break and no chain may be fused wrongly.
    Also: an SQ order at an integer minute always finds that minute's queue sample (the process-return hop that precedes
    it gets a fresh eid behind the monitors'), so OM_IERR_SQ_KEYERROR must not fire."""
    mine = _mine(integer_mine, len(DISPATCHERS))
    mine.reset(DISPATCHERS, seeds=[0] * len(DISPATCHERS), rng_mode="det")
    res = mine.run(trace_capacity=2048, site_ticks=True, order_capacity=1024)
    for i, d in enumerate(DISPATCHERS):
        diffs = engine_vs_ref(Golden("det_integer_mine_%s_s0" % d), mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
    assert (res.summary_np()["err"] == 0).all()


def test_ragged_mine_live_distribution_vs_reference(ragged_mine):
    """Live RNG on the ragged mine, whose own road_event_params make road repairs and detours frequent (dozens per episode
    instead of ~1 on north_pit): against 2048 runs of the unmodified reference per dispatcher
    (tests/golden/make_golden_dist.py ragged_mine 2048).  Welch z-tests at |z| < 3.5 on the means of reported tons
    (CI +- 3.5 * 258 / sqrt(2048) = +- 20 t = 0.6 % for SQ), truck events, orders, jams, road repairs, truck breakdowns, unrepairable trucks, the mean
    production / waiting curves, and KS on the tonnage."""
    import os
    from helpers import GOLDEN
    from test_gpu_parity import live_distribution_check
    ref = np.load(os.path.join(GOLDEN, "live_distribution_ragged_mine.npz"))
    assert len(ref["SQDispatcher_tons"]) >= 2048
    names = ["SQDispatcher", "SPTFDispatcher", "NearestDispatcher", "FixedGroupDispatcher"]
    fields = (("reported_tons", "tons"), ("n_truck_events", "events"), ("n_wrapper_orders", "orders"), ("n_jams", "jams"),
              ("n_road_repairs", "road_repairs"), ("n_truck_repairs", "breakdowns"), ("n_unrepairable", "unrepairable"))
    T = int(ragged_mine["sim_time"])
    live_distribution_check(ragged_mine, ref, names, 4096, 77_000, fields, 3.5, [T // 4, T // 2, 3 * T // 4, T - 1],
                            jam_capacity=128)


def test_limits_mine_32_sites_32_units_per_site_vs_oracle():
    """The engine's stated maxima at once: L + D = 32 sites, 32 shovels at one load site and 32 dumpers at one dump site
    (unit selection and the wait estimates run one lane per unit of a site).  det Naive (every truck to the 32-shovel
    site) / SQ / SPTF bit-exact vs the oracle and a tape replay of a live run."""
    from oracle import des_oracle as do
    from helpers import engine_vs_ref
    from test_host_side import _limits_mine
    import openmines_b200 as om
    cfg = _limits_mine()
    cases = [("NaiveDispatcher", 0), ("SQDispatcher", 0), ("SPTFDispatcher", 0), ("NearestDispatcher", 0)]
    oracles = [do.run_oracle(cfg, d, "det", seed=s) for d, s in cases]
    mine = om.BatchedMine(cfg, len(cases), jam_capacity=256)
    assert mine.pc.L + mine.pc.D == 32
    mine.reset([d for d, _ in cases], seeds=[s for _, s in cases], rng_mode="det")
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
    for i, (d, s) in enumerate(cases):
        diffs = engine_vs_ref(oracles[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
    live = do.run_oracle(cfg, "SQDispatcher", "live", seed=1)
    mine = om.BatchedMine(cfg, 1, jam_capacity=256)
    mine.reset("SQDispatcher", rng_mode="tape", tapes=[live.tape])
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
    diffs = engine_vs_ref(live, mine, res, 0)
    assert not diffs, diffs[:5]


def test_more_than_32_units_per_site_vs_oracle():
    """40 shovels at one load site and 40 dumpers at one dump site: get_available_shovel / get_available_dumper
    (load_site.py:209-220, dump_site.py:105-116) take a second lane trip.  det Naive (every truck to the 40-shovel site, so
    the first idle unit is found in the second trip once 32 are busy) / SQ / SPTF bit-exact vs the oracle + a tape replay."""
    from oracle import des_oracle as do
    from helpers import engine_vs_ref
    from test_host_side import _limits_mine
    import openmines_b200 as om
    cfg = _limits_mine(n_load=6, n_dump=5, units_site0=40)
    cases = [("NaiveDispatcher", 0), ("SQDispatcher", 0), ("SPTFDispatcher", 0)]
    oracles = [do.run_oracle(cfg, d, "det", seed=s) for d, s in cases]
    mine = om.BatchedMine(cfg, len(cases), jam_capacity=256)
    mine.reset([d for d, _ in cases], seeds=[s for _, s in cases], rng_mode="det")
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
    for i, (d, s) in enumerate(cases):
        diffs = engine_vs_ref(oracles[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
    # the second lane trip was taken: some shovel / dumper beyond the first 32 of site 0 served trucks
    sa = mine.state_arrays(0)
    assert sa["u_svc"][32:40].sum() > 0
    live = do.run_oracle(cfg, "NaiveDispatcher", "live", seed=1)
    mine = om.BatchedMine(cfg, 1, jam_capacity=256)
    mine.reset("NaiveDispatcher", rng_mode="tape", tapes=[live.tape])
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
    diffs = engine_vs_ref(live, mine, res, 0)
    assert not diffs, diffs[:5]


def test_older_handle_survives_a_smaller_one(north_pit):
    """The kernels' dynamic shared-memory limit is a per-process function attribute: a second, smaller handle must not
    lower it under a live, larger one (a 4096-env training env followed by a small eval env, run_external(n=1) ...)."""
    import openmines_b200 as om
    big = om.BatchedMine(north_pit, 4096)
    big.reset("SQDispatcher", rng_mode="philox")
    small = om.BatchedMine(north_pit, 1)
    small.reset("SQDispatcher", rng_mode="philox")
    a = small.run(until=30.0).summary_np()
    assert big.launch_info()["smem_bytes"] > small.launch_info()["smem_bytes"]
    b = big.run(until=30.0).summary_np()           # used to fail with cudaErrorInvalidValue
    assert (b["err"] == 0).all() and b["reported_tons"][0] == a["reported_tons"][0]


def test_run_refuses_host_side_dispatchers(north_pit):
    """om_run on instances of a user-defined (host-side) dispatcher would pause them at their first order and return empty
    results: BatchedMine.run raises and points to run_external; run_dispatch_sim routes such dispatchers there."""
    import openmines_b200 as om
    from openmines_b200 import experiments

    class AlwaysFirst(om.BaseDispatcher):
        def __init__(self):
            super().__init__()
            self.name = "AlwaysFirst"

        def give_init_order(self, truck, mine):
            return 0

        give_haul_order = give_back_order = give_init_order

    mine = om.BatchedMine(north_pit, 2)
    mine.reset(AlwaysFirst(), rng_mode="det")
    with pytest.raises(RuntimeError, match="run_external"):
        mine.run()
    short = dict(north_pit, sim_time=30)
    out = experiments.run_dispatch_sim(AlwaysFirst(), short, n_seeds=2, stochastic=False)
    ref = experiments.run_dispatch_sim("NaiveDispatcher", short, n_seeds=2, stochastic=False, want_ticks=False)
    assert np.array_equal(out["produced_tons"], ref["summary"]["reported_tons"])     # both always answer 0


def test_stats_are_exact_and_order_independent(north_pit):
    """om_stats_reduce accumulates integers (fixed-point tonnages): bit-identical to the numpy mirror, identical for a
    shuffled instance order, identical when the batch is cut into 'rank' shards whose partial statistics are added in any
    order -- i.e. all-reduced N-rank totals equal the single-GPU totals over the same global instance ids.  ShortestTrip
    (dispatcher id 6) has its own row."""
    import openmines_b200 as om
    from openmines_b200 import parallel
    n = 1400
    ids = (np.arange(n) % 7).astype(np.int32)
    seeds = np.arange(n, dtype=np.uint64) + 11
    mine = om.BatchedMine(north_pit, n)
    mine.reset(ids, seeds=seeds, rng_mode="philox")
    res = mine.run(until=120.0)
    s = res.summary_np()
    got = mine.stats(res).cpu().numpy()
    again = mine.stats(res).cpu().numpy()
    want = parallel.stats_from_summaries(s, ids)
    assert got.shape == (7, 16) and np.array_equal(got, want) and np.array_equal(got, again)
    assert (got == np.rint(got)).all() and got[6, 0] == 200
    rows = parallel.describe(got)
    for d, name in enumerate(parallel.DISPATCHER_NAMES):
        t = s["reported_tons"][ids == d]
        assert abs(rows[name]["mean_tons"] - t.mean()) <= 1e-6 and abs(rows[name]["std_tons"] - t.std()) <= 0.05
    # shuffled placement
    perm = np.random.default_rng(3).permutation(n)
    m2 = om.BatchedMine(north_pit, n)
    m2.reset(ids[perm], seeds=seeds[perm], rng_mode="philox")
    assert np.array_equal(m2.stats(m2.run(until=120.0)).cpu().numpy(), got)
    # 8 "ranks": partial statistics of contiguous shards, summed in two different orders
    parts = []
    for r in range(8):
        lo, hi = parallel.shard_range(n, 8, r)
        mr = om.BatchedMine(north_pit, hi - lo)
        mr.reset(ids[lo:hi], seeds=seeds[lo:hi], rng_mode="philox")
        parts.append(mr.stats(mr.run(until=120.0)).cpu().numpy())
    fwd = np.zeros_like(got)
    for p in parts:
        fwd = fwd + p
    tree = ((parts[7] + parts[3]) + (parts[5] + parts[1])) + ((parts[6] + parts[2]) + (parts[4] + parts[0]))
    assert np.array_equal(fwd, got) and np.array_equal(tree, got)


def test_split_layout_hot_truck_arrays_and_its_limit(north_pit, monkeypatch):
    """Split layouts keep a truck's unit / queue link / order epoch / service count in a hot 8-byte record, the minute of
    its last breakdown as a uint16 and a pending-repair flag: after the replay of a live reference tape (breakdowns, repairs,
    jams) the decoded state must equal the unsplit layout's field by field, and a clock beyond what a uint16 minute holds is
    refused, not wrapped.  (Replay, not Philox: the two kernels pick the truck a jam position is drawn around from
    differently ordered lists, so their live trajectories are different samples of the same law.)"""
    g = Golden("live_north_pit_mine_SQDispatcher_s11")
    outs = []
    for split in ("0", "1"):
        monkeypatch.setenv("OM_SPLIT", split)
        mine = _mine(north_pit, 2)
        assert bool(mine.layout.t_aux) == (split == "1") and bool(mine.layout.t_lastbrk_hot) == (split == "1")
        assert bool(mine.layout.slot_cache) == (split == "1")
        mine.reset("SQDispatcher", rng_mode="tape", tapes=[g.tape, g.tape])
        s = mine.run(until=200.0).summary_np().copy()          # stop mid-episode: queues, moves and repairs under way
        outs.append((s, [mine.state_arrays(i) for i in range(2)]))
        if split == "1":
            with pytest.raises(RuntimeError, match="65535"):
                mine.run(until=70000.0)
    assert np.array_equal(outs[0][0], outs[1][0]) and (outs[0][0]["err"] == 0).all()
    assert outs[0][0]["n_truck_repairs"].sum() > 0            # breakdowns happened: t_lastbrk / t_repair were exercised
    for a, b in zip(outs[0][1], outs[1][1]):
        assert (a["t_lastbrk"] > 0).any() and (a["t_unit"] != 0xFFFF).any()
        for f in ("t_unit", "t_next", "t_epoch", "t_svc", "t_lastbrk", "t_repair", "t_status", "t_cur", "t_tgt", "t_total",
                  "t_cycle", "t_dist", "t_load", "t_time", "t_key", "u_tons", "u_qlen", "u_head", "u_tail"):
            assert np.array_equal(a[f], b[f]), f
