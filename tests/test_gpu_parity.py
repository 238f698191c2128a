"""GPU parity tests: the CUDA engine (through the C ABI) against the oracle and the golden
vectors recorded from the unmodified reference.  Integer / index / event-sequence work is compared
bit-exact; fp64 times and tonnages are compared bit-exact too (tolerance 0), which is stricter
than the 1e-9 relative bound of BASELINE.json's north_star.
"""
import numpy as np
import pytest

from helpers import DISPATCHERS, Golden, engine_vs_ref

pytestmark = pytest.mark.gpu


def _mine(cfg, n, **kw):
    import openmines_b200 as om
    return om.BatchedMine(cfg, n, **kw)


@pytest.mark.parametrize("cfg_name", ["north_pit_mine", "north_pit_mine_short"])
def test_det_all_dispatchers_vs_golden(cfg_name, north_pit, short_mine):
    """det mode: event sequence, tick arrays, final tonnage bit-exact vs the reference goldens."""
    cfg = north_pit if cfg_name == "north_pit_mine" else short_mine
    cases = [(d, 0) for d in DISPATCHERS] + [("RandomDispatcher", 7)]
    mine = _mine(cfg, len(cases))
    mine.reset([d for d, _ in cases], seeds=[s for _, s in cases], rng_mode="det")
    res = mine.run(trace_capacity=4096, site_ticks=True, order_capacity=2048)
    for i, (d, seed) in enumerate(cases):
        g = Golden("det_%s_%s_s%d" % (cfg_name, d, seed))
        diffs = engine_vs_ref(g, mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s seed %d: %s" % (d, seed, diffs[:5])


def test_tape_replay_vs_golden(north_pit):
    """tape mode: the engine consumes the reference's recorded draws -> trajectories bit-exact."""
    goldens = [Golden("live_north_pit_mine_%s_s11" % d) for d in DISPATCHERS]
    mine = _mine(north_pit, len(DISPATCHERS))
    mine.reset(DISPATCHERS, rng_mode="tape", tapes=[g.tape for g in goldens])
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=2048)
    for i, d in enumerate(DISPATCHERS):
        diffs = engine_vs_ref(goldens[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
        s = res.summary_np()[i]
        assert s["err"] == 0


def test_split_state_mode_vs_golden(north_pit, monkeypatch):
    """The split layout (truck records + jam pool accessed in place in HBM, only the hot part staged in shared
    memory; chosen automatically for big mines) must give the same bits: forced here on north_pit, det + tape."""
    monkeypatch.setenv("OM_SPLIT", "1")
    cases = [(d, 0) for d in DISPATCHERS]
    mine = _mine(north_pit, len(cases))
    assert mine.layout.hot_bytes < mine.layout.total_bytes
    assert mine.launch_info()["smem_bytes"] >= mine.launch_info()["warps_per_cta"] * mine.layout.hot_bytes
    mine.reset([d for d, _ in cases], seeds=[s for _, s in cases], rng_mode="det")
    res = mine.run(trace_capacity=4096, site_ticks=True, order_capacity=2048)
    for i, (d, seed) in enumerate(cases):
        g = Golden("det_north_pit_mine_%s_s%d" % (d, seed))
        diffs = engine_vs_ref(g, mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s seed %d: %s" % (d, seed, diffs[:5])
    goldens = [Golden("live_north_pit_mine_%s_s11" % d) for d in DISPATCHERS]
    mine = _mine(north_pit, len(DISPATCHERS))
    mine.reset(DISPATCHERS, rng_mode="tape", tapes=[g.tape for g in goldens])
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=2048)
    for i, d in enumerate(DISPATCHERS):
        diffs = engine_vs_ref(goldens[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])


def test_tape_replay_short(short_mine):
    names = ["RandomDispatcher", "SQDispatcher"]
    goldens = [Golden("live_north_pit_mine_short_%s_s3" % d) for d in names]
    mine = _mine(short_mine, 2)
    mine.reset(names, rng_mode="tape", tapes=[g.tape for g in goldens])
    res = mine.run(trace_capacity=1024, site_ticks=True, order_capacity=256)
    for i, d in enumerate(names):
        diffs = engine_vs_ref(goldens[i], mine, res, i)
        assert not diffs, "%s: %s" % (d, diffs[:5])


def test_det_vs_oracle_seeds(north_pit):
    """det mode on seeds the goldens do not cover: RandomDispatcher (keyed Philox) vs the CPU oracle."""
    from oracle import des_oracle as do
    seeds = [1, 2, 3, 1000003, 2 ** 40 + 5]
    mine = _mine(north_pit, len(seeds))
    mine.reset("RandomDispatcher", seeds=seeds, rng_mode="det")
    res = mine.run(trace_capacity=4096, site_ticks=True, order_capacity=2048)
    for i, seed in enumerate(seeds):
        o = do.run_oracle(north_pit, "RandomDispatcher", "det", seed=seed)
        diffs = engine_vs_ref(o, mine, res, i)
        assert not diffs, "seed %d: %s" % (seed, diffs[:5])


def test_tape_replay_vs_oracle_fresh_seeds(north_pit):
    """live oracle runs on fresh seeds -> tape -> engine replay, all six dispatchers."""
    from oracle import des_oracle as do
    cases = [(d, 100 + k) for k, d in enumerate(DISPATCHERS)]
    oracles = [do.run_oracle(north_pit, d, "live", seed=s) for d, s in cases]
    mine = _mine(north_pit, len(cases))
    mine.reset([d for d, _ in cases], rng_mode="tape", tapes=[o.tape for o in oracles])
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=2048)
    for i, (d, s) in enumerate(cases):
        diffs = engine_vs_ref(oracles[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s seed %d: %s" % (d, s, diffs[:5])


def test_c2_shape_identical_trajectories(north_pit):
    """BASELINE config 2 property: for the five seed-independent dispatchers every seed gives the
    identical det trajectory (same events, tons); checked on 6 x 64 instances."""
    n_seeds = 64
    names = [d for d in DISPATCHERS for _ in range(n_seeds)]
    seeds = [s for _ in DISPATCHERS for s in range(n_seeds)]
    mine = _mine(north_pit, len(names))
    mine.reset(names, seeds=seeds, rng_mode="det")
    res = mine.run()
    s = res.summary_np().reshape(len(DISPATCHERS), n_seeds)
    ticks = res.ticks_np().reshape(len(DISPATCHERS), n_seeds, -1)
    for k, d in enumerate(DISPATCHERS):
        assert (s[k]["err"] == 0).all()
        if d == "RandomDispatcher":
            assert len(np.unique(s[k]["reported_tons"])) > 8
            continue
        g = Golden("det_north_pit_mine_%s_s0" % d)
        assert (s[k]["reported_tons"] == g.reported_tons).all(), d
        assert (s[k]["n_truck_events"] == g.n_truck_events).all(), d
        assert (ticks[k]["produced_tons"][:, :g.n_ticks] == g.tick["produced_tons"][None, :]).all(), d


def test_resume_equals_single_run(north_pit):
    """om_run(until=a) then om_run(until=b) == om_run(until=b): state round-trips HBM exactly."""
    mine = _mine(north_pit, 6)
    mine.reset(DISPATCHERS, rng_mode="philox", seeds=np.arange(6) + 50)
    r1 = mine.run(trace_capacity=8192)
    a = r1.summary_np().copy()
    t1 = [r1.trace_np(i)[0].copy() for i in range(6)]
    mine.reset(DISPATCHERS, rng_mode="philox", seeds=np.arange(6) + 50)
    r2 = mine.run(until=77.5, trace_capacity=8192)
    r2 = mine.run(until=150.0, result=r2)
    r2 = mine.run(result=r2)
    b = r2.summary_np()
    for f in ("reported_tons", "dumped_tons", "n_truck_events", "n_decisions", "n_jams", "n_truck_repairs", "err"):
        assert np.array_equal(a[f], b[f]), f
    for i in range(6):
        assert np.array_equal(t1[i], r2.trace_np(i)[0])


def _z(a_mean, a_var, n_a, b_mean, b_var, n_b):
    return (a_mean - b_mean) / np.sqrt(a_var / n_a + b_var / n_b + 1e-300)


def live_distribution_check(cfg, ref, names, n_seeds, seed0, fields, zmax, curve_minutes, **mine_kw):
    """Engine (Philox, ``n_seeds`` per dispatcher) against per-seed statistics of the unmodified reference (``ref``: npz of
    tests/golden/make_golden_dist.py).  Welch z on the means of every field, two-sample KS on tons, and Welch z on the mean
    production curve and mean waiting-truck curve at ``curve_minutes``.  Returns the report rows."""
    from scipy import stats
    mine = _mine(cfg, len(names) * n_seeds, **mine_kw)
    mine.reset([d for d in names for _ in range(n_seeds)],
               seeds=np.tile(np.arange(n_seeds, dtype=np.uint64) + seed0, len(names)), rng_mode="philox")
    res = mine.run()
    s = res.summary_np().reshape(len(names), n_seeds)
    ticks = res.ticks_np().reshape(len(names), n_seeds, -1)
    assert (s["err"] == 0).all()
    report, worst = [], 0.0
    for k, d in enumerate(names):
        n_ref = len(ref[d + "_tons"])
        for field, key in fields:
            a = s[k][field].astype(np.float64)
            b = ref["%s_%s" % (d, key)].astype(np.float64)
            z = _z(a.mean(), a.var(ddof=1), len(a), b.mean(), b.var(ddof=1), len(b))
            report.append((d, key, a.mean(), b.mean(), z))
            worst = max(worst, abs(z))
        ks = stats.ks_2samp(s[k]["reported_tons"], ref[d + "_tons"])
        report.append((d, "KS p(tons)", ks.pvalue, 1.0, 0.0))
        assert ks.pvalue > 1e-3, "%s tons KS p=%.2e" % (d, ks.pvalue)
        for what, col in (("curve", "produced_tons"), ("wait", "waiting")):
            rs, rq = ref["%s_%s_sum" % (d, what)], ref["%s_%s_sumsq" % (d, what)]
            rmean = rs / n_ref
            rvar = np.maximum(rq / n_ref - rmean ** 2, 0.0) * n_ref / (n_ref - 1)
            e = ticks[k][col][:, :len(rs)].astype(np.float64)
            for m in curve_minutes:
                if rvar[m] == 0.0 and e[:, m].var() == 0.0:
                    assert e[:, m].mean() == rmean[m], (d, what, m)
                    continue
                z = _z(e[:, m].mean(), e[:, m].var(ddof=1), n_seeds, rmean[m], rvar[m], n_ref)
                report.append((d, "%s@%d" % (what, m), e[:, m].mean(), rmean[m], z))
                worst = max(worst, abs(z))
    for r in report:
        print("%-22s %-14s engine %10.3f  reference %10.3f  z=%+.2f" % r)
    bad = [r for r in report if abs(r[4]) >= zmax]
    assert not bad, "beyond |z| < %.1f: %s" % (zmax, bad)
    return report


def test_live_distribution_vs_reference(north_pit):
    """live RNG (Philox) against 1024 stochastic runs of the UNMODIFIED reference per dispatcher
    (tests/golden/live_distribution_north_pit_mine.npz, generated by tests/golden/make_golden_dist.py).

    Stated confidence interval: the reference sample fixes the mean reported tonnage of a dispatcher to
    +- 3.5 * sigma / sqrt(1024) = +- 34 t for SQ (sigma 311 t; +- 0.39 % of 8842 t), +- 35 t for SPTF; the engine runs 8192
    seeds per dispatcher, so its own sampling error is 2.8x smaller.  Every check is a Welch z-test on the difference of the
    means at |z| < 3.5 (two-sided p = 4.7e-4 per check): reported tons, truck events, wrapper orders, jams, road repairs,
    truck breakdowns, unrepairable trucks per episode, and the MEAN PRODUCTION CURVE and mean waiting-truck count at minutes
    30, 60, ..., 210, 239; plus a two-sample KS test on the tonnage distribution (p > 1e-3)."""
    import os
    from helpers import GOLDEN
    ref = np.load(os.path.join(GOLDEN, "live_distribution_north_pit_mine.npz"))
    assert len(ref["SQDispatcher_tons"]) >= 1024
    fields = (("reported_tons", "tons"), ("n_truck_events", "events"), ("n_wrapper_orders", "orders"), ("n_jams", "jams"),
              ("n_road_repairs", "road_repairs"), ("n_truck_repairs", "breakdowns"), ("n_unrepairable", "unrepairable"))
    live_distribution_check(north_pit, ref, DISPATCHERS, 8192, 10_000, fields, 3.5, [30, 60, 90, 120, 150, 180, 210, 239])


def test_stats_kernel_matches_host_mirror(north_pit):
    """om_stats_reduce (the partial sums fed to the NCCL all-reduce) == the numpy mirror."""
    from openmines_b200 import parallel
    n = 600
    ids = np.arange(n) % 6
    mine = _mine(north_pit, n)
    mine.reset(ids.astype(np.int32), seeds=np.arange(n) + 7, rng_mode="philox")
    res = mine.run(until=90.0)
    got = mine.stats(res).cpu().numpy()
    want = parallel.stats_from_summaries(res.summary_np(), ids)
    assert np.array_equal(got, want)            # integer accumulation: exact
    rows = parallel.describe(got)
    assert set(rows) == set(DISPATCHERS) and all(r["instances"] == 100 for r in rows.values())


def test_scaled_mine_c5_vs_oracle():
    """BASELINE configs[4] shape (500 trucks, 20 load sites, 10 dump sites, 420 roads): det SQ / SPTF /
    Random bit-exact vs the oracle, and a tape replay of a live SQ oracle run."""
    from oracle import des_oracle as do
    from openmines_b200.synthetic import scaled_mine
    cfg = scaled_mine()
    cases = [("SQDispatcher", "det", 0), ("SPTFDispatcher", "det", 0), ("RandomDispatcher", "det", 5)]
    oracles = [do.run_oracle(cfg, d, m, seed=s) for d, m, s in cases]
    mine = _mine(cfg, len(cases), jam_capacity=512)
    mine.reset([d for d, _, _ in cases], seeds=[s for _, _, s in cases], rng_mode="det")
    res = mine.run(trace_capacity=16384, site_ticks=True, order_capacity=8192)
    for i, (d, m, s) in enumerate(cases):
        diffs = engine_vs_ref(oracles[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
    live = do.run_oracle(cfg, "SQDispatcher", "live", seed=1)
    mine = _mine(cfg, 1, jam_capacity=512)
    mine.reset("SQDispatcher", rng_mode="tape", tapes=[live.tape])
    res = mine.run(trace_capacity=32768, site_ticks=True, order_capacity=8192)
    diffs = engine_vs_ref(live, mine, res, 0)
    assert not diffs, diffs[:5]


def test_mine_with_more_than_1024_trucks_vs_oracle():
    """1100 trucks: more than 32 slots of 32 pending entries, so the cached slot minima of the SPLIT kernels take their
    multi-trip paths (and 160 units / 420 roads span several lane trips everywhere).  det SQ / SPTF bit-exact vs the
    oracle, plus a tape replay of a live oracle run."""
    from oracle import des_oracle as do
    from openmines_b200.synthetic import scaled_mine
    cfg = scaled_mine(n_trucks=(400, 400, 300), sim_time=40)
    cases = [("SQDispatcher", "det", 0), ("SPTFDispatcher", "det", 0)]
    oracles = [do.run_oracle(cfg, d, m, seed=s) for d, m, s in cases]
    mine = _mine(cfg, len(cases), jam_capacity=1024)
    assert mine.pc.T == 1100 and mine.layout.hot_bytes < mine.layout.total_bytes        # split mode
    mine.reset([d for d, _, _ in cases], seeds=[s for _, _, s in cases], rng_mode="det")
    res = mine.run(trace_capacity=16384, site_ticks=True, order_capacity=8192)
    for i, (d, m, s) in enumerate(cases):
        diffs = engine_vs_ref(oracles[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
    live = do.run_oracle(cfg, "SQDispatcher", "live", seed=1)
    mine = _mine(cfg, 1, jam_capacity=1024)
    mine.reset("SQDispatcher", rng_mode="tape", tapes=[live.tape])
    res = mine.run(trace_capacity=16384, site_ticks=True, order_capacity=8192)
    diffs = engine_vs_ref(live, mine, res, 0)
    assert not diffs, diffs[:5]


def test_run_before_reset_is_an_error_not_a_hang(north_pit):
    """om_run / om_step before om_reset must fail on the host (a never-written blob could make the kernel spin)."""
    import torch
    from openmines_b200.rl_env import BatchedMineEnv
    mine = _mine(north_pit, 8)
    with pytest.raises(RuntimeError, match="om_reset has not been called"):
        mine.run()
    env = BatchedMineEnv(north_pit, 4)
    with pytest.raises(RuntimeError, match="om_reset has not been called"):
        env.step(torch.zeros(4, dtype=torch.int32, device=env.mine.device))


def test_lean_and_rl_layouts_agree(north_pit):
    """Handles created with and without the RL surface (OM_CFG_RL_SURFACE: bigger header, reward-window fields, both
    sides' wait estimates live at once) must simulate identically; om_step is refused on the lean layout."""
    import torch
    lean = _mine(north_pit, len(DISPATCHERS))
    full = _mine(north_pit, len(DISPATCHERS), rl_surface=True)
    assert lean.layout.total_bytes < full.layout.total_bytes and lean.layout.hdr_bytes == 128 and full.layout.hdr_bytes == 192
    assert lean.launch_info()["warps_per_cta"] >= full.launch_info()["warps_per_cta"]
    outs = []
    for m in (lean, full):
        m.reset(DISPATCHERS, seeds=np.arange(len(DISPATCHERS)) + 3, rng_mode="philox")
        res = m.run(trace_capacity=8192, site_ticks=True)
        outs.append((res.summary_np().copy(), res.ticks_np().copy(), [res.trace_np(i)[0].copy() for i in range(len(DISPATCHERS))]))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2], outs[1][2]):
        assert np.array_equal(a, b)
    io = {"obs": torch.zeros((len(DISPATCHERS), lean.obs_dim()), device=lean.device)}
    with pytest.raises(RuntimeError, match="OM_CFG_RL_SURFACE"):
        lean.step(None, io)


def test_episode_pipeline_matches_direct_runs(north_pit):
    """EpisodePipeline (double-buffered outputs, copies on a side stream) returns, in submission order, exactly what
    synchronous reset + run calls return."""
    from openmines_b200.pipeline import EpisodePipeline
    n = 96
    ids = (np.arange(n) % 6).astype(np.int32)
    mine = _mine(north_pit, n)
    want = []
    for b in range(4):
        mine.reset(ids, seeds=np.arange(n) + 100 * b, rng_mode="philox")
        r = mine.run()
        want.append((r.summary_np().copy(), r.ticks_np().copy()))
    pipe = EpisodePipeline(mine, rng_mode="philox")
    got = []
    for b in range(4):
        pipe.submit(ids, np.arange(n, dtype=np.uint64) + 100 * b)
        if pipe.pending == pipe.depth:
            s, t = pipe.collect()
            got.append((s.copy(), t.copy()))
    while pipe.pending:
        s, t = pipe.collect()
        got.append((s.copy(), t.copy()))
    assert len(got) == 4
    for (ws, wt), (gs, gt) in zip(want, got):
        assert np.array_equal(ws, gs) and np.array_equal(wt, gt)
    with pytest.raises(RuntimeError):
        for _ in range(3):
            pipe.submit(ids, np.arange(n, dtype=np.uint64))


def _shard_properties(cfg, names, seeds, n_sub, n_decode, **mine_kw):
    """Size-independent properties of one per-GPU shard of live (Philox) instances -- see the two tests below."""
    n = len(names)
    T = int(cfg["sim_time"])
    mine = _mine(cfg, n, **mine_kw)
    mine.reset(names, seeds=seeds, rng_mode="philox")
    res = mine.run()
    s, ticks = res.summary_np(), res.ticks_np()
    assert (s["err"] == 0).all()
    assert (s["n_ticks"] == T).all()

    # placement independence
    rng = np.random.default_rng(7)
    perm = rng.permutation(n)
    mine_b = _mine(cfg, n, **mine_kw)
    mine_b.reset([names[j] for j in perm], seeds=seeds[perm], rng_mode="philox")
    res_b = mine_b.run()
    assert res_b.summary_np().tobytes() == s[perm].tobytes()
    assert res_b.ticks_np().tobytes() == ticks[perm].tobytes()
    sub = perm[:n_sub]
    mine_c = _mine(cfg, len(sub), **mine_kw)
    mine_c.reset([names[j] for j in sub], seeds=seeds[sub], rng_mode="philox")
    res_c = mine_c.run()
    assert res_c.summary_np().tobytes() == s[sub].tobytes()
    assert res_c.ticks_np().tobytes() == ticks[sub].tobytes()

    # production curves
    prod = ticks["produced_tons"][:, :T]
    assert (np.diff(prod, axis=1) >= 0).all()
    assert (prod[:, -1] == s["reported_tons"]).all()
    assert (s["reported_tons"] <= s["dumped_tons"] * (1 + 1e-12)).all()     # the two sums run in different orders
    n_trucks = mine.pc.T
    for f in ("working", "waiting", "load_unload", "moving", "repairing"):
        assert (ticks[f][:, :T] <= n_trucks).all(), f
    assert (s["n_truck_events"] >= s["n_decisions"]).all()
    assert (s["n_decisions"] >= n_trucks).all()            # every truck got at least its init order

    # conservation, on a sample of instances (the state blobs are decoded on the host)
    S = mine.pc.S
    for j in rng.choice(n, size=n_decode, replace=False):
        st = mine.state_arrays(int(j))
        dumped = float(s["dumped_tons"][j])
        assert abs(st["u_tons"][S:].sum() - dumped) <= 1e-9 * max(1.0, dumped)
        assert abs(st["t_total"].sum() - dumped) <= 1e-9 * max(1.0, dumped)
        loaded = st["u_tons"][:S].sum()
        carried = st["t_load"].sum()                       # loaded but not yet dumped
        assert abs(loaded - carried - dumped) <= 1e-9 * max(1.0, loaded)
        assert int(st["u_svc"][S:].sum()) == int(st["t_svc"].sum())
    return s


def test_c3_full_shard_properties(north_pit):
    """BASELINE configs[2] at its full per-GPU size (8192 seeds x {SQ, SPTF}, stochastic events on, 16384 instances):
    no oracle finishes that in seconds, so the check is through size-independent properties --
    * placement independence: the same (dispatcher, seed) pairs in a shuffled instance order and in a smaller batch
      give bit-identical summaries and production curves (Philox streams are keyed by seed, not by warp or slot);
    * conservation: tons dumped == sum over dumpers == sum over trucks (a load counts once at the dumper and once at the
      truck), reported tons == the last sampled point of the production curve <= dumped tons;
    * monotone production curves, one sample per simulated minute, truck state counts within the fleet size;
    * no instance ends with an engine error."""
    n_seeds = 8192
    names = ["SQDispatcher", "SPTFDispatcher"] * n_seeds
    seeds = np.repeat(np.arange(n_seeds, dtype=np.int64), 2)
    s = _shard_properties(north_pit, names, seeds, n_sub=1000, n_decode=48)
    # the two dispatchers see the same mines but decide differently
    assert s[0::2]["reported_tons"].mean() != s[1::2]["reported_tons"].mean()


def test_c5_full_shard_properties():
    """BASELINE configs[4] at its per-GPU size (500 trucks, 20 load sites, 10 dump sites; 1024 live instances, the split
    state mode with its shared-memory helper arrays and cached slot minima): the same properties."""
    from openmines_b200.synthetic import scaled_mine
    n = 1024
    names = ["SQDispatcher", "SPTFDispatcher"] * (n // 2)
    seeds = np.arange(n, dtype=np.int64)
    _shard_properties(scaled_mine(), names, seeds, n_sub=100, n_decode=8, jam_capacity=512)


def test_c_program_through_the_c_abi_vs_golden(tmp_path):
    """examples/c_api_demo.c: a plain C program (gcc, CUDA runtime only -- no Python, no PyTorch in the process) packs the
    one-truck mine by hand, runs it through om_create / om_reset / om_run and must print the unmodified reference's
    tonnage, event count and production curve for that mine."""
    import re
    import subprocess
    from test_host_side import _build_c_demo
    exe = _build_c_demo(tmp_path)
    r = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    g = Golden("det_one_truck_mine_NaiveDispatcher_s0")
    rows = re.findall(r"instance (\d+): reported_tons (\S+) dumped_tons (\S+) truck_events (\d+) decisions (\d+) ticks (\d+) "
                      r"err (\d+) tons@60min (\S+)", r.stdout)
    assert len(rows) == 3, r.stdout
    for row in rows:
        assert float(row[1]) == float(g.reported_tons)
        assert int(row[3]) == int(g.n_truck_events)
        assert int(row[5]) == int(g.n_ticks) and int(row[6]) == 0
        assert float(row[7]) == float(g.tick["produced_tons"][60])


def test_shortest_trip_dispatcher_vs_reference_goldens(north_pit, ragged_mine):
    """OM_DISP_SHORTESTTRIP in om_run against goldens of the unmodified reference running ShortestTripDispatcher as a
    dispatcher (tests/golden/make_golden_teacher.py): det and tape replay, north_pit and the ragged mine."""
    d = "ShortestTripDispatcher"
    for cname, cfg, jam in (("north_pit_mine", north_pit, 64), ("ragged_mine", ragged_mine, 128)):
        gd, gl = Golden("det_%s_%s_s0" % (cname, d)), Golden("live_%s_%s_s17" % (cname, d))
        mine = _mine(cfg, 1, jam_capacity=jam)
        mine.reset(d, rng_mode="det")
        res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
        assert not engine_vs_ref(gd, mine, res, 0), cname
        mine.reset(d, rng_mode="tape", tapes=[gl.tape])
        res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
        assert not engine_vs_ref(gl, mine, res, 0), cname


def test_naive_sum_flavour_vs_oracle(north_pit):
    """OmConfig.sum_mode = 0: the reference under CPython <= 3.11 (its CI pins 3.9.6, .github/workflows/pypi.yml:19), whose
    sum() adds left to right instead of the Neumaier-compensated sum of CPython >= 3.12.  Decisions of SPTF
    (sptf_dispatcher.py:35) and every reported tonnage (dump_site.py:69, mine.py:61) go through sum(): live oracle runs with
    sum_mode="naive" -> tape -> engine with sum_mode="naive", bit-exact; and the two flavours do differ on the same tape."""
    from oracle import des_oracle as do
    names = ["SPTFDispatcher", "SQDispatcher", "FixedGroupDispatcher", "RandomDispatcher"]
    oracles = [do.run_oracle(north_pit, d, "live", seed=40 + k, sum_mode="naive") for k, d in enumerate(names)]
    mine = _mine(north_pit, len(names), sum_mode="naive")
    mine.reset(names, rng_mode="tape", tapes=[o.tape for o in oracles])
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=2048)
    for i, d in enumerate(names):
        diffs = engine_vs_ref(oracles[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (d, diffs[:5])
    naive_curves = res.ticks_np()["produced_tons"].copy()
    comp = _mine(north_pit, len(names))                       # default: Neumaier
    comp.reset(names, rng_mode="tape", tapes=[o.tape for o in oracles])
    res2 = comp.run(trace_capacity=8192)
    comp_curves = res2.ticks_np()["produced_tons"]
    assert (naive_curves != comp_curves).any(), "the summation flavour should show in the last bits of the production curves"
    assert np.allclose(naive_curves, comp_curves, rtol=1e-12)


def test_jam_pool_overflow_is_regrown_not_lost(north_pit):
    """A jam pool that is too small stops the instance with OM_IERR_JAM_OVERFLOW; run(regrow=True) re-simulates exactly those
    instances with a larger pool and patches their outputs in, so the batch equals one run with a big pool from the start."""
    n = 256
    names = ["SQDispatcher", "SPTFDispatcher"] * (n // 2)
    seeds = np.arange(n, dtype=np.uint64) + 900
    big = _mine(north_pit, n, jam_capacity=128)
    big.reset(names, seeds=seeds, rng_mode="philox")
    want = big.run()
    small = _mine(north_pit, n, jam_capacity=12)
    small.reset(names, seeds=seeds, rng_mode="philox")
    raw = small.run().summary_np()
    n_over = int((raw["err"] == 3).sum())
    assert n_over > 0                                        # 12 entries do overflow on north_pit
    small.reset(names, seeds=seeds, rng_mode="philox")
    got = small.run(regrow=True)
    assert got.regrown == n_over
    assert got.summary_np().tobytes() == want.summary_np().tobytes()
    assert got.ticks_np().tobytes() == want.ticks_np().tobytes()



def test_episode_batch_in_a_cuda_graph(north_pit):
    """om_reset + om_run + om_stats_reduce captured in a CUDA graph and replayed: the work-queue counter and the per-instance
    segment flags are left as the launch found them (no memset, no epoch in the kernel parameters), so every replay gives
    the eager result -- on a batch big enough to be cut into episode segments."""
    import torch
    n = 4096                                         # 1.4 waves of 2960 resident warps -> segmented episodes
    ids = np.where(np.arange(n) % 2 == 0, 5, 4).astype(np.int32)
    mine = _mine(north_pit, n)
    mine.reset(ids, seeds=np.arange(n) // 2, rng_mode="philox")
    res = mine.run(until=60.0)
    want_s, want_t = res.summary_np().copy(), res.ticks_np().copy()
    want_stats = mine.stats(res).cpu().numpy()
    import ctypes
    from openmines_b200.engine import RNG_MODES
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    stats_box = []

    def body():
        mine._ck(mine.lib.om_reset(mine.h, mine.seeds.data_ptr(), mine.dispatcher_ids.data_ptr(), RNG_MODES["philox"],
                                   None, None, None, mine.state.data_ptr(), mine._stream()))
        mine.run(until=60.0, result=res)
        stats_box[:] = [mine.stats(res)]
    with torch.cuda.stream(side):
        body()                                       # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    for _ in range(3):
        res.summary.zero_(); res.ticks.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert res.summary_np().tobytes() == want_s.tobytes() and res.ticks_np().tobytes() == want_t.tobytes()
        assert np.array_equal(stats_box[0].cpu().numpy(), want_stats)
