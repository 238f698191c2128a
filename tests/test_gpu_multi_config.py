"""Mixed-config batches (the "config" axis of "seed x dispatcher x config"): om_create_multi holds several packed mines
and every instance names the one it simulates; one launch runs them all.

* parity: instances of five different mines (north_pit, the 3-truck short mine, the ragged mine with L != D, the one-truck
  mine, the integer-duration mine) x dispatchers in ONE batch reproduce, bit for bit, the goldens recorded from the
  unmodified reference for each mine on its own (each instance stops at its own config's sim_time);
* the reference's own fleet-ablation drivers (cli/run.py:283-472, goldens in tests/golden/ablation_det.json from
  tests/golden/make_golden_ablation.py) as one launch each;
* the Philox kernels on a mixed batch: draw tape -> CPU replay, per instance with its own config.
"""
import json
import os

import numpy as np
import pytest

from conftest import load_cfg
from helpers import DISPATCHERS, GOLDEN, Golden, HostResult, engine_vs_ref, oracle_replays

pytestmark = pytest.mark.gpu


def test_mixed_config_batch_vs_reference_goldens():
    import openmines_b200 as om
    mines = ["north_pit_mine", "north_pit_mine_short", "ragged_mine", "one_truck_mine", "integer_mine"]
    cfgs = [load_cfg(m) for m in mines]
    cases = []                                   # (config index, dispatcher, golden name)
    for k, m in enumerate(mines):
        for d in (["NaiveDispatcher"] if m == "one_truck_mine" else DISPATCHERS):
            cases.append((k, d, "det_%s_%s_s0" % (m, d)))
    # interleave the mines so neighbouring warps run different configs
    order = np.random.default_rng(5).permutation(len(cases))
    cases = [cases[i] for i in order]
    mine = om.BatchedMine(cfgs, len(cases), config_ids=[c[0] for c in cases], jam_capacity=[64, 64, 128, 64, 64])
    assert mine.multi and mine.max_T == 130 and mine.layout.total_bytes >= om.BatchedMine(cfgs[2], 1).layout.total_bytes
    mine.reset([c[1] for c in cases], seeds=np.zeros(len(cases), dtype=np.uint64), rng_mode="det")
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
    s = res.summary_np()
    for i, (k, d, gname) in enumerate(cases):
        g = Golden(gname)
        diffs = engine_vs_ref(g, mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (gname, diffs[:5])
        assert int(s["n_ticks"][i]) == g.n_ticks
    # the tape goldens of three different mines, replayed in one batch
    tcases = [(0, "SQDispatcher", "live_north_pit_mine_SQDispatcher_s11"), (2, "SPTFDispatcher", "live_ragged_mine_SPTFDispatcher_s5"),
              (1, "RandomDispatcher", "live_north_pit_mine_short_RandomDispatcher_s3"), (0, "SPTFDispatcher", "live_north_pit_mine_SPTFDispatcher_s11"),
              (3, "NaiveDispatcher", "live_one_truck_mine_NaiveDispatcher_s2"), (2, "RandomDispatcher", "live_ragged_mine_RandomDispatcher_s5")]
    goldens = [Golden(t[2]) for t in tcases]
    mine = om.BatchedMine(cfgs, len(tcases), config_ids=[t[0] for t in tcases], jam_capacity=[64, 64, 128, 64, 64])
    mine.reset([t[1] for t in tcases], rng_mode="tape", tapes=[g.tape for g in goldens])
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096)
    for i, (k, d, gname) in enumerate(tcases):
        diffs = engine_vs_ref(goldens[i], mine, res, i, sptf=(d == "SPTFDispatcher"))
        assert not diffs, "%s: %s" % (gname, diffs[:5])


def test_fleet_ablation_drivers_vs_reference_golden():
    """scene_based_fleet_ablation / algo_based_fleet_ablation (ONE mixed-config launch each) reproduce exactly what the
    unmodified reference's run_scene_based_fleet_ablation_experiment / run_algo_based_fleet_ablation_experiment hand to
    Charter in deterministic mode; the serial driver (one launch per fleet size) gives the same numbers."""
    from openmines_b200 import experiments as ex
    with open(os.path.join(GOLDEN, "ablation_det.json"), encoding="utf-8") as fh:
        gold = json.load(fh)
    for cname, g in gold["scene"].items():
        cfg = dict(load_cfg(cname), sim_time=g["sim_time"])
        names = list(g["results"].keys())
        got = ex.scene_based_fleet_ablation(cfg, g["min_truck"], g["max_truck"], dispatchers=names, n_seeds=1, stochastic=False)
        for d in names:
            assert got[d]["fleet_sizes"] == g["results"][d]["fleet_sizes"], (cname, d)
            assert got[d]["productions"] == g["results"][d]["productions"], (cname, d, got[d]["productions"], g["results"][d]["productions"])
        if cname == "north_pit_mine_short":
            ser = ex.scene_based_fleet_ablation_serial(cfg, g["min_truck"], g["max_truck"], dispatchers=names, n_seeds=1, stochastic=False)
            assert all(ser[d]["productions"] == got[d]["productions"] for d in names)
    a = gold["algo"]
    scenes = [dict(load_cfg(n), sim_time=a["sim_times"][n]) for n in a["scenes"]]
    for sc, n in zip(scenes, a["scenes"]):
        sc["mine"] = dict(sc["mine"], name=n)
    got = ex.algo_based_fleet_ablation(scenes, a["baseline"], a["target"], a["min_truck"], a["max_truck"], n_seeds=1, stochastic=False)
    for n, g in a["scenes"].items():
        assert got[n]["fleet_sizes"] == g["fleet_sizes"], n
        assert got[n]["ratios"] == g["ratios"], (n, got[n]["ratios"], g["ratios"])


def test_mixed_config_philox_draw_tape_replay(north_pit, ragged_mine, short_mine):
    """Live (Philox) instances of three mines x {SQ, SPTF, Random} in one batch: every instance's exported draw tape,
    replayed by the CPU restatement on that instance's own config, reproduces its trajectory bit for bit; and the mixed batch
    gives each instance exactly what a one-config batch gives it."""
    import openmines_b200 as om
    cfgs = [north_pit, ragged_mine, short_mine]
    names, cids, seeds = [], [], []
    for r in range(16):
        for k in range(3):
            for d in ("SQDispatcher", "SPTFDispatcher", "RandomDispatcher"):
                if k == 1 and d == "RandomDispatcher":
                    continue                   # L = 3 > D = 2: the reference's RandomDispatcher dies with IndexError there
                names.append(d); cids.append(k); seeds.append(1000 + r)
    n = len(names)
    mine = om.BatchedMine(cfgs, n, config_ids=cids, jam_capacity=[64, 128, 64])
    mine.reset(names, seeds=seeds, rng_mode="philox")
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096, draw_capacity=mine.draw_tape_slots())
    host = HostResult(mine, res)
    s = host.summary_np()
    assert (s["err"] == 0).all()
    tapes = [host.draw_tape(i) for i in range(n)]
    oracles = oracle_replays([(cfgs[cids[i]], names[i], tapes[i], {}) for i in range(n)])
    for i in range(n):
        diffs = engine_vs_ref(oracles[i], mine, host, i, sptf=(names[i] == "SPTFDispatcher"))
        assert not diffs, "instance %d (config %d, %s): %s" % (i, cids[i], names[i], diffs[:5])
        assert np.array_equal(oracles[i].tape[1], tapes[i][1])
    # one-config batches of the same (dispatcher, seed) pairs
    for k in range(3):
        idx = [i for i in range(n) if cids[i] == k]
        one = om.BatchedMine(cfgs[k], len(idx), jam_capacity=[64, 128, 64][k])
        one.reset([names[i] for i in idx], seeds=[seeds[i] for i in idx], rng_mode="philox")
        s1 = one.run().summary_np()
        assert s1.tobytes() == s[idx].tobytes(), k
