"""GPU test of the tick recorder: the engine's outputs rendered into the reference's result
JSON must equal the JSON the unmodified reference's TickGenerator wrote for the same run."""
import gzip
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def _load(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as fh:
        return json.load(fh)


def _close(a, b, tol=1e-9):
    if isinstance(a, (int, float)) and isinstance(b, (int, float)) and not isinstance(a, bool):
        return abs(a - b) <= tol * max(1.0, abs(a), abs(b))
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(_close(x, y, tol) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_close(a[k], b[k], tol) for k in a)
    return a == b


def _compare(got, ref, skip_mine_keys=(), skip_summary=("avg_time_per_order",), tol=1e-9, skip_truck_states=()):
    got = json.loads(json.dumps(got))          # what a consumer of the file sees
    assert set(got.keys()) == set(ref.keys())
    problems = []
    for k in ref:
        if k == "summary":
            for f in ref[k]:
                if f in skip_summary:
                    continue
                if not _close(got[k][f], ref[k][f], tol):
                    problems.append(("summary", f, got[k][f] if not isinstance(got[k][f], list) else "list", ref[k][f] if not isinstance(ref[k][f], list) else "list"))
            continue
        g, r = got[k], ref[k]
        for sec in ("load_site_states", "dump_site_states", "shovel_states", "dumper_states"):
            if not _close(g[sec], r[sec], tol):
                bad = [n for n in r[sec] if not _close(g[sec].get(n), r[sec][n], tol)]
                problems.append((k, sec, bad[:2], g[sec].get(bad[0]) if bad else None, r[sec][bad[0]] if bad else None))
        for f in r["mine_states"]:
            if f in skip_mine_keys:
                continue
            if not _close(g["mine_states"][f], r["mine_states"][f], tol):
                problems.append((k, "mine_states", f, g["mine_states"][f], r["mine_states"][f]))
        for name, ts in r["truck_states"].items():
            if ts["state"] in skip_truck_states:
                continue
            if not _close(g["truck_states"][name], ts, 1e-9):
                problems.append((k, "truck", name, g["truck_states"][name], ts))
    return problems


def test_result_json_det_sq(north_pit):
    import openmines_b200 as om
    from openmines_b200.ticker import TickRecorder
    ref = _load("ticks_north_pit_mine_SQDispatcher_det_t120.json.gz")
    cfg = dict(north_pit, sim_time=120)
    mine = om.BatchedMine(cfg, 1)
    mine.reset("SQDispatcher", rng_mode="det")
    res = mine.run(trace_capacity=4096, site_ticks=True, unit_ticks=True, order_capacity=1024)
    rec = TickRecorder(mine, res, "SQDispatcher")
    ticks = rec.ticks(0)
    problems = _compare(ticks, ref)
    assert not problems, problems[:6]
    sd = rec.states_dict(ticks)
    assert sd["produced_tons_list"][-1] == ref["summary"]["produced_tons"] and len(sd["times"]) == 120
    path = rec.write_to_file(ticks, result_dir=str(__import__("tempfile").mkdtemp()))
    assert os.path.basename(path).startswith("MINE-NorthPitMine-ALGO-SQDispatcher-TIME-")
    assert json.load(open(path))["summary"]["total_order_count"] == ref["summary"]["total_order_count"]


def test_result_json_tape_sptf(north_pit):
    """Stochastic run (tape replay of the reference's draws), including the reference's dict-collision
    event counters, which the recorder re-derives from the event trace."""
    import openmines_b200 as om
    from openmines_b200.ticker import TickRecorder
    ref = _load("ticks_north_pit_mine_SPTFDispatcher_live_s11.json.gz")
    z = np.load(os.path.join(GOLDEN, "live_north_pit_mine_SPTFDispatcher_s11.npz"))
    mine = om.BatchedMine(north_pit, 1)
    mine.reset("SPTFDispatcher", rng_mode="tape", tapes=[(z["tape_sites"], z["tape_values"])])
    res = mine.run(trace_capacity=8192, site_ticks=True, unit_ticks=True, order_capacity=2048)
    ticks = TickRecorder(mine, res, "SPTFDispatcher").ticks(0)
    problems = _compare(ticks, ref)
    assert not problems, problems[:6]


def test_experiment_drivers(north_pit):
    """run_dispatch_sim / run_simulation / fleet ablation facades (cli/run.py:48-112,115-273,283-370,363-472)."""
    from openmines_b200 import experiments as ex
    import openmines_b200 as om
    cfg = dict(north_pit, sim_time=60)
    cfg["dispatcher"] = {"type": ["NaiveDispatcher", "SQDispatcher", "PPODispatcher"]}
    d = om.SQDispatcher()
    out = ex.run_dispatch_sim(d, cfg, n_seeds=4, stochastic=False)
    assert set(out["ticks"].keys()) == set(range(60)) | {"summary"}
    assert (out["summary"]["reported_tons"] == out["summary"]["reported_tons"][0]).all()
    assert d.total_order_count == out["ticks"]["summary"]["total_order_count"] > 71
    out["mine"].close()
    states = ex.run_simulation(cfg, stochastic=False)
    assert set(states) == {"NaiveDispatcher", "SQDispatcher"}
    assert len(states["SQDispatcher"]["produced_tons_list"]) == 60
    assert ex.fleet_sizes_for(10, 100) == [10, 20, 30, 40, 50, 60, 70, 80, 90, 100]
    c = ex.scaled_fleet_config(north_pit, 40)
    assert sum(t["count"] for t in c["charging_site"]["trucks"]) == 40
    res = ex.scene_based_fleet_ablation(cfg, 20, 20, dispatchers=["NaiveDispatcher", "SQDispatcher"], n_seeds=8)
    assert res["SQDispatcher"]["fleet_sizes"] == [20] and res["SQDispatcher"]["productions"][0] > res["NaiveDispatcher"]["productions"][0]
    # algo-based ablation (cli/run.py:363-472): ratio target / baseline per scene and fleet size; a directory of scene JSONs works too
    data = ex.algo_based_fleet_ablation([cfg, dict(cfg, mine={"name": "SecondScene"})], "NaiveDispatcher", "SQDispatcher",
                                        min_truck=15, max_truck=33, n_seeds=4, stochastic=False)
    assert set(data) == {north_pit["mine"]["name"], "SecondScene"}
    first = data[north_pit["mine"]["name"]]
    assert first["fleet_sizes"] == ex.fleet_sizes_for(15, 33) and len(first["ratios"]) == 10
    assert all(r > 1.0 for r in first["ratios"])                   # SQ out-produces Naive at every fleet size
    assert first["ratios"] == data["SecondScene"]["ratios"]        # det mode: same mine, same numbers
    assert abs(first["ratios"][3] - first["target"][3] / first["baseline"][3]) < 1e-15
    scene_dir = os.path.join(GOLDEN, "configs")
    by_dir = ex.algo_based_fleet_ablation(scene_dir, "NaiveDispatcher", "NearestDispatcher", min_truck=5, max_truck=5, n_seeds=2,
                                          stochastic=False)
    assert "north_pit_mine" in by_dir and by_dir["north_pit_mine"]["fleet_sizes"] == [5]


def test_result_json_tape_ragged_mine(ragged_mine):
    """TickGenerator parity on the ragged mine (L != D, 1..9 units per site, 130 trucks, frequent road repairs): tape
    replay of the reference's live SQ run (tests/golden/make_golden_ticks_edge.py)."""
    import openmines_b200 as om
    from openmines_b200.ticker import TickRecorder
    ref = _load("ticks_ragged_mine_SQDispatcher_live_s5.json.gz")
    z = np.load(os.path.join(GOLDEN, "live_ragged_mine_SQDispatcher_s5.npz"))
    mine = om.BatchedMine(ragged_mine, 1)
    mine.reset("SQDispatcher", rng_mode="tape", tapes=[(z["tape_sites"], z["tape_values"])])
    res = mine.run(trace_capacity=8192, site_ticks=True, unit_ticks=True, order_capacity=2048)
    ticks = TickRecorder(mine, res, "SQDispatcher").ticks(0)
    problems = _compare(ticks, ref)
    assert not problems, problems[:6]
