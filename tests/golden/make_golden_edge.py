"""Edge-case golden vectors from the UNMODIFIED reference (authoring container only; see make_golden.py).

Two synthetic mines in the reference's JSON schema exercise what north_pit does not:

* ``ragged_mine``: L = 3 != D = 2, 1 / 4 / 7 shovels and 1 / 9 dumpers per site (lane-parallel unit picks with one
  and with many units), 130 trucks of four types (more than 96 = three scan slots per lane, not a multiple of 32,
  a fractional capacity, two speeds), asymmetric road matrices, sim_time 90;
* ``one_truck_mine``: the smallest legal mine (one truck, one shovel, one dumper) -- SURVEY.md section 8c anchor:
  2 + 8/cycle event records.

    python tests/golden/make_golden_edge.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from make_golden import save_result, DISPATCHERS   # noqa: E402


def _site(name, pos, units_key, units):
    return {"name": name, "position": pos, units_key: units,
            "parkinglot": {"name": name + " Parking Lot", "position_offset": [0.06, 0.06]}}


def ragged_mine():
    shovels = lambda site, n: [{"name": "Ragged-LS%d-Shovel-%d" % (site, k + 1), "tons": [2.25, 20.32, 5.5][k % 3],
                                "cycle_time": [1, 1.5, 0.75][k % 3], "position_offset": [0.1, 0.03 * k]} for k in range(n)]
    return {
        "mine": {"name": "RaggedMine"},
        "dispatcher": {"type": list(DISPATCHERS)},
        "charging_site": {"name": "RaggedChargingSite", "position": [0, 0],
                          "trucks": [{"type": "A", "count": 41, "capacity": 77, "speed": 25},
                                     {"type": "B", "count": 37, "capacity": 35, "speed": 30},
                                     {"type": "C", "count": 29, "capacity": 42.5, "speed": 25},
                                     {"type": "D", "count": 23, "capacity": 55, "speed": 30}]},
        "load_sites": [_site("RaggedLoadSite%d" % (i + 1), [0.1, 0.2 + 0.3 * i], "shovels", shovels(i + 1, n))
                       for i, n in enumerate((1, 4, 7))],
        "dump_sites": [_site("RaggedDumpSite%d" % (j + 1), [0.8, 0.3 + 0.4 * j], "dumpers",
                             [{"count": n, "cycle_time": [1, 1.25][j], "position_offset": [0.0, 0.05]}])
                       for j, n in enumerate((1, 9))],
        "road": {"l2d_road_matrix": [[3.1, 7.45], [5.0, 2.2], [9.75, 4.4]],
                 "d2l_road_matrix": [[3.1, 11.0], [6.5, 2.2], [9.75, 5.05]],
                 "charging_to_load_road_matrix": [2.4, 3.3, 1.9],
                 "road_event_params": {"lambda_repair": 0.02, "mu_repair_duration": 20, "sigma_repair_duration": 5,
                                       "mu_punish_distance": 0.4, "sigma_punish_distance": 0.1}},
        "sim_time": 90,
    }


def one_truck_mine():
    return {
        "mine": {"name": "OneTruckMine"},
        "dispatcher": {"type": ["NaiveDispatcher"]},
        "charging_site": {"name": "OneChargingSite", "position": [0, 0],
                          "trucks": [{"type": "Solo", "count": 1, "capacity": 40, "speed": 20}]},
        "load_sites": [_site("OneLoadSite", [0.1, 0.5], "shovels",
                             [{"name": "One-Shovel", "tons": 8, "cycle_time": 1.25, "position_offset": [0.1, 0.1]}])],
        "dump_sites": [_site("OneDumpSite", [0.8, 0.5], "dumpers",
                             [{"count": 1, "cycle_time": 2, "position_offset": [0.0, 0.05]}])],
        "road": {"l2d_road_matrix": [[4.0]], "d2l_road_matrix": [[4.5]], "charging_to_load_road_matrix": [1.5]},
        "sim_time": 120,
    }


def integer_mine():
    """Every duration is a whole number of minutes (30 km/h, distances in half kilometres, 30 t trucks on 10 t/min
    shovels): truck events, Releases and the per-minute monitor block tie at integer timestamps all the time -- the
    SimPy tie-breaking stress test.  Built from the reference's own short config."""
    with open(os.path.join(HERE, "configs", "north_pit_mine_short.json"), encoding="utf-8") as fh:
        c = json.load(fh)
    c["mine"]["name"] = "IntegerMine"
    for k, t in enumerate(c["charging_site"]["trucks"]):
        t["speed"], t["capacity"], t["count"] = 30, 30, 2 + k
    for ls in c["load_sites"]:
        for sh in ls["shovels"]:
            sh["tons"], sh["cycle_time"] = 10, 1
    L, D = len(c["load_sites"]), len(c["dump_sites"])
    c["road"]["charging_to_load_road_matrix"] = [1.0 + 0.5 * (i % 2) for i in range(L)]
    c["road"]["l2d_road_matrix"] = [[2.0 + 0.5 * ((i + j) % 3) for j in range(D)] for i in range(L)]
    c["road"]["d2l_road_matrix"] = [[2.0 + 0.5 * ((i + 2 * j) % 3) for j in range(D)] for i in range(L)]
    return c


def main():
    cfg = integer_mine()
    with open(os.path.join(HERE, "configs", "integer_mine.json"), "w", encoding="utf-8") as fh:
        json.dump(cfg, fh, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    for d in DISPATCHERS:
        r = rh.run_reference(cfg, d, "det", seed=0)
        save_result(os.path.join(HERE, "det_integer_mine_%s_s0.npz" % d), r)
        print("det integer", d, r.reported_tons, r.n_truck_events, r.order_count, r.error)
    for name, cfg in (("ragged_mine", ragged_mine()), ("one_truck_mine", one_truck_mine())):
        with open(os.path.join(HERE, "configs", name + ".json"), "w", encoding="utf-8") as fh:
            json.dump(cfg, fh, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    cfg = ragged_mine()
    for d in DISPATCHERS:
        r = rh.run_reference(cfg, d, "det", seed=0)
        save_result(os.path.join(HERE, "det_ragged_mine_%s_s0.npz" % d), r)
        print("det ragged", d, r.reported_tons, r.n_truck_events, r.order_count, r.error)
    for d in ("SQDispatcher", "SPTFDispatcher", "RandomDispatcher"):
        r = rh.run_reference(cfg, d, "live", seed=5)
        s, v = r.tape
        save_result(os.path.join(HERE, "live_ragged_mine_%s_s5.npz" % d), r, {"tape_sites": s, "tape_values": v})
        print("live ragged", d, r.reported_tons, r.n_truck_events, len(v), r.error)
    cfg = one_truck_mine()
    r = rh.run_reference(cfg, "NaiveDispatcher", "det", seed=0)
    save_result(os.path.join(HERE, "det_one_truck_mine_NaiveDispatcher_s0.npz"), r)
    print("det one", r.reported_tons, r.n_truck_events, r.order_count, r.n_steps, r.error)
    r = rh.run_reference(cfg, "NaiveDispatcher", "live", seed=2)
    s, v = r.tape
    save_result(os.path.join(HERE, "live_one_truck_mine_NaiveDispatcher_s2.npz"), r, {"tape_sites": s, "tape_values": v})
    print("live one", r.reported_tons, r.n_truck_events, len(v), r.error)


if __name__ == "__main__":
    main()
