"""Synthetic mine configs in the reference's JSON schema (conf/north_pit_mine.json).

``scaled_mine()`` is BASELINE.json configs[4] / SURVEY.md section 8d "C5": 500 trucks, 20 load sites x 4
shovels, 10 dump sites x 8 dumpers, generated with ``np.random.default_rng(20240622)``.
"""
import numpy as np


def scaled_mine(n_trucks=(63, 204, 233), n_load=20, n_dump=10, seed=20240622, sim_time=240):
    rng = np.random.default_rng(seed)
    l2d = np.round(rng.uniform(1, 35, size=(n_load, n_dump)), 2)
    d2l = l2d.copy()
    mask = rng.uniform(size=d2l.shape) < 0.15
    d2l[mask] = np.round(d2l[mask] * rng.uniform(1, 6, size=int(mask.sum())), 2)
    c2l = np.round(rng.uniform(2, 5, size=n_load), 2)
    caps = (77, 35, 55)
    names = ("OfficalTruck", "CLTruck", "XHTruck")
    cfg = {
        "mine": {"name": "SyntheticScaledMine"},
        "dispatcher": {"type": ["SQDispatcher"]},
        "charging_site": {"name": "ScaledChargingSite", "position": [0, 0],
                          "trucks": [{"type": names[k], "count": int(n_trucks[k]), "capacity": caps[k], "speed": 25}
                                     for k in range(3)]},
        "load_sites": [], "dump_sites": [],
        "road": {"l2d_road_matrix": l2d.tolist(), "d2l_road_matrix": d2l.tolist(),
                 "charging_to_load_road_matrix": c2l.tolist()},
        "sim_time": sim_time,
    }
    for i in range(n_load):
        y = 0.05 + 0.9 * i / max(n_load - 1, 1)
        shovels = [{"name": "LS%d-Shovel-1" % (i + 1), "tons": 20.32, "cycle_time": 1.5, "position_offset": [0.1, 0.1]}]
        shovels += [{"name": "LS%d-Shovel-%d" % (i + 1, k + 2), "tons": 2.25, "cycle_time": 1,
                     "position_offset": [0.1, 0.05 * k]} for k in range(3)]
        cfg["load_sites"].append({"name": "ScaledLoadSite%d" % (i + 1), "position": [0.1, y], "shovels": shovels,
                                  "parkinglot": {"name": "ScaledLoadSite%d-ParkingLot" % (i + 1),
                                                 "position_offset": [-0.08, 0]}})
    for j in range(n_dump):
        y = 0.05 + 0.9 * j / max(n_dump - 1, 1)
        cfg["dump_sites"].append({"name": "ScaledDumpSite%d" % (j + 1), "position": [0.8, y],
                                  "dumpers": [{"count": 8, "cycle_time": 1, "position_offset": [0.0, 0.05]}],
                                  "parkinglot": {"name": "ScaledDumpSite%d-ParkingLot" % (j + 1),
                                                 "position_offset": [0.06, 0.06]}})
    return cfg
