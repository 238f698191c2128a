"""Batched counterparts of the reference's experiment drivers (openmines/src/cli/run.py).

* ``run_dispatch_sim(dispatcher, config, ...)``  -- cli/run.py:48-112: one dispatcher on one mine; here N seeds at
  once; returns the reference's ticks dict for the first seed plus per-seed summaries.
* ``run_simulation(config)``                     -- cli/run.py:115-273 without the plotting: every dispatcher named in
  ``config['dispatcher']['type']`` -> ``states_dict`` (what ``Charter.draw`` consumes).
* ``scene_based_fleet_ablation(config, lo, hi)`` -- cli/run.py:283-370: production vs fleet size for every dispatcher,
  10 fleet sizes, truck types scaled by their original ratios; each point is the mean over ``n_seeds`` instances.
* ``algo_based_fleet_ablation(scenes, baseline, target, lo, hi)`` -- cli/run.py:363-472: for every scene (a directory of
  mine JSONs, or a list of paths / dicts) the production ratio target / baseline over the same 10 fleet sizes; one
  batch per (scene, fleet size) holds both dispatchers' instances.
"""
import copy
import glob
import math
import os

import numpy as np

from .config import load_config
from .dispatcher import BUILTIN, by_name
from .engine import BatchedMine, DISPATCHER_IDS
from .ticker import TickRecorder


def run_dispatch_sim(dispatcher, config_file, n_seeds=1, stochastic=True, seeds=None, device=0, want_ticks=True):
    """N independent runs of ``dispatcher`` (a BaseDispatcher object or a built-in name) on the mine."""
    cfg = load_config(config_file)
    if isinstance(dispatcher, str):
        dispatcher = by_name(dispatcher)
    seeds = np.arange(n_seeds, dtype=np.uint64) if seeds is None else np.asarray(seeds, dtype=np.uint64)
    mine = BatchedMine(cfg, len(seeds), device=device)
    mine.reset(dispatcher, seeds=seeds, rng_mode="philox" if stochastic else "det")
    trace_cap = 16384 if want_ticks else 0
    res = mine.run(trace_capacity=trace_cap, site_ticks=want_ticks, unit_ticks=want_ticks,
                   order_capacity=8192 if want_ticks else 0)
    summary = res.summary_np()
    ticks = None
    if want_ticks:
        ticks = TickRecorder(mine, res, dispatcher.name).ticks(0)
        dispatcher.record_orders(res.orders_np()[0], sptf_double_log=(dispatcher.name == "SPTFDispatcher"))
    out = {"ticks": ticks, "summary": summary, "mine": mine, "result": res}
    return out


def run_simulation(config_file, n_seeds=1, stochastic=True, device=0):
    cfg = load_config(config_file)
    states = {}
    for name in cfg["dispatcher"]["type"]:
        if name not in BUILTIN:
            continue            # PPO / LLM / Gurobi dispatchers are out of scope
        r = run_dispatch_sim(by_name(name), cfg, n_seeds=n_seeds, stochastic=stochastic, device=device)
        rec = TickRecorder(r["mine"], r["result"], name)
        states[name] = rec.states_dict(r["ticks"])
        r["mine"].close()
    return states


def fleet_sizes_for(min_truck, max_truck):
    lo, hi = int(min_truck), int(max_truck)
    if hi < lo:
        lo, hi = hi, lo
    if hi == lo:
        return [lo]
    step = (hi - lo) / 9
    return [int(math.floor(lo + i * step)) for i in range(10)]


def scaled_fleet_config(config, fleet_size):
    """cli/run.py:330-341: scale every truck type by its original share, leftovers go to the last type."""
    cfg = copy.deepcopy(config)
    counts = [t["count"] for t in config["charging_site"]["trucks"]]
    total = sum(counts)
    assigned = 0
    for i, t in enumerate(cfg["charging_site"]["trucks"]):
        t["count"] = int(math.floor(fleet_size * (counts[i] / total)))
        assigned += t["count"]
    if fleet_size - assigned > 0:
        cfg["charging_site"]["trucks"][-1]["count"] += fleet_size - assigned
    return cfg


def scene_based_fleet_ablation(config_file, min_truck, max_truck, dispatchers=None, n_seeds=32, stochastic=True, device=0):
    cfg = load_config(config_file)
    names = dispatchers or [n for n in cfg["dispatcher"]["type"] if n in DISPATCHER_IDS]
    results = {n: {"fleet_sizes": [], "productions": [], "std": []} for n in names}
    for fs in fleet_sizes_for(min_truck, max_truck):
        c = scaled_fleet_config(cfg, fs)
        c["charging_site"]["trucks"] = [t for t in c["charging_site"]["trucks"] if t["count"] > 0]
        mine = BatchedMine(c, n_seeds * len(names), device=device)
        ids = np.repeat([DISPATCHER_IDS[n] for n in names], n_seeds).astype(np.int32)
        mine.reset(ids, seeds=np.tile(np.arange(n_seeds, dtype=np.uint64), len(names)),
                   rng_mode="philox" if stochastic else "det")
        s = mine.run(ticks=True).summary_np().reshape(len(names), n_seeds)
        for k, n in enumerate(names):
            results[n]["fleet_sizes"].append(fs)
            results[n]["productions"].append(float(s[k]["reported_tons"].mean()))
            results[n]["std"].append(float(s[k]["reported_tons"].std()))
        mine.close()
    return results


def algo_based_fleet_ablation(scenes, baseline, target, min_truck=1, max_truck=160, n_seeds=32, stochastic=True, device=0):
    """``scenes``: directory with ``*.json`` mine configs (the reference's ``-d``), or a list of paths / config dicts.
    Returns ``{scene name: {"fleet_sizes": [...], "ratios": [...], "baseline": [...], "target": [...]}}`` -- the
    ``scenes_data`` that ``Charter.draw_algo_based_fleet_ablation_experiment`` plots (ratio 0 where the baseline produced
    nothing, cli/run.py:461-463) plus the two mean productions."""
    if isinstance(scenes, (str, os.PathLike)):
        paths = sorted(glob.glob(os.path.join(str(scenes), "*.json")))
        scenes = [(os.path.splitext(os.path.basename(p))[0], load_config(p)) for p in paths]
    else:
        named = []
        for k, sc in enumerate(scenes):
            cfg = load_config(sc)
            name = os.path.splitext(os.path.basename(sc))[0] if isinstance(sc, (str, os.PathLike)) else \
                cfg.get("mine", {}).get("name", "scene%d" % k)
            named.append((name, cfg))
        scenes = named
    ids = np.repeat([DISPATCHER_IDS[by_name(baseline).name], DISPATCHER_IDS[by_name(target).name]], n_seeds).astype(np.int32)
    seeds = np.tile(np.arange(n_seeds, dtype=np.uint64), 2)
    out = {}
    for name, cfg in scenes:
        data = {"fleet_sizes": [], "ratios": [], "baseline": [], "target": []}
        for fs in fleet_sizes_for(min_truck, max_truck):
            c = scaled_fleet_config(cfg, fs)
            c["charging_site"]["trucks"] = [t for t in c["charging_site"]["trucks"] if t["count"] > 0]
            mine = BatchedMine(c, 2 * n_seeds, device=device)
            mine.reset(ids, seeds=seeds, rng_mode="philox" if stochastic else "det")
            tons = mine.run(ticks=True).summary_np()["reported_tons"].reshape(2, n_seeds).mean(axis=1)
            mine.close()
            data["fleet_sizes"].append(fs)
            data["baseline"].append(float(tons[0]))
            data["target"].append(float(tons[1]))
            data["ratios"].append(float(tons[1] / tons[0]) if tons[0] > 1e-9 else 0.0)
        out[name] = data
    return out
