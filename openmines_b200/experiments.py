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
from .trace import IERR_NAMES

OM_IERR_TRACE_OVERFLOW = 5


def default_jam_capacity(config):
    """Jam-pool size for a mine: 64 entries hold every simultaneously active jam of the 71-truck north_pit mine with margin
    (tools/jam_capacity_sweep.py); more trucks on the same roads keep more jams alive, so the pool grows with the fleet."""
    n_trucks = sum(int(t["count"]) for t in load_config(config)["charging_site"]["trucks"])
    return int(min(4096, max(64, 64 * ((n_trucks + 63) // 64) * 2)))


def check_instance_errors(summary, what):
    """Raise if any instance stopped on an engine error (jam-pool overflow ...) or where the reference itself raises (SQ's
    KeyError, an out-of-range order): their truncated tonnage must not be averaged into a curve."""
    bad = np.nonzero((summary["err"] != 0) & (summary["err"] != OM_IERR_TRACE_OVERFLOW))[0]
    if len(bad):
        codes = sorted(set(int(e) for e in summary["err"][bad]))
        raise RuntimeError("%s: %d of %d instances stopped early: %s (first: instance %d)" % (
            what, len(bad), len(summary), "; ".join("%d = %s" % (c, IERR_NAMES.get(c, "?")) for c in codes), int(bad[0])))


def run_dispatch_sim(dispatcher, config_file, n_seeds=1, stochastic=True, seeds=None, device=0, want_ticks=True,
                     jam_capacity=None):
    """N independent runs of ``dispatcher`` (a BaseDispatcher object or a built-in name) on the mine.  A user-defined
    ``BaseDispatcher`` subclass (no device function) is routed to the host-decision path (``run_external``)."""
    cfg = load_config(config_file)
    if isinstance(dispatcher, str):
        dispatcher = by_name(dispatcher)
    seeds = np.arange(n_seeds, dtype=np.uint64) if seeds is None else np.asarray(seeds, dtype=np.uint64)
    jam_capacity = jam_capacity or default_jam_capacity(cfg)
    if int(getattr(dispatcher, "device_id", 255)) == DISPATCHER_IDS["external"]:
        from .external import run_external
        ext = run_external(dispatcher, cfg, n_instances=len(seeds), seeds=seeds, rng_mode="philox" if stochastic else "det",
                           device=device, jam_capacity=jam_capacity)
        bad = np.nonzero((ext["headers"]["err"] != 0) & (ext["headers"]["err"] != OM_IERR_TRACE_OVERFLOW))[0]
        if len(bad):
            raise RuntimeError("run_dispatch_sim(%s): instance %d stopped early: %s" % (
                getattr(dispatcher, "name", type(dispatcher).__name__), int(bad[0]),
                IERR_NAMES.get(int(ext["headers"]["err"][bad[0]]), "?")))
        return {"ticks": None, "summary": None, "headers": ext["headers"], "produced_tons": ext["produced_tons"],
                "mine": ext["mine"], "result": None, "dispatchers": ext["dispatchers"]}
    mine = BatchedMine(cfg, len(seeds), device=device, jam_capacity=jam_capacity)
    mine.reset(dispatcher, seeds=seeds, rng_mode="philox" if stochastic else "det")
    trace_cap = 16384 if want_ticks else 0
    res = mine.run(trace_capacity=trace_cap, site_ticks=want_ticks, unit_ticks=want_ticks,
                   order_capacity=8192 if want_ticks else 0)
    summary = res.summary_np()
    check_instance_errors(summary, "run_dispatch_sim(%s)" % dispatcher.name)
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


def scene_based_fleet_ablation_serial(config_file, min_truck, max_truck, dispatchers=None, n_seeds=32, stochastic=True,
                                      device=0, jam_capacity=None):
    """One launch per fleet size (the round-1 driver; kept as the baseline the one-launch sweep is measured against)."""
    cfg = load_config(config_file)
    names = dispatchers or [n for n in cfg["dispatcher"]["type"] if n in DISPATCHER_IDS]
    results = {n: {"fleet_sizes": [], "productions": [], "std": []} for n in names}
    for fs in fleet_sizes_for(min_truck, max_truck):
        c = scaled_fleet_config(cfg, fs)
        c["charging_site"]["trucks"] = [t for t in c["charging_site"]["trucks"] if t["count"] > 0]
        mine = BatchedMine(c, n_seeds * len(names), device=device, jam_capacity=jam_capacity or default_jam_capacity(c))
        ids = np.repeat([DISPATCHER_IDS[n] for n in names], n_seeds).astype(np.int32)
        mine.reset(ids, seeds=np.tile(np.arange(n_seeds, dtype=np.uint64), len(names)),
                   rng_mode="philox" if stochastic else "det")
        s = mine.run(ticks=True).summary_np()
        check_instance_errors(s, "scene_based_fleet_ablation(fleet size %d)" % fs)
        s = s.reshape(len(names), n_seeds)
        for k, n in enumerate(names):
            results[n]["fleet_sizes"].append(fs)
            results[n]["productions"].append(float(s[k]["reported_tons"].mean()))
            results[n]["std"].append(float(s[k]["reported_tons"].std()))
        mine.close()
    return results


def _fleet_configs(cfg, sizes):
    out = []
    for fs in sizes:
        c = scaled_fleet_config(cfg, fs)
        c["charging_site"]["trucks"] = [t for t in c["charging_site"]["trucks"] if t["count"] > 0]
        out.append(c)
    return out


def scene_based_fleet_ablation(config_file, min_truck, max_truck, dispatchers=None, n_seeds=32, stochastic=True, device=0,
                               jam_capacity=None, seeds=None):
    """cli/run.py:283-362 as ONE batch: 10 fleet sizes x dispatchers x seeds are the instances of a single mixed-config
    ``BatchedMine`` (``om_create_multi``) and one ``om_run`` launch simulates them all.  Returns, per dispatcher,
    ``{"fleet_sizes", "productions" (mean reported tons over the seeds; with ``stochastic=False`` and one seed exactly the
    value the reference plots), "std"}``."""
    cfg = load_config(config_file)
    names = dispatchers or [n for n in cfg["dispatcher"]["type"] if n in DISPATCHER_IDS]
    sizes = fleet_sizes_for(min_truck, max_truck)
    configs = _fleet_configs(cfg, sizes)
    seeds = np.arange(n_seeds, dtype=np.uint64) if seeds is None else np.asarray(seeds, dtype=np.uint64)
    n_seeds = len(seeds)
    # instance order: fleet size (slowest), dispatcher, seed
    F, K = len(sizes), len(names)
    config_ids = np.repeat(np.arange(F, dtype=np.int32), K * n_seeds)
    disp_ids = np.tile(np.repeat([DISPATCHER_IDS[n] for n in names], n_seeds), F).astype(np.int32)
    all_seeds = np.tile(seeds, F * K)
    jams = [jam_capacity or default_jam_capacity(c) for c in configs]
    mine = BatchedMine(configs, F * K * n_seeds, device=device, jam_capacity=jams, config_ids=config_ids)
    mine.reset(disp_ids, seeds=all_seeds, rng_mode="philox" if stochastic else "det")
    s = mine.run(ticks=True).summary_np()
    mine.close()
    check_instance_errors(s, "scene_based_fleet_ablation")
    tons = s["reported_tons"].reshape(F, K, n_seeds)
    results = {}
    for k, n in enumerate(names):
        results[n] = {"fleet_sizes": list(sizes), "productions": [float(x) for x in tons[:, k].mean(axis=1)],
                      "std": [float(x) for x in tons[:, k].std(axis=1)]}
    return results


def algo_based_fleet_ablation(scenes, baseline, target, min_truck=1, max_truck=160, n_seeds=32, stochastic=True, device=0,
                              jam_capacity=None):
    """cli/run.py:363-472 as ONE batch.  ``scenes``: directory with ``*.json`` mine configs (the reference's ``-d``), or a list
    of paths / config dicts.  Every (scene, fleet size) pair is one config of a mixed-config ``BatchedMine``; both dispatchers'
    seeds of all of them run in one launch.  Returns ``{scene name: {"fleet_sizes": [...], "ratios": [...], "baseline": [...],
    "target": [...]}}`` -- the ``scenes_data`` that ``Charter.draw_algo_based_fleet_ablation_experiment`` plots (ratio 0 where
    the baseline produced nothing, cli/run.py:461-463) plus the two mean productions."""
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
    sizes = fleet_sizes_for(min_truck, max_truck)
    F, NSC = len(sizes), len(scenes)
    configs = []
    for _, cfg in scenes:
        configs += _fleet_configs(cfg, sizes)
    pair = np.repeat([DISPATCHER_IDS[by_name(baseline).name], DISPATCHER_IDS[by_name(target).name]], n_seeds).astype(np.int32)
    config_ids = np.repeat(np.arange(NSC * F, dtype=np.int32), 2 * n_seeds)
    disp_ids = np.tile(pair, NSC * F)
    seeds = np.tile(np.arange(n_seeds, dtype=np.uint64), 2 * NSC * F)
    jams = [jam_capacity or default_jam_capacity(c) for c in configs]
    mine = BatchedMine(configs, len(config_ids), device=device, jam_capacity=jams, config_ids=config_ids)
    mine.reset(disp_ids, seeds=seeds, rng_mode="philox" if stochastic else "det")
    s = mine.run(ticks=True).summary_np()
    mine.close()
    check_instance_errors(s, "algo_based_fleet_ablation")
    tons = s["reported_tons"].reshape(NSC, F, 2, n_seeds).mean(axis=3)
    out = {}
    for k, (name, _) in enumerate(scenes):
        b, t = tons[k, :, 0], tons[k, :, 1]
        out[name] = {"fleet_sizes": list(sizes), "baseline": [float(x) for x in b], "target": [float(x) for x in t],
                     "ratios": [float(t[i] / b[i]) if b[i] > 1e-9 else 0.0 for i in range(F)]}
    return out
