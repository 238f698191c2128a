"""Generates the committed golden vectors from the UNMODIFIED reference at /root/reference.

Run in the authoring container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

For each (config, dispatcher, rng mode, seed) it runs the reference's own ``run_dispatch_sim``
(cli/run.py:48-112) through ``oracle/ref_harness.py`` (SimPy stand-in + import stubs, nothing of the
reference modified) and stores the canonical event trace, per-tick KPI arrays, final per-truck /
per-shovel / per-dumper values, the dispatcher's order lists and -- for live runs -- the RNG tape.
The mine configs themselves (input data of the benchmark) are stored re-serialised under
``tests/golden/configs`` so that tests and bench.py can run where /root/reference is absent.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

REF_CONF = "/root/reference/openmines/src/conf"
DISPATCHERS = ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher",
               "SPTFDispatcher", "SQDispatcher"]


def save_result(path, r, extra=None):
    d = {"trace": r.trace, "reported_tons": np.float64(r.reported_tons), "dumped_tons": np.float64(r.dumped_tons),
         "n_truck_events": np.int64(r.n_truck_events), "order_count": np.int64(r.order_count),
         "n_steps": np.int64(r.n_steps), "n_ticks": np.int64(r.n_ticks), "python": np.array(r.python),
         "error": np.array(r.error or "")}
    for k, v in r.tick.items():
        d["tick_" + k] = v
    for k, v in r.final.items():
        d["final_" + k] = v
    for k, v in r.orders.items():
        d["orders_" + k] = v
    if extra:
        d.update(extra)
    np.savez_compressed(path, **d)


def main():
    os.makedirs(os.path.join(HERE, "configs"), exist_ok=True)
    cfgs = {}
    for name in ("north_pit_mine", "north_pit_mine_short"):
        cfg = rh.load_config(os.path.join(REF_CONF, name + ".json"))
        cfgs[name] = cfg
        with open(os.path.join(HERE, "configs", name + ".json"), "w", encoding="utf-8") as fh:
            json.dump(cfg, fh, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    t0 = time.time()
    # deterministic mode: every dispatcher on both configs (Random: keyed Philox, seeds 0 and 7)
    for cname, cfg in cfgs.items():
        for d in DISPATCHERS:
            seeds = (0, 7) if d == "RandomDispatcher" else (0,)
            for seed in seeds:
                r = rh.run_reference(cfg, d, "det", seed=seed)
                save_result(os.path.join(HERE, "det_%s_%s_s%d.npz" % (cname, d, seed)), r)
                print("det", cname, d, seed, r.reported_tons, r.n_truck_events, r.order_count, r.error)
    # live mode with tapes: every dispatcher on north_pit (seed 11), short config (seed 3, sim 120)
    for d in DISPATCHERS:
        r = rh.run_reference(cfgs["north_pit_mine"], d, "live", seed=11)
        s, v = r.tape
        save_result(os.path.join(HERE, "live_north_pit_mine_%s_s11.npz" % d), r, {"tape_sites": s, "tape_values": v})
        print("live", d, r.reported_tons, r.n_truck_events, len(v), r.error)
    for d in ("RandomDispatcher", "SQDispatcher"):
        r = rh.run_reference(cfgs["north_pit_mine_short"], d, "live", seed=3)
        s, v = r.tape
        save_result(os.path.join(HERE, "live_north_pit_mine_short_%s_s3.npz" % d), r, {"tape_sites": s, "tape_values": v})
    # live-mode distribution anchors: reported tons of the reference for seeds 0..47 per dispatcher
    dist = {}
    for d in DISPATCHERS:
        tons, ev, jams = [], [], []
        for seed in range(48):
            r = rh.run_reference(cfgs["north_pit_mine"], d, "live", seed=seed)
            tons.append(r.reported_tons)
            ev.append(r.n_truck_events)
            jams.append(int((r.trace["kind"] == 12).sum()))
        dist[d + "_tons"] = np.array(tons)
        dist[d + "_events"] = np.array(ev)
        dist[d + "_jams"] = np.array(jams)
        print("dist", d, np.mean(tons), np.std(tons), np.mean(ev), np.mean(jams))
    np.savez_compressed(os.path.join(HERE, "live_distribution_north_pit_mine.npz"), **dist)
    print("done in %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
