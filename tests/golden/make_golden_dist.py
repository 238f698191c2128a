"""Live-mode distribution anchors from the UNMODIFIED reference (authoring container only), many seeds in parallel.

For every dispatcher it runs the reference's own ``run_dispatch_sim`` through ``oracle/ref_harness.py`` with the global
numpy / ``random`` streams seeded by ``seed`` and stores, per seed: reported tons, truck events, wrapper order count, jams,
road repairs, truck breakdowns, unrepairable trucks; and per minute the sum and the sum of squares over the seeds of the
reported production curve and of the waiting-truck count (so a test can compare curve means with a z-bound).

    python tests/golden/make_golden_dist.py north_pit_mine 1024
    python tests/golden/make_golden_dist.py ragged_mine 512

Output: tests/golden/live_distribution_<config>.npz (replaces the 48-/64-seed files of round 1).
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

DISPATCHERS = {
    "north_pit_mine": ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher",
                       "SPTFDispatcher", "SQDispatcher"],
    "ragged_mine": ["SQDispatcher", "SPTFDispatcher", "NearestDispatcher", "FixedGroupDispatcher"],
}
SEED0 = {"north_pit_mine": 0, "ragged_mine": 1000}
_CFG = None


def _init(cfg_name):
    global _CFG
    with open(os.path.join(HERE, "configs", cfg_name + ".json"), encoding="utf-8") as fh:
        _CFG = json.load(fh)


def _one(job):
    from oracle import ref_harness as rh
    from oracle import trace_format as tf
    d, seed = job
    r = rh.run_reference(_CFG, d, "live", seed=seed)
    k = r.trace["kind"]
    nt = r.n_ticks
    curve = np.zeros(int(np.ceil(_CFG["sim_time"])), dtype=np.float64)
    wait = np.zeros_like(curve)
    curve[:nt] = r.tick["produced_tons"]
    wait[:nt] = r.tick["waiting"]
    return (d, seed, r.error or "", float(r.reported_tons), int(r.n_truck_events), int(r.order_count),
            int((k == tf.K_JAM).sum()), int((k == tf.K_ROAD_REPAIR).sum()), int((k == tf.K_BREAKDOWN).sum()),
            int((k == tf.K_UNREPAIRABLE).sum()), curve, wait)


def main():
    cfg_name = sys.argv[1] if len(sys.argv) > 1 else "north_pit_mine"
    n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    nproc = int(sys.argv[3]) if len(sys.argv) > 3 else os.cpu_count()
    names = DISPATCHERS[cfg_name]
    jobs = [(d, SEED0[cfg_name] + s) for d in names for s in range(n_seeds)]
    t0 = time.time()
    out = {}
    cols = {d: {k: [] for k in ("seed", "tons", "events", "orders", "jams", "road_repairs", "breakdowns", "unrepairable")}
            for d in names}
    curves = {d: [0.0, 0.0, 0.0, 0.0] for d in names}
    errors = []
    with mp.Pool(nproc, initializer=_init, initargs=(cfg_name,)) as pool:
        for n, res in enumerate(pool.imap_unordered(_one, jobs, chunksize=4)):
            d, seed, err, tons, ev, orders, jams, rr, br, un, curve, wait = res
            if err:
                errors.append((d, seed, err))
                continue
            c = cols[d]
            for key, v in zip(("seed", "tons", "events", "orders", "jams", "road_repairs", "breakdowns", "unrepairable"),
                              (seed, tons, ev, orders, jams, rr, br, un)):
                c[key].append(v)
            a = curves[d]
            a[0] = a[0] + curve
            a[1] = a[1] + curve * curve
            a[2] = a[2] + wait
            a[3] = a[3] + wait * wait
            if n % 256 == 0:
                print("%d / %d  %.0fs" % (n, len(jobs), time.time() - t0), flush=True)
    for d in names:
        order = np.argsort(cols[d]["seed"])
        for key, v in cols[d].items():
            out["%s_%s" % (d, key)] = np.asarray(v, dtype=np.float64)[order]
        out[d + "_curve_sum"], out[d + "_curve_sumsq"], out[d + "_wait_sum"], out[d + "_wait_sumsq"] = curves[d]
        t = out[d + "_tons"]
        print(d, len(t), "tons %.1f +- %.1f" % (t.mean(), t.std()), "events %.1f" % out[d + "_events"].mean(),
              "jams %.2f" % out[d + "_jams"].mean(), flush=True)
    out["python"] = np.array(sys.version_info[:3])
    out["errors"] = np.array(["%s:%d:%s" % e for e in errors])
    np.savez_compressed(os.path.join(HERE, "live_distribution_%s.npz" % cfg_name), **out)
    print("done in %.0fs, %d errors" % (time.time() - t0, len(errors)))


if __name__ == "__main__":
    main()
