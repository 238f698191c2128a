"""Warp-convergence probes at scale: runs the C3 shard (16384 live SQ/SPTF instances), the C2 det shard and a C5 shard once
under a library built with -DOM_CONV_PROBES (OM_LIB_PATH=...) and prints how many times a warp-uniform continuation was
entered by a partial warp (must be 0: the engine's redundant all-lane stores are only safe when the warp is converged)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import openmines_b200 as om
    from openmines_b200.synthetic import scaled_mine
    import sanitize_target
    cfg = om.builtin_config("north_pit_mine")
    n = 16384
    mine = om.BatchedMine(cfg, n)
    mine.reset(np.where(np.arange(n) % 2 == 0, 5, 4).astype(np.int32), seeds=np.arange(n) // 2, rng_mode="philox")
    s = mine.run().summary_np()
    print("c3: events", int(s["n_truck_events"].sum()), "errors", int((s["err"] != 0).sum()))
    sanitize_target.bounds_report()
    mine = om.BatchedMine(cfg, 6144)
    mine.reset(np.repeat(np.arange(6, dtype=np.int32), 1024), seeds=np.tile(np.arange(1024), 6), rng_mode="det")
    s = mine.run(trace_capacity=4096).summary_np()
    print("c2 (trace variant): events", int(s["n_truck_events"].sum()), "errors", int((s["err"] != 0).sum()))
    sanitize_target.bounds_report()
    mine = om.BatchedMine(scaled_mine(), 256, jam_capacity=512)
    mine.reset("SQDispatcher", seeds=np.arange(256), rng_mode="philox")
    s = mine.run().summary_np()
    print("c5 (split): events", int(s["n_truck_events"].sum()), "errors", int((s["err"] != 0).sum()))
    sanitize_target.bounds_report()


if __name__ == "__main__":
    main()
