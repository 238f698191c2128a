"""GPU box: run a few Philox-mode instances with the draw-tape export and save, per instance, the tape and everything the
device produced (trace, tick arrays, summary, order log) to gpurun_out/device_tapes.npz.  The authoring container then
replays the tapes through the UNMODIFIED reference (tests/golden/make_golden_device_tapes.py) and commits the pair as a
fixture: the Philox kernels' trajectories reproduced by the reference itself."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import openmines_b200 as om
    from conftest import load_cfg
    out = {}
    cases = [("north_pit_mine", ["SQDispatcher", "SPTFDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher",
                                 "NaiveDispatcher", "ShortestTripDispatcher", "SQDispatcher"], 64),
             ("ragged_mine", ["SQDispatcher", "SPTFDispatcher", "NearestDispatcher", "ShortestTripDispatcher"], 128)]
    for cname, names, jam in cases:
        cfg = load_cfg(cname)
        n = len(names)
        mine = om.BatchedMine(cfg, n, jam_capacity=jam)
        seeds = np.arange(n, dtype=np.uint64) + 2024
        mine.reset(names, seeds=seeds, rng_mode="philox")
        res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=4096, draw_capacity=mine.draw_tape_slots())
        s = res.summary_np()
        assert (s["err"] == 0).all()
        ticks = res.ticks_np()
        for i in range(n):
            key = "%s_%d_%s" % (cname, i, names[i])
            sites, values = res.draw_tape(i)
            tr, _ = res.trace_np(i)
            nrec = int(np.nonzero((tr["time"] != 0) | (tr["kind"] != 0) | (tr["truck"] != 0))[0].max()) + 1
            out[key + "_tape_sites"], out[key + "_tape_values"] = sites, values
            out[key + "_trace"] = tr[:nrec]
            out[key + "_produced"] = ticks["produced_tons"][i]
            out[key + "_waiting"] = ticks["waiting"][i]
            out[key + "_summary"] = s[i:i + 1]
            row = res.orders_np()[i]
            out[key + "_orders"] = row[row >= 0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "device_tapes.npz"), **out)
    print("saved", len(out), "arrays")


if __name__ == "__main__":
    main()
