"""Goldens of ShortestTripDispatcher run AS A DISPATCHER by the unmodified reference (authoring container only): det on
north_pit_mine and the ragged mine, one live run with its RNG tape on north_pit.  The device engine carries it as its 7th
device function (OM_DISP_SHORTESTTRIP); round 1 pinned it to the reference only through the RL teacher's suggestions.

    python tests/golden/make_golden_teacher.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_harness as rh  # noqa: E402
from make_golden import save_result  # noqa: E402


def main():
    d = "ShortestTripDispatcher"
    for cname in ("north_pit_mine", "ragged_mine"):
        with open(os.path.join(HERE, "configs", cname + ".json"), encoding="utf-8") as fh:
            cfg = json.load(fh)
        r = rh.run_reference(cfg, d, "det", seed=0)
        save_result(os.path.join(HERE, "det_%s_%s_s0.npz" % (cname, d)), r)
        print("det", cname, r.reported_tons, r.n_truck_events, r.order_count, r.error)
        r = rh.run_reference(cfg, d, "live", seed=17)
        s, v = r.tape
        save_result(os.path.join(HERE, "live_%s_%s_s17.npz" % (cname, d)), r, {"tape_sites": s, "tape_values": v})
        print("live", cname, r.reported_tons, r.n_truck_events, len(v), r.error)


if __name__ == "__main__":
    main()
