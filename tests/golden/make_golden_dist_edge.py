"""Live-mode distribution anchors for the ragged mine from the UNMODIFIED reference (authoring container only).

The ragged mine carries its own ``road_event_params`` (lambda_repair 0.02, repairs of ~20 minutes), so road repairs and
detours -- rare on north_pit -- happen dozens of times per episode: this pins the stochastic road path at the
distribution level.  64 seeds per dispatcher; stored: reported tons, truck events, jams, road repairs, truck breakdowns.

    python tests/golden/make_golden_dist_edge.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from oracle import trace_format as tf         # noqa: E402


def main():
    with open(os.path.join(HERE, "configs", "ragged_mine.json"), encoding="utf-8") as fh:
        cfg = json.load(fh)
    out = {}
    for d in ("SQDispatcher", "SPTFDispatcher", "NearestDispatcher", "FixedGroupDispatcher"):
        cols = {k: [] for k in ("tons", "events", "jams", "road_repairs", "breakdowns")}
        for seed in range(64):
            r = rh.run_reference(cfg, d, "live", seed=1000 + seed)
            assert r.error is None, r.error
            k = r.trace["kind"]
            cols["tons"].append(r.reported_tons)
            cols["events"].append(r.n_truck_events)
            cols["jams"].append(int((k == tf.K_JAM).sum()))
            cols["road_repairs"].append(int((k == tf.K_ROAD_REPAIR).sum()))
            cols["breakdowns"].append(int((k == tf.K_BREAKDOWN).sum()))
        for key, v in cols.items():
            out["%s_%s" % (d, key)] = np.array(v, dtype=np.float64)
        print(d, {k: (round(float(np.mean(v)), 2), round(float(np.std(v)), 2)) for k, v in cols.items()})
    np.savez_compressed(os.path.join(HERE, "live_distribution_ragged_mine.npz"), **out)


if __name__ == "__main__":
    main()
