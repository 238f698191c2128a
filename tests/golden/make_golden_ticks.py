"""Golden result JSONs (TickGenerator output, utils/ticker.py:42-470) recorded from the UNMODIFIED reference
through oracle/ref_harness.run_reference(dump_frames=True); stored gzip-compressed.

    python tests/golden/make_golden_ticks.py
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


class _Enc(json.JSONEncoder):
    def default(self, o):
        if isinstance(o, np.generic):
            return o.item()
        if isinstance(o, np.ndarray):
            return o.tolist()
        return super().default(o)


def main():
    cfg = rh.load_config(os.path.join(HERE, "configs", "north_pit_mine.json"))
    r = rh.run_reference(cfg, "SQDispatcher", "det", seed=0, sim_time=120, dump_frames=True)
    with gzip.open(os.path.join(HERE, "ticks_north_pit_mine_SQDispatcher_det_t120.json.gz"), "wt") as fh:
        json.dump(r.ticks, fh, cls=_Enc)
    z = np.load(os.path.join(HERE, "live_north_pit_mine_SPTFDispatcher_s11.npz"))
    r = rh.run_reference(cfg, "SPTFDispatcher", "replay", tape=(z["tape_sites"], z["tape_values"]), dump_frames=True)
    with gzip.open(os.path.join(HERE, "ticks_north_pit_mine_SPTFDispatcher_live_s11.json.gz"), "wt") as fh:
        json.dump(r.ticks, fh, cls=_Enc)
    print("ok", r.ticks["summary"]["produced_tons"])


if __name__ == "__main__":
    main()
