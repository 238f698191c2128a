"""Result-JSON golden for the ragged mine (see make_golden_edge.py) from the UNMODIFIED reference's TickGenerator:
a tape replay of the live SQ run, i.e. with jams, frequent road repairs, truck breakdowns, 1..9 units per site.

    python tests/golden/make_golden_ticks_edge.py
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from make_golden_ticks import _Enc            # noqa: E402


def main():
    cfg = rh.load_config(os.path.join(HERE, "configs", "ragged_mine.json"))
    z = np.load(os.path.join(HERE, "live_ragged_mine_SQDispatcher_s5.npz"))
    r = rh.run_reference(cfg, "SQDispatcher", "replay", tape=(z["tape_sites"], z["tape_values"]), dump_frames=True)
    with gzip.open(os.path.join(HERE, "ticks_ragged_mine_SQDispatcher_live_s5.json.gz"), "wt") as fh:
        json.dump(r.ticks, fh, cls=_Enc)
    print("ok", r.ticks["summary"]["produced_tons"], r.ticks["summary"]["RoadJams"])


if __name__ == "__main__":
    main()
