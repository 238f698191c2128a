#!/usr/bin/env python
"""The reference's `openmines run -f conf/north_pit_mine.json` (cli/run.py:115-273) on the batched engine.

Runs every built-in dispatcher named in the config on N seeds at once, writes the reference's result JSON for the first
seed of each dispatcher (results/MINE-{mine}-ALGO-{dispatcher}-TIME-{ts}.json, ticker.py:455-470) and prints the table
the reference draws (produced tons, matching factor, total wait time, road jams; charter.py:41-66) as mean +- std over
the seeds.

    python examples/run_north_pit.py [config.json] [--seeds 64] [--det]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import openmines_b200 as om                                    # noqa: E402
from openmines_b200.dispatcher import BUILTIN                  # noqa: E402
from openmines_b200.experiments import run_dispatch_sim        # noqa: E402
from openmines_b200.ticker import TickRecorder                 # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default=os.path.join(ROOT, "openmines_b200", "conf", "north_pit_mine.json"))
    ap.add_argument("--seeds", type=int, default=64)
    ap.add_argument("--det", action="store_true", help="stochastic events off")
    ap.add_argument("--out", default="results")
    a = ap.parse_args()
    cfg = om.load_config(a.config)
    # the shipped north_pit config names only PPODispatcher (checkpoint not included): fall back to the six built-ins
    names = [n for n in cfg["dispatcher"]["type"] if n in BUILTIN and n != "ShortestTripDispatcher"] or \
        [n for n in BUILTIN if n != "ShortestTripDispatcher"]
    print("%-22s %12s %10s %10s %8s" % ("dispatcher", "tons", "+-", "events", "jams"))
    for name in names:
        r = run_dispatch_sim(name, cfg, n_seeds=a.seeds, stochastic=not a.det)
        s = r["summary"]
        rec = TickRecorder(r["mine"], r["result"], name)
        path = rec.write_to_file(r["ticks"], result_dir=a.out, dispatcher_name=name)
        print("%-22s %12.1f %10.1f %10.1f %8.1f   -> %s" % (name, s["reported_tons"].mean(), s["reported_tons"].std(),
                                                           s["n_truck_events"].mean(), s["n_jams"].mean(), path))
        r["mine"].close()


if __name__ == "__main__":
    main()
