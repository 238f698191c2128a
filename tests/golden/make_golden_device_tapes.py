"""Philox-mode device trajectories replayed by the UNMODIFIED reference (authoring container only).

Input: gpurun_out/device_tapes.npz, written on a B200 by tools/export_device_tapes.py -- per instance the draw tape the
Philox kernel exported (OmOutputs.draw_values / draw_sites) and what the device produced (event trace, production and
waiting-truck curves, summary, order log).  Each tape is fed to the reference's own run_dispatch_sim through
oracle/ref_harness.py (RngControl "replay": every np.random.* / random.* call of the reference returns the next tape
entry and checks its site id); the reference's event trace, curves, tonnage and order lists must equal the device's bit
for bit, and the tape must be consumed to its last entry.  What passes is committed as tests/golden/device_tapes.npz;
tests/test_oracle_golden.py::test_device_philox_trajectories_replayed_by_the_reference re-checks it through the reference
where /root/reference exists and through the CPU restatement everywhere.

    python tests/golden/make_golden_device_tapes.py
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import trace_format as tf  # noqa: E402


def cases(z):
    keys = sorted(k[:-len("_tape_values")] for k in z.files if k.endswith("_tape_values"))
    for key in keys:
        cname = "north_pit_mine" if key.startswith("north_pit_mine") else "ragged_mine"
        yield key, cname, key.split("_")[-1]


def compare(key, z, r, sptf):
    """device arrays of `key` in z vs a reference / oracle result object; returns a list of differences"""
    diffs = []
    dev = z[key + "_trace"].copy()
    ref = r.trace.copy()
    for t in (dev, ref):
        t["truck"][t["kind"] >= 14] = tf.NO_TRUCK
    if len(dev) != len(ref):
        diffs.append("trace length %d vs %d" % (len(dev), len(ref)))
    n = min(len(dev), len(ref))
    for f in dev.dtype.names:
        if f == "flags":
            continue
        bad = np.nonzero(dev[f][:n] != ref[f][:n])[0]
        if len(bad):
            diffs.append("trace.%s differs first at %d: device %r reference %r" % (f, bad[0], dev[bad[0]], ref[bad[0]]))
    nt = r.n_ticks
    if not np.array_equal(z[key + "_produced"][:nt], r.tick["produced_tons"]):
        diffs.append("production curve differs")
    if not np.array_equal(z[key + "_waiting"][:nt].astype(np.int64), r.tick["waiting"].astype(np.int64)):
        diffs.append("waiting-truck curve differs")
    s = z[key + "_summary"][0]
    if float(s["reported_tons"]) != r.reported_tons or float(s["dumped_tons"]) != r.dumped_tons:
        diffs.append("tons %r/%r vs %r/%r" % (s["reported_tons"], s["dumped_tons"], r.reported_tons, r.dumped_tons))
    if int(s["n_truck_events"]) != r.n_truck_events or int(s["n_wrapper_orders"]) != r.order_count:
        diffs.append("event / order counts differ")
    row = z[key + "_orders"]
    kinds, sites = row >> 12, row & 0xFFF
    got_init = list(sites[(kinds == 0) | (kinds == 2)]) if sptf else list(sites[kinds == 0])
    if got_init != list(r.orders["init"]) or list(sites[kinds == 1]) != list(r.orders["haul"]) or \
            list(sites[kinds == 2]) != list(r.orders["back"]):
        diffs.append("order lists differ")
    if not (np.array_equal(r.tape[0], z[key + "_tape_sites"]) and np.array_equal(r.tape[1], z[key + "_tape_values"])):
        diffs.append("the replay did not consume the tape entry for entry")
    return diffs


def main():
    from conftest import load_cfg
    from oracle import ref_harness as rh
    src = os.path.join(ROOT, "gpurun_out", "device_tapes.npz")
    z = np.load(src)
    n_ok = 0
    for key, cname, disp in cases(z):
        cfg = load_cfg(cname)
        r = rh.run_reference(cfg, disp, "replay", tape=(z[key + "_tape_sites"], z[key + "_tape_values"]))
        assert r.error is None, (key, r.error)
        diffs = compare(key, z, r, disp == "SPTFDispatcher")
        print(key, "events", r.n_truck_events, "tons %.3f" % r.reported_tons, "draws", len(r.tape[1]), "OK" if not diffs else diffs[:3])
        assert not diffs, key
        n_ok += 1
    shutil.copy(src, os.path.join(HERE, "device_tapes.npz"))
    print("%d device trajectories reproduced by the unmodified reference; fixture written" % n_ok)


if __name__ == "__main__":
    main()
