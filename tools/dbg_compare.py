"""Debugging aid: run one det case under the plain and the bounds-checking build (two subprocesses) and show where their
event traces / order logs part.  usage: python tools/dbg_compare.py [dispatcher] [until]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(disp, until, out):
    import openmines_b200 as om
    from conftest import load_cfg
    cfg = load_cfg("north_pit_mine")
    mine = om.BatchedMine(cfg, 1)
    mine.reset(disp, rng_mode="det")
    res = mine.run(until=until, trace_capacity=4096, order_capacity=2048, site_ticks=True)
    tr, s = res.trace_np(0)
    np.savez(out, trace=tr[:int(s["n_truck_events"]) + 50], orders=res.orders_np()[0], summary=np.array([s]))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sanitize_target
    sanitize_target.bounds_report()


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "--child":
        child(sys.argv[2], float(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    disp = sys.argv[1] if len(sys.argv) > 1 else "SPTFDispatcher"
    until = sys.argv[2] if len(sys.argv) > 2 else "45"
    outs = []
    second = os.environ.get("DBG_LIB")          # explicit path of the library to compare with the plain build
    for dbg in ("", "1"):
        env = dict(os.environ)
        env.pop("OM_DEBUG_LIB", None)
        env.pop("OM_LIB_PATH", None)
        if dbg:
            if second:
                env["OM_LIB_PATH"] = second
            else:
                env["OM_DEBUG_LIB"] = "1"
        out = "/tmp/dbgcmp_%s.npz" % (dbg or "0")
        subprocess.run([sys.executable, __file__, "--child", disp, until, out], env=env, check=True)
        outs.append(np.load(out))
    a, b = outs
    oa, ob = a["orders"], b["orders"]
    ne = np.nonzero(oa != ob)[0]
    print("orders: first difference at decision", ne[0] if len(ne) else None)
    if len(ne):
        k = int(ne[0])
        print(" plain", [(int(x) >> 12, int(x) & 0xFFF) for x in oa[max(0, k - 3):k + 3]])
        print(" debug", [(int(x) >> 12, int(x) & 0xFFF) for x in ob[max(0, k - 3):k + 3]])
    ta, tb = a["trace"], b["trace"]
    n = min(len(ta), len(tb))
    ne = np.nonzero(ta[:n] != tb[:n])[0]
    print("trace: first difference at record", ne[0] if len(ne) else None)
    if len(ne):
        k = int(ne[0])
        for r in range(max(0, k - 4), min(n, k + 4)):
            print(" ", r, "plain", ta[r], "| debug", tb[r])
    print("summaries", a["summary"], b["summary"])
    na, nb = int(a["summary"]["n_truck_events"][0]), int(b["summary"]["n_truck_events"][0])
    print("kind histogram plain", np.bincount(ta["kind"][:len(ta) - 50], minlength=16))
    print("kind histogram debug", np.bincount(tb["kind"][:len(tb) - 50], minlength=16))
    # records of plain that are missing in debug (multiset difference on (time, truck, kind, arg0))
    from collections import Counter
    ka = Counter((float(r["time"]), int(r["truck"]), int(r["kind"]), int(r["arg0"])) for r in ta[:len(ta) - 50])
    kb = Counter((float(r["time"]), int(r["truck"]), int(r["kind"]), int(r["arg0"])) for r in tb[:len(tb) - 50])
    miss = sorted((ka - kb).elements())
    extra = sorted((kb - ka).elements())
    print("missing in debug (%d):" % len(miss), miss[:12])
    print("extra in debug (%d):" % len(extra), extra[:12])
