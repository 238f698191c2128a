"""Workload of tools/sanitize.sh: short runs of every kernel family under compute-sanitizer (memcheck / racecheck /
initcheck).  Each leg checks its result against a golden or the oracle as well, so a sanitizer-clean run is also a correct run.

legs: north_pit det + tape with trace / order log / site + unit ticks, Philox with the draw-tape export; the ragged mine
(L != D); the forced SPLIT layout; a mixed-config batch (MULTI kernels); om_step with per-env auto reset; om_stats_reduce."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def bounds_report():
    import ctypes
    from openmines_b200 import _lib
    rep = (ctypes.c_int32 * 8)()
    checks = _lib.load().om_debug_bounds_report(rep)
    print("bounds-check build: %s; index violations %d (first: om_engine.cuh line %d, index %d, extent %d); "
          "convergence violations %d (first: line %d, active mask %08x)" % (
              {1: "yes", 2: "convergence probes only"}.get(checks, "NO (plain build)"), rep[0], rep[1], rep[2], rep[3], rep[4], rep[5], rep[6] & 0xFFFFFFFF), flush=True)
    return checks, rep[0] + rep[4]


def main():
    try:
        legs()
    finally:
        checks, bad = bounds_report()
    if checks in (1, 2) and bad != 0:
        sys.exit(9)


def legs():
    import torch
    import openmines_b200 as om
    from conftest import load_cfg
    from helpers import DISPATCHERS, Golden, engine_vs_ref
    from oracle import des_oracle as do
    legs = sys.argv[1:] or ["det", "tape", "philox", "ragged", "split", "multi", "step", "stats"]
    np_cfg, rag = load_cfg("north_pit_mine"), load_cfg("ragged_mine")
    until = 45.0
    if "det" in legs:
        mine = om.BatchedMine(np_cfg, 6)
        mine.reset(DISPATCHERS, rng_mode="det")
        res = mine.run(until=until, trace_capacity=2048, site_ticks=True, unit_ticks=True, order_capacity=1024)
        for i, d in enumerate(DISPATCHERS):
            o = do.run_oracle(np_cfg, d, "det", seed=i, sim_time=until)
            diffs = engine_vs_ref(o, mine, res, i, sptf=(d == "SPTFDispatcher"))
            assert not diffs, (d, diffs[:4])
        print("det ok", flush=True)
    if "tape" in legs:
        gs = [Golden("live_north_pit_mine_%s_s11" % d) for d in ("SQDispatcher", "SPTFDispatcher")]
        mine = om.BatchedMine(np_cfg, 2)
        mine.reset(["SQDispatcher", "SPTFDispatcher"], rng_mode="tape", tapes=[g.tape for g in gs])
        res = mine.run(until=until, trace_capacity=2048, site_ticks=True, order_capacity=1024)
        s = res.summary_np()
        for i, g in enumerate(gs):
            tr = g.trace[g.trace["time"] < until]
            got = res.trace_np(i)[0][:len(tr)]
            assert np.array_equal(got["kind"], tr["kind"]) and np.array_equal(got["time"], tr["time"])
        assert (s["err"] == 0).all()
        print("tape ok", flush=True)
    if "philox" in legs:
        mine = om.BatchedMine(np_cfg, 4)
        names = ["SQDispatcher", "SPTFDispatcher", "RandomDispatcher", "FixedGroupDispatcher"]
        mine.reset(names, rng_mode="philox", seeds=[5, 6, 7, 8])
        res = mine.run(until=until, trace_capacity=2048, site_ticks=True, order_capacity=1024, draw_capacity=mine.draw_tape_slots(until))
        for i, d in enumerate(names):
            o = do.run_oracle(np_cfg, d, "replay", tape=res.draw_tape(i), sim_time=until)
            assert not engine_vs_ref(o, mine, res, i, sptf=(d == "SPTFDispatcher")), d
        print("philox + draw tape ok", flush=True)
    if "ragged" in legs:
        mine = om.BatchedMine(rag, 3, jam_capacity=128)
        names = ["SQDispatcher", "SPTFDispatcher", "NearestDispatcher"]
        mine.reset(names, rng_mode="philox", seeds=[1, 2, 3])
        res = mine.run(until=until, trace_capacity=4096, site_ticks=True, order_capacity=2048, draw_capacity=mine.draw_tape_slots(until))
        for i, d in enumerate(names):
            o = do.run_oracle(rag, d, "replay", tape=res.draw_tape(i), sim_time=until)
            assert not engine_vs_ref(o, mine, res, i, sptf=(d == "SPTFDispatcher")), d
        print("ragged ok", flush=True)
    if "split" in legs:
        os.environ["OM_SPLIT"] = "1"
        mine = om.BatchedMine(np_cfg, 3)
        names = ["SQDispatcher", "SPTFDispatcher", "NaiveDispatcher"]
        mine.reset(names, rng_mode="philox", seeds=[11, 12, 13])
        res = mine.run(until=until, trace_capacity=2048, site_ticks=True, order_capacity=1024, draw_capacity=mine.draw_tape_slots(until))
        for i, d in enumerate(names):
            o = do.run_oracle(np_cfg, d, "replay", tape=res.draw_tape(i), sim_time=until)
            assert not engine_vs_ref(o, mine, res, i, sptf=(d == "SPTFDispatcher")), d
        del os.environ["OM_SPLIT"]
        print("split ok", flush=True)
    if "multi" in legs:
        short = load_cfg("north_pit_mine_short")
        cfgs = [np_cfg, rag, short]
        cids = [0, 1, 2, 1, 0, 2]
        names = ["SQDispatcher", "SPTFDispatcher", "NaiveDispatcher", "SQDispatcher", "SPTFDispatcher", "NearestDispatcher"]
        mine = om.BatchedMine(cfgs, 6, config_ids=cids, jam_capacity=[64, 128, 64])
        mine.reset(names, rng_mode="philox", seeds=np.arange(6) + 21)
        res = mine.run(until=until, trace_capacity=4096, site_ticks=True, order_capacity=2048, draw_capacity=mine.draw_tape_slots(until))
        for i, d in enumerate(names):
            o = do.run_oracle(cfgs[cids[i]], d, "replay", tape=res.draw_tape(i), sim_time=until)
            assert not engine_vs_ref(o, mine, res, i, sptf=(d == "SPTFDispatcher")), (i, d)
        print("multi ok", flush=True)
    if "step" in legs:
        from openmines_b200.rl_env import BatchedMineEnv
        env = BatchedMineEnv(dict(np_cfg, sim_time=12), 8, rng_mode="philox", auto_reset=True)
        obs, info = env.reset(seed=3)
        ends = 0
        for k in range(400):
            a = info["sug_action"].clamp(min=0)
            obs, rew, term, trunc, info = env.step(a)
            ends += int(term.sum().item())
        assert ends >= 8 and torch.isfinite(obs).all()
        print("step ok (%d episode ends)" % ends, flush=True)
    if "stats" in legs:
        mine = om.BatchedMine(np_cfg, 300)
        ids = (np.arange(300) % 7).astype(np.int32)
        mine.reset(ids, rng_mode="philox")
        res = mine.run(until=20.0)
        from openmines_b200 import parallel
        assert np.array_equal(mine.stats(res).cpu().numpy(), parallel.stats_from_summaries(res.summary_np(), ids))
        print("stats ok", flush=True)


if __name__ == "__main__":
    main()
