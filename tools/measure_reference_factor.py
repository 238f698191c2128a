"""How much faster is the CPU port (oracle/des_oracle.py, what bench.py's reference arm times on the GPU box) than the
UNMODIFIED reference (what that arm stands in for)?  Authoring container only: needs /root/reference.

Both run the same workload (BASELINE configs[2]: north_pit_mine, SQ and SPTF, live RNG, same seeds), simulation only
(``dump_frames`` skipped on the reference side, no trace recording on the port side), one process, best of ``--repeat``
per episode to take scheduling noise out.  Prints the per-dispatcher seconds per episode and the factor
``port_speedup = t_reference / t_port``: the reference arm's throughput divided by this factor is an estimate of the
unmodified reference's throughput on the same cores.  Result recorded in BASELINE.md section 5 and bench.py
(REFERENCE_FACTOR)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=6)
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    from oracle import des_oracle as do
    from oracle import ref_harness as rh
    with open(os.path.join(ROOT, "openmines_b200", "conf", "north_pit_mine.json"), encoding="utf-8") as fh:
        cfg = json.load(fh)
    rows = {}
    for d in ("SQDispatcher", "SPTFDispatcher"):
        t_ref = t_port = 0.0
        ev = 0
        for seed in range(a.seeds):
            best_r = best_p = 1e9
            for _ in range(a.repeat):
                t0 = time.perf_counter()
                r = rh.run_reference(cfg, d, "live", seed=seed)
                best_r = min(best_r, time.perf_counter() - t0)
                t0 = time.perf_counter()
                o = do.DesOracle(cfg, d, "live", seed, record_trace=False).run()
                best_p = min(best_p, time.perf_counter() - t0)
            assert r.reported_tons == o.reported_tons and r.n_truck_events == o.n_truck_events      # same trajectories
            t_ref += best_r
            t_port += best_p
            ev += o.n_truck_events
        rows[d] = (t_ref / a.seeds, t_port / a.seeds, ev / a.seeds)
        print("%-16s reference %.3f s/episode  port %.3f s/episode  factor %.2f  (%.0f truck events/episode)" % (
            d, rows[d][0], rows[d][1], rows[d][0] / rows[d][1], rows[d][2]))
    tr = sum(v[0] for v in rows.values())
    tp = sum(v[1] for v in rows.values())
    print("workload (SQ + SPTF): reference %.3f s  port %.3f s  port_speedup %.2f" % (tr / 2, tp / 2, tr / tp))


if __name__ == "__main__":
    main()
