"""Golden vectors of the reference's RL step surface (RLDispatcher + Mine.start_rl + preprocess_observation),
recorded from the UNMODIFIED reference at /root/reference through oracle/ref_harness.run_reference_rl.

    python tests/golden/make_golden_rl.py

The agent is the deterministic keyed policy oracle/rl_oracle.keyed_policy(policy_seed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh, rl_oracle as ro  # noqa: E402


def save(path, log, tape, meta):
    d = {"obs": np.stack([x["obs"] for x in log]).astype(np.float32),
         "reward": np.array([x["reward"] for x in log], dtype=np.float64),
         "done": np.array([x["done"] for x in log], dtype=np.bool_),
         "kind": np.array([x["kind"] for x in log], dtype=np.int32),
         "sug_action": np.array([x["sug_action"] for x in log], dtype=np.int32),
         "action": np.array([x["action"] for x in log], dtype=np.int32),
         "time": np.array([x["time"] for x in log], dtype=np.float64),
         "produce_tons": np.array([x["produce_tons"] for x in log], dtype=np.float64),
         "tape_sites": tape[0], "tape_values": tape[1]}
    d.update({k: np.array(v) for k, v in meta.items()})
    np.savez_compressed(path, **d)


def main():
    cfg = rh.load_config(os.path.join(HERE, "configs", "north_pit_mine.json"))
    cases = [("det", 0, "ShortestTripDispatcher", 1), ("live", 21, "ShortestTripDispatcher", 2),
             ("live", 22, "SQDispatcher", 3)]
    for mode, seed, teacher, pseed in cases:
        log, tape = rh.run_reference_rl(cfg, ro.keyed_policy(pseed), mode, seed, teacher=teacher)
        name = "rl_north_pit_mine_%s_s%d_%s_p%d.npz" % (mode, seed, teacher, pseed)
        save(os.path.join(HERE, name), log, tape, {"mode": mode, "seed": seed, "teacher": teacher, "policy_seed": pseed})
        print(name, len(log), sum(x["reward"] for x in log))
    # sparse reward mode (rl_dispatch.py:186-192, mine.py:456-459)
    log, tape = rh.run_reference_rl(cfg, ro.keyed_policy(4), "live", 23, teacher="ShortestTripDispatcher", reward_mode="sparse")
    name = "rl_north_pit_mine_live_s23_ShortestTripDispatcher_p4_sparse.npz"
    save(os.path.join(HERE, name), log, tape, {"mode": "live", "seed": 23, "teacher": "ShortestTripDispatcher", "policy_seed": 4})
    print(name, len(log), sum(x["reward"] for x in log), log[-1]["reward"])


if __name__ == "__main__":
    main()
