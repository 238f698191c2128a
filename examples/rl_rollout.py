#!/usr/bin/env python
"""A PPO-style rollout collector on the batched env (the role gym.vector.AsyncVectorEnv plays in the reference's
trainers, test/cleanrl/ppo_single_net.py:593-760): N envs step together, observations / rewards / dones stay on the
device, every env restarts by itself when its episode ends, and one rollout iteration (policy forward, action sampling,
om_step) is a single CUDA-graph replay.

    python examples/rl_rollout.py [--envs 4096] [--steps 256]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import openmines_b200 as om                              # noqa: E402
from openmines_b200.rl_env import BatchedMineEnv         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=256)
    a = ap.parse_args()
    cfg = om.load_config(os.path.join(ROOT, "openmines_b200", "conf", "north_pit_mine.json"))
    env = BatchedMineEnv(cfg, a.envs, rng_mode="philox", auto_reset=True)
    dev = env.mine.device
    policy = torch.nn.Sequential(torch.nn.Linear(env.obs_dim, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                                 torch.nn.Linear(256, env.n_actions)).to(dev)
    replay = env.capture_rollout_step(policy)
    env.reset(seed=0)
    obs_buf = torch.zeros((a.steps, a.envs, env.obs_dim), device=dev)
    act_buf = torch.zeros((a.steps, a.envs), dtype=torch.int32, device=dev)
    rew_buf = torch.zeros((a.steps, a.envs), dtype=torch.float64, device=dev)
    done_buf = torch.zeros((a.steps, a.envs), dtype=torch.bool, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(a.steps):
        obs_buf[t].copy_(env.io["obs"])          # the observation the policy is about to act on
        replay()                                 # policy(obs) -> sample -> om_step, one graph launch
        act_buf[t].copy_(env.last_actions)
        rew_buf[t].copy_(env.io["reward"])
        done_buf[t].copy_(env.io["done"].bool())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    teacher_agree = float((act_buf[-1] == env.io["sug_action"]).float().mean())
    print("%d envs x %d steps in %.3f s = %.2e env-steps/s; mean reward %.3f; episodes finished %d; "
          "last-step agreement with the ShortestTrip teacher %.2f" % (
              a.envs, a.steps, dt, a.envs * a.steps / dt, float(rew_buf.mean()), int(env.io["episode_totals"][:, 0].sum()),
              teacher_agree))


if __name__ == "__main__":
    main()
