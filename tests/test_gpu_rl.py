"""GPU parity of the batched RL step (om_step through the C ABI) against the reference goldens and
the RL oracle: order kinds, asking truck, teacher suggestion bit-exact; observation float32[194] within
1 float32 ulp (device fp64 log vs libm before the float32 cast); reward within 1e-9 relative (the
reference sums truck intervals in truck order, the engine integrates them incrementally)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def _drive(env, policies, max_steps=2000):
    """Step the batched env with one host policy per env; returns per-env logs."""
    import torch
    n = env.num_envs
    logs = [[] for _ in range(n)]
    obs, info = env.reset(seed=np.array([p["seed"] for p in policies], dtype=np.uint64),
                          tapes=[p.get("tape") for p in policies] if env.rng_mode == "tape" else None)
    done_prev = np.zeros(n, dtype=bool)
    reward = np.zeros(n)
    done = np.zeros(n, dtype=bool)
    for _ in range(max_steps):
        o = obs.cpu().numpy()
        kind = info["order_kind"].cpu().numpy()
        sug = info["sug_action"].cpu().numpy()
        truck = info["truck"].cpu().numpy()
        tons = info["produce_tons"].cpu().numpy()
        acts = np.zeros(n, dtype=np.int32)
        for e in range(n):
            if done_prev[e]:
                continue
            rec = {"obs": o[e].copy(), "reward": float(reward[e]), "done": bool(done[e]), "kind": int(kind[e]),
                   "sug_action": int(sug[e]), "truck": int(truck[e]), "produce_tons": float(tons[e])}
            if not done[e]:
                nch = env.mine.pc.D if kind[e] == 1 else env.mine.pc.L
                acts[e] = policies[e]["fn"](len(logs[e]), int(kind[e]), o[e], nch)
                rec["action"] = int(acts[e])
            logs[e].append(rec)
        done_prev = done_prev | done
        if done_prev.all():
            break
        obs, r, term, _, info = env.step(torch.from_numpy(acts).to(env.mine.device))
        reward = r.cpu().numpy()
        done = term.cpu().numpy()
    return logs


def _compare(log, ref_obs, ref_reward, ref_kind, ref_sug, ref_done, ref_action):
    assert len(log) == len(ref_reward), (len(log), len(ref_reward))
    for k, rec in enumerate(log):
        assert rec["kind"] == ref_kind[k] and rec["done"] == bool(ref_done[k]), k
        assert rec["sug_action"] == ref_sug[k], (k, rec["sug_action"], ref_sug[k])
        if not rec["done"]:
            assert rec["action"] == ref_action[k], k
            a, b = rec["obs"], ref_obs[k]
            tol = np.maximum(np.abs(b), 1e-30) * 2.4e-7
            bad = np.nonzero(np.abs(a - b) > tol)[0]
            assert len(bad) == 0, (k, bad[:8], a[bad[:8]], b[bad[:8]])
        assert abs(rec["reward"] - ref_reward[k]) <= 1e-9 * max(1.0, abs(ref_reward[k])), (k, rec["reward"], ref_reward[k])


@pytest.mark.parametrize("split", ["0", "1"])
def test_rl_step_vs_reference_goldens(north_pit, monkeypatch, split):
    """det + tape replays of the reference's own RL loop (ShortestTrip and SQ teachers; dense and sparse reward modes), on
    the unsplit layout and on a forced split layout (the big-mine step kernels: hot truck arrays, uint16 breakdown minutes,
    pending-repair flag, register-carried chain values)."""
    import openmines_b200 as om
    monkeypatch.setenv("OM_SPLIT", split)
    from openmines_b200.rl_env import BatchedMineEnv
    from oracle import rl_oracle as ro
    names = ["rl_north_pit_mine_det_s0_ShortestTripDispatcher_p1", "rl_north_pit_mine_live_s21_ShortestTripDispatcher_p2",
             "rl_north_pit_mine_live_s22_SQDispatcher_p3", "rl_north_pit_mine_live_s23_ShortestTripDispatcher_p4_sparse"]
    for name in names:
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        mode = "det" if str(z["mode"]) == "det" else "tape"
        env = BatchedMineEnv(north_pit, 1, rng_mode=mode, sug_dispatcher=str(z["teacher"]),
                             reward_mode="sparse" if name.endswith("_sparse") else "dense")
        assert env.obs_dim == 194 and bool(env.mine.layout.t_aux) == (split == "1")
        pol = {"fn": ro.keyed_policy(int(z["policy_seed"])), "seed": int(z["seed"]),
               "tape": (z["tape_sites"], z["tape_values"])}
        logs = _drive(env, [pol])
        _compare(logs[0], z["obs"], z["reward"], z["kind"], z["sug_action"], z["done"], z["action"])
        env.close()


def test_rl_step_batch_vs_oracle(north_pit):
    """A batch of envs with different policies / teachers in one om_step call vs the RL oracle."""
    from openmines_b200.rl_env import BatchedMineEnv
    from oracle import rl_oracle as ro
    for teacher in ("SPTFDispatcher", "NearestDispatcher"):
        pols = [{"fn": ro.keyed_policy(40 + k), "seed": k} for k in range(4)]
        env = BatchedMineEnv(north_pit, len(pols), rng_mode="det", sug_dispatcher=teacher)
        logs = _drive(env, pols)
        for k, p in enumerate(pols):
            r = ro.run_rl_oracle(north_pit, p["fn"], "det", seed=k, teacher=teacher)
            ol = r.rl_log
            _compare(logs[k], [x["obs"] for x in ol], [x["reward"] for x in ol], [x["kind"] for x in ol],
                     [x["sug_action"] for x in ol], [x["done"] for x in ol], [x["action"] for x in ol])
        env.close()


def test_rl_live_runs_and_finishes(north_pit):
    """Philox mode smoke: 256 envs with a uniform random policy all reach done with sane returns."""
    import torch
    from openmines_b200.rl_env import BatchedMineEnv
    env = BatchedMineEnv(north_pit, 256, rng_mode="philox")
    obs, info = env.reset(seed=123)
    g = torch.Generator(device="cpu").manual_seed(0)
    total = torch.zeros(256, dtype=torch.float64, device=obs.device)
    steps = 0
    while True:
        a = torch.randint(0, 5, (256,), generator=g)
        obs, r, term, trunc, info = env.step(a.to(obs.device))
        total += r
        steps += 1
        assert torch.isfinite(obs).all()
        if bool(term.all()) or steps > 2000:
            break
    assert bool(term.all()) and 300 < steps < 1500
    tons = info["produce_tons"].cpu().numpy()
    assert 5000 < tons.mean() < 11000            # random dispatching: reference mean 8178 +- 421
    env.close()


def test_shortest_trip_as_builtin_dispatcher(north_pit):
    """ShortestTripDispatcher as a 7th device dispatcher in om_run, det mode, vs the oracle."""
    import openmines_b200 as om
    from helpers import engine_vs_ref
    from oracle import des_oracle as do
    o = do.DesOracle(north_pit, "SQDispatcher", "det")
    o.dispatcher = "ShortestTripDispatcher"
    ref = o.run()
    mine = om.BatchedMine(north_pit, 1)
    mine.reset("ShortestTripDispatcher", rng_mode="det")
    res = mine.run(trace_capacity=8192, site_ticks=True, order_capacity=2048)
    diffs = engine_vs_ref(ref, mine, res, 0)
    assert not diffs, diffs[:5]


def test_graphed_rollout_matches_eager(north_pit):
    """The CUDA-graph replay of one rollout iteration (policy + om_step) gives the same trajectories as eager stepping."""
    import torch
    from openmines_b200.rl_env import BatchedMineEnv
    n = 64
    lin = torch.nn.Linear(194, 5).cuda()
    policy = lambda o: lin(o)
    argmax = lambda logits: logits.argmax(dim=1)
    outs = []
    for graphed in (False, True):
        env = BatchedMineEnv(north_pit, n, rng_mode="philox")
        replay = env.capture_rollout_step(policy, sample=argmax) if graphed else None
        obs, info = env.reset(seed=5)
        rewards = []
        with torch.no_grad():
            for _ in range(40):
                if graphed:
                    replay()
                else:
                    env.step(argmax(policy(obs)))
                rewards.append(env.io["reward"].clone())
        outs.append((torch.stack(rewards).cpu(), env.io["obs"].clone().cpu(), env.io["info"].clone().cpu()))
        env.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_per_env_auto_reset(north_pit):
    """auto_reset=True: on the step an episode ends the env reports the final reward / terminated, restarts itself in the
    same om_step call with seed + num_envs and returns the next episode's first observation; the second episode must be
    identical to a fresh env reset with that seed, and the terminal observation lands in info['final_observation']."""
    import torch
    from openmines_b200.rl_env import BatchedMineEnv
    n = 8
    cfg = dict(north_pit)
    cfg["sim_time"] = 40
    lin = torch.nn.Linear(194, 5).cuda()
    act_of = lambda o: lin(o).argmax(dim=1)

    def rollout(env, seed, steps):
        obs, info = env.reset(seed=seed)
        log = []
        with torch.no_grad():
            for _ in range(steps):
                prev = obs.clone()
                obs, rew, term, trunc, info = env.step(act_of(obs))
                log.append((prev.cpu(), rew.clone().cpu(), term.clone().cpu(), obs.clone().cpu(),
                            info["order_kind"].clone().cpu(), info["sug_action"].clone().cpu()))
        return log

    auto = BatchedMineEnv(cfg, n, rng_mode="philox", auto_reset=True)
    a = rollout(auto, 100, 260)
    totals = auto.io["episode_totals"].cpu().numpy()
    final_obs = auto.io["final_obs"].cpu()
    assert (totals[:, 0] >= 1).all() and (totals[:, 1] > 0).all() and (totals[:, 2] > 0).all()
    # env e: steps of its first episode, then its second episode must replay a fresh env seeded 100 + n + e
    plain = BatchedMineEnv(cfg, n, rng_mode="philox")
    b = rollout(plain, 100, 260)            # first episodes (no auto reset)
    c = rollout(plain, 100 + n, 260)        # what the second episodes must look like
    for e in range(n):
        t_end = next(k for k, rec in enumerate(a) if bool(rec[2][e]))
        for k in range(t_end + 1):          # first episode identical incl. the final reward
            assert torch.equal(a[k][1][e], b[k][1][e]) and bool(a[k][2][e]) == bool(b[k][2][e])
        for k in range(t_end):
            assert torch.equal(a[k][3][e], b[k][3][e])
        assert torch.equal(a[t_end][3][e], c[0][0][e])                 # obs returned with done = first obs of episode 2
        assert int(a[t_end][4][e]) == 0                                 # pending decision of the new episode: an init order
        if totals[e, 0] == 1:
            assert torch.equal(final_obs[e], b[t_end][3][e])            # terminal observation of episode 1
        m = min(len(a) - t_end - 1, 40)
        for k in range(m):                  # second episode == fresh env with the next seed
            assert torch.equal(a[t_end + 1 + k][1][e], c[k][1][e]), (e, k)
            assert torch.equal(a[t_end + 1 + k][3][e], c[k][3][e]), (e, k)


@pytest.mark.parametrize("which", ["ragged", "scaled"])
def test_rl_env_on_other_mines(which, ragged_mine):
    """The batched env on mines with L != D (generalised observation layout) and on the 500-truck mine (SPLIT step
    kernel, slot cache), with per-env auto reset: finite observations, valid pending decisions, teacher suggestions in
    range, episodes finish and restart, no engine errors."""
    import torch
    from openmines_b200.rl_env import BatchedMineEnv
    if which == "ragged":
        cfg, n, jam = dict(ragged_mine), 64, 64
        cfg["sim_time"] = 30
    else:
        from openmines_b200.synthetic import scaled_mine
        cfg, n, jam = scaled_mine(sim_time=12), 16, 512
    env = BatchedMineEnv(cfg, n, rng_mode="philox", auto_reset=True, jam_capacity=jam)
    L, D = env.mine.pc.L, env.mine.pc.D
    assert env.obs_dim == 6 + (1 + L + D) + 2 + 3 * max(L, D) + 2 * env.mine.pc.R + 5 * (L + D)
    obs, info = env.reset(seed=9)
    g = torch.Generator(device="cpu").manual_seed(0)
    steps = 1500 if which == "ragged" else 2500
    for _ in range(steps):
        kind = info["order_kind"]
        nmax = torch.where(kind == 1, torch.tensor(D, device=kind.device), torch.tensor(L, device=kind.device))
        a = (torch.rand(n, generator=g).to(kind.device) * nmax).long().clamp_(max=max(L, D) - 1)
        obs, rew, term, trunc, info = env.step(a)
        sug = info["sug_action"]
        assert bool(((sug >= 0) & (sug < nmax_next(info, L, D))).all())
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    hdr = env.mine.headers()
    assert (hdr["err"] == 0).all()
    totals = env.io["episode_totals"].cpu().numpy()
    assert (totals[:, 0] >= 1).all(), totals[:, 0]


def nmax_next(info, L, D):
    import torch
    kind = info["order_kind"]
    return torch.where(kind == 1, torch.tensor(D, device=kind.device), torch.tensor(L, device=kind.device))


def test_c4_full_size_teacher_rollout_equals_episode_kernel(north_pit):
    """BASELINE configs[3] at its full size (16384 envs, stochastic events on).  Size-independent property: a rollout
    that always plays the teacher's suggestion (ShortestTrip, computed inside om_step) must reproduce, env by env and bit
    for bit, the episode kernel running ShortestTripDispatcher as a built-in device dispatcher on the same seeds -- the
    pause / resume of the step path, the RL-surface state layout and the per-step staging of the blobs change nothing in
    the simulated trajectory."""
    import torch
    import openmines_b200 as om
    from openmines_b200.rl_env import BatchedMineEnv
    n = 16384
    env = BatchedMineEnv(north_pit, n, rng_mode="philox")
    assert env.single_action_space.n == 5 and env.single_observation_space.shape == (194,)      # mine_env.py:217-226
    assert np.dtype(env.single_observation_space.dtype) == np.float32
    obs, info = env.reset(seed=0)
    steps, returns = 0, torch.zeros(n, dtype=torch.float64, device=env.mine.device)
    while True:
        a = info["sug_action"].clamp(min=0)                 # finished envs report -1 and ignore the action
        obs, r, term, trunc, info = env.step(a)
        returns += r
        steps += 1
        if steps % 64 == 0 and bool(term.all()):
            break
        assert steps < 4096, "episodes do not finish"
    assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(returns).all())
    h = env.mine.headers()
    assert (h["err"] == 0).all() and (h["finished"] == 1).all()

    mine = om.BatchedMine(north_pit, n)
    mine.reset(["ShortestTripDispatcher"] * n, seeds=np.arange(n, dtype=np.int64), rng_mode="philox")
    s = mine.run().summary_np()
    assert (s["err"] == 0).all()
    assert (h["reported_tons"] == s["reported_tons"]).all()
    assert (h["n_events"] == s["n_truck_events"]).all()
    assert (h["n_decisions"] == s["n_decisions"]).all()
    assert (h["n_jams"] == s["n_jams"]).all() and (h["n_brk"] == s["n_truck_repairs"]).all()
    assert (h["rl_steps"] == s["n_decisions"] + 1).all()    # one env step per dispatch decision + the terminal step
    assert len(np.unique(s["reported_tons"])) > n // 2      # the seeds really differ
