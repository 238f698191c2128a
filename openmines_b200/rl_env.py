"""Batched gym-style environment over the device engine (BASELINE.json configs[3]).

Mirrors the reference's ``mine/Mine-v1`` env (utils/gym/openmines_gym/envs/mine_env.py:192-353):
``reset() -> (obs, info)`` and ``step(action) -> (obs, reward, terminated, truncated, info)``, where one
step is one dispatch decision of whichever truck asks next -- but for N independent mines at once, with
observations, rewards and flags as **device tensors** (no queues, threads or processes).  The
observation is the reference's float32[194] vector (utils/feature_processing.py:3-75; generalised to
C = max(L, D) candidate slots when L != D), the reward its dense reward (rl_dispatch.py:194-308)
including the final-step bonus, ``info["sug_action"]`` the teacher's decision (default
ShortestTripDispatcher, mine_env.py:55,196).
"""
import numpy as np

from .engine import BatchedMine, DISPATCHER_IDS


class _Space:
    """Minimal stand-in for ``gymnasium.spaces.Discrete`` / ``Box`` (gymnasium is optional): what vector-env trainers read
    (``single_action_space.n``, ``single_observation_space.shape`` / ``.dtype`` / ``.low`` / ``.high``)."""

    def __init__(self, shape, dtype, n=None, low=None, high=None):
        self.shape, self.dtype, self.n, self.low, self.high = tuple(shape), np.dtype(dtype), n, low, high

    def __repr__(self):
        return "Discrete(%d)" % self.n if self.n is not None else "Box(%r, %r, %r, %s)" % (self.low, self.high, self.shape,
                                                                                            self.dtype)


def _spaces(n_actions, obs_dim):
    """(action space, observation space) of one env: Discrete(max(L, D)), Box(0, inf, (obs_dim,), float32)
    (mine_env.py:217-226); real gymnasium spaces when the package is importable."""
    try:
        from gymnasium import spaces
        return spaces.Discrete(n_actions), spaces.Box(low=0.0, high=np.inf, shape=(obs_dim,), dtype=np.float32)
    except Exception:
        return (_Space((), np.int64, n=int(n_actions)),
                _Space((obs_dim,), np.float32, low=0.0, high=float("inf")))


class BatchedMineEnv:
    def __init__(self, config, num_envs, device=0, rng_mode="philox", sug_dispatcher="ShortestTripDispatcher",
                 auto_reset=False, jam_capacity=52, reward_mode="dense", copy_outputs=False):
        """``jam_capacity``: stored road jams per env.  52 (a pool of 48 never overflowed in 49152 live north_pit episodes, DESIGN.md
        section 2) instead of the episode runs' 64: with the RL surface's bigger state that is what lets 19 instead of 17 envs share an SM's
        shared memory (12192 bytes per env).  An env whose pool overflows ends its episode with ``OM_IERR_JAM_OVERFLOW``.

        ``copy_outputs``: ``reset()`` / ``step()`` return the engine's OWN output buffers (``env.io[...]``), which the next
        ``step()`` overwrites in place -- zero-copy, what a fused rollout loop wants, but a trainer that stores the returned
        ``obs`` / ``reward`` / ``terminated`` / ``info`` tensors in a rollout buffer without ``.clone()`` would see every
        stored step turn into the latest one.  ``copy_outputs=True`` returns fresh tensors every call instead (one extra
        device copy of N x (obs_dim x 4 + 41) bytes per step)."""
        import torch
        self.copy_outputs = bool(copy_outputs)
        self.torch = torch
        self.mine = BatchedMine(config, num_envs, device=device, jam_capacity=jam_capacity, rl_surface=True)
        self.num_envs = int(num_envs)
        self.rng_mode = rng_mode
        self.teacher = DISPATCHER_IDS[sug_dispatcher] if sug_dispatcher else -1
        self.auto_reset = auto_reset
        if reward_mode not in ("dense", "sparse"):
            raise ValueError("Unknown reward mode: %s" % reward_mode)          # rl_dispatch.py:72
        self.reward_mode = reward_mode                     # GymMineEnv(reward_mode=...), mine_env.py:55
        self.obs_dim = self.mine.obs_dim()
        self.n_actions = max(self.mine.pc.L, self.mine.pc.D)          # spaces.Discrete(max(L, D)), mine_env.py:217-221
        self.single_action_space, self.single_observation_space = _spaces(self.n_actions, self.obs_dim)
        dev, n = self.mine.device, self.num_envs
        self.io = {
            "obs": torch.zeros((n, self.obs_dim), dtype=torch.float32, device=dev),
            "reward": torch.zeros(n, dtype=torch.float64, device=dev),
            "done": torch.zeros(n, dtype=torch.uint8, device=dev),
            "order_kind": torch.full((n,), -1, dtype=torch.int32, device=dev),
            "truck": torch.full((n,), -1, dtype=torch.int32, device=dev),
            "sug_action": torch.full((n,), -1, dtype=torch.int32, device=dev),
            "info": torch.zeros((n, 4), dtype=torch.float64, device=dev),
        }
        if auto_reset:
            # per-env automatic reset on the device (gym.vector semantics, see OmStepIO in include/openmines_b200.h)
            self.io["final_obs"] = torch.zeros((n, self.obs_dim), dtype=torch.float32, device=dev)
            self.io["episode_totals"] = torch.zeros((n, 3), dtype=torch.int64, device=dev)
        self._episode = 0
        self._seeds = None
        self._tapes = None

    def _out(self, t):
        return t.clone() if self.copy_outputs else t

    def _info(self):
        io, o = self.io, self._out
        info = o(io["info"])
        return {"produce_tons": info[:, 0], "time": info[:, 1], "delta_time": info[:, 2],
                "sug_action": o(io["sug_action"]), "order_kind": o(io["order_kind"]), "truck": o(io["truck"]),
                "load_num": self.mine.pc.L, "unload_num": self.mine.pc.D}

    def reset(self, seed=None, tapes=None):
        """Start N fresh episodes and advance each to its first decision.  ``seed``: int base or array of N."""
        n = self.num_envs
        if seed is None:
            seeds = np.arange(n, dtype=np.uint64) + np.uint64(self._episode * n)
        elif np.isscalar(seed):
            seeds = np.arange(n, dtype=np.uint64) + np.uint64(seed)
        else:
            seeds = np.asarray(seed, dtype=np.uint64)
        self._seeds, self._tapes = seeds, tapes
        self._episode += 1
        self.mine.reset("external", seeds=seeds, rng_mode=self.rng_mode, tapes=tapes)
        if self.auto_reset:
            self.io["episode_totals"].zero_()
        self._step(None)
        return self._out(self.io["obs"]), self._info()

    def _step(self, a):
        self.mine.step(a, self.io, teacher=self.teacher, auto_reset=self.auto_reset, seed_stride=self.num_envs,
                       sparse_reward=(self.reward_mode == "sparse"))

    def step(self, actions):
        """actions: int tensor [N] on the env's device (int64 or int32).

        With ``auto_reset=True`` every env restarts by itself, inside the same ``om_step`` call, on the step its
        episode ends (the behaviour of ``gym.vector.AsyncVectorEnv`` that the reference's trainers run under,
        test/cleanrl/ppo_single_net.py:593): that step returns the final reward and ``terminated=True`` together with
        the FIRST observation of the next episode; ``info["final_observation"]`` holds the terminal observation rows and
        ``info["episode_totals"]`` the per-env counts of finished episodes / their truck events / their decisions.
        The next episode's seed is the previous one + num_envs.  Without it, finished envs idle (reward 0,
        ``terminated=True``) until the caller resets the batch.

        ALIASING: unless the env was built with ``copy_outputs=True`` the returned tensors are views of ``env.io`` and are
        overwritten by the next ``step()``; ``.clone()`` what you keep."""
        torch = self.torch
        a = actions.to(device=self.mine.device, dtype=torch.int32).contiguous()
        self._step(a)
        terminated = self.io["done"].bool()
        truncated = torch.zeros_like(terminated)
        info = self._info()
        if self.auto_reset:
            info["final_observation"] = self._out(self.io["final_obs"])
            info["episode_totals"] = self._out(self.io["episode_totals"])
        return self._out(self.io["obs"]), self._out(self.io["reward"]), terminated, truncated, info

    def capture_rollout_step(self, policy, sample=None, warmup=3):
        """Capture ONE rollout iteration -- policy forward on the current observation, action sampling, ``om_step``
        -- into a CUDA graph and return ``replay()``.  The iteration is launch-bound (a small MLP, a sampler and one
        engine kernel per env step), so replaying a graph removes most of the host time between env steps.  The
        env's ``io`` tensors are the graph's static buffers: after ``replay()`` read ``env.io["obs"]``,
        ``env.io["reward"]``, ``env.io["done"]`` (``env.last_actions`` holds the sampled actions).  Capturing
        advances the envs by ``warmup + 1`` steps: call ``reset()`` afterwards."""
        torch = self.torch
        if self._seeds is None:
            self.reset()          # the warm-up / capture steps need initialised envs
        if sample is None:
            def sample(logits):
                # validate_args=False: argument validation synchronises with the host, which is illegal during capture
                return torch.distributions.Categorical(logits=logits, validate_args=False).sample()
        self.last_actions = torch.zeros(self.num_envs, dtype=torch.int32, device=self.mine.device)

        def body():
            act = sample(policy(self.io["obs"]))
            self.last_actions.copy_(act)
            self._step(self.last_actions)

        side = torch.cuda.Stream(device=self.mine.device)
        side.wait_stream(torch.cuda.current_stream(self.mine.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(self.mine.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            body()
        self._graph = graph
        return graph.replay

    def close(self):
        self.mine.close()
