"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's RL step surface on top of des_oracle.

Follows ``RLDispatcher._step/_get_observation/_get_reward_dense`` (dispatch_algorithms/rl_dispatch.py:58-123,
194-509), ``preprocess_observation`` (utils/feature_processing.py:3-75) and the final step emitted by
``Mine.start_rl`` (mine.py:453-476).  One "env step" = one dispatch decision of whichever truck asks
next; the log holds for every step the float32[194] observation, reward, done flag, order kind, truck,
teacher suggestion and the info dict values.  Pinned against the unmodified reference by
``tests/golden/rl_*.npz`` (made by tests/golden/make_golden.py through oracle/ref_harness.py).
"""
import math

import numpy as np

from . import des_oracle as do

KINDS = {"init": 0, "haul": 1, "back": 2}


class RlRecorder:
    def __init__(self, oracle, policy, teacher="ShortestTripDispatcher", reward_mode="dense"):
        self.o = oracle
        self.reward_mode = reward_mode
        self.policy = policy          # policy(step_index, kind_code, obs194, n_choices) -> int
        self.teacher = teacher
        # per truck, per record type, in creation order: the reference concatenates each truck's
        # events_by_type lists in truck order (rl_dispatch.py:222,235,243,265-267), which fixes the fp sum order
        nt = len(oracle.trucks)
        self.iv = {k: [[] for _ in range(nt)] for k in ("wait", "shovel", "dumper", "init", "haul", "unhaul")}
        self._open = {}
        self.last_production = 0
        self.last_time = 0
        self.last_order_time = 0
        self.last_truck = None
        self.last_action = None
        self.log = []
        self.current_obs = None

    # interval bookkeeping = the start_time / end_time fields of the reference's event records
    def open(self, kind, t, now):
        rec = [now, None]
        key = kind
        if kind == "move":
            key = "unhaul" if t.cur > self.o.L else ("init" if t.cur == 0 else "haul")
        self.iv[key][t.idx].append(rec)
        self._open[(kind, t.idx)] = rec

    def close(self, kind, t, now):
        self._open.pop((kind, t.idx))[1] = now

    def _overlap(self, kinds, last, cur):
        total = 0
        for ti in range(len(self.o.trucks)):
            for kind in kinds:
                for start, end in self.iv[kind][ti]:
                    e = end or cur
                    if max(start, last) < min(e, cur):
                        total += min(e, cur) - max(start, last)
        return total

    def reward_dense(self, final=False):
        """rl_dispatch.py:194-308."""
        o = self.o
        last, cur = self.last_order_time, o.env.now
        wait = self._overlap(("wait",), last, cur)
        service = self._overlap(("shovel",), last, cur) + self._overlap(("dumper",), last, cur)
        jam = 0
        for (_rid, _pos, start, end) in o.jams:
            e = end or cur
            if max(start, last) < min(e, cur):
                jam += min(e, cur) - max(start, last)
        move = self._overlap(("init", "haul", "unhaul"), last, cur)
        final_reward = 0
        if final and cur != 0:
            final_tons = self.last_production - 10000
            if final_tons > 0:
                final_reward = 0.5 * final_tons
        cur_prod = o.mine_produce_tons
        delta = cur_prod - self.last_production
        self.last_production = cur_prod
        reward = -0.1 * wait + -0.05 * service + -0.05 * jam + -0.01 * move + 5.0 * np.log(delta + 1) + final_reward
        dt = cur - last
        if dt > 0:
            reward += 0.1 * (delta / (dt + 1e-6))
        return float(reward)

    def observation(self, kind, truck):
        """rl_dispatch.py:309-509 + feature_processing.py:3-75 -> float32[194] (L = D = 5 layout generalised)."""
        o = self.o
        L, D, T = o.L, o.D, len(o.trucks)
        now = o.env.now
        onehot = [0] * (1 + L + D)
        onehot[truck.cur] = 1
        if kind == "haul":
            on = [o._on_road(truck.cur, 1 + L + j) for j in range(D)]
            extra = [(sum([1 for _ in r]) / len(o.dump_sites[j].units)) * o.dump_sites[j].units[0].cycle
                     for j, r in enumerate(on)]
            sig = np.array([s.est for s in o.dump_sites]) + np.array(extra)
        else:
            start = 0 if kind == "init" else truck.cur
            on = [o._on_road(start, 1 + i) for i in range(L)]
            extra = [sum([t.cap for t in r]) / o.load_sites[i].productivity for i, r in enumerate(on)]
            sig = np.array([s.est for s in o.load_sites]) + np.array(extra)
        sites = o.load_sites + o.dump_sites
        queue_lengths = [s.queue_total_cur for s in sites]
        capacities = [s.productivity for s in o.load_sites] + [s.dump_productivity for s in o.dump_sites]
        est_wait = [s.est for s in sites]
        ratio = [s.ability for s in o.load_sites] + [1 for _ in o.dump_sites]
        produced = [s.produced for s in sites]
        # road status (mine.update_road_status, mine.py:157-225): pair counts regardless of truck status
        pair = {}
        for t in o.trucks:
            if t.cur == do.NO_LOC or t.tgt == do.NO_LOC:
                continue
            # ChargingSite (loc 0) is falsy-safe in the reference (objects), so every located truck counts
            pair[(t.cur, t.tgt)] = pair.get((t.cur, t.tgt), 0) + 1
        jam_cnt = {}
        for (rid, _pos, start, end) in o.jams:
            if start <= now and end >= now:
                jam_cnt[rid] = jam_cnt.get(rid, 0) + 1
        roads = [(0, 1 + i) for i in range(L)] + [(1 + i, 1 + L + j) for i in range(L) for j in range(D)] + \
                [(1 + L + j, 1 + i) for j in range(D) for i in range(L)]
        oh_dist = [float(o._base_distance(a, b)) for a, b in roads]
        oh_jam = [jam_cnt.get(k, 0) for k in range(len(roads))]
        if kind == "init":
            cand = [(0, 1 + i) for i in range(L)]
        elif kind == "haul":
            cand = [(truck.cur, 1 + L + j) for j in range(D)]
        else:
            cand = [(truck.cur, 1 + i) for i in range(L)]
        distances = [float(o._base_distance(a, b)) for a, b in cand]
        truck_counts = [pair.get(k, 0) for k in cand]
        delta_time = now - self.last_time
        self.last_time = now
        ev = [[1, 0, 0], [0, 1, 0], [0, 0, 1]][KINDS[kind]]
        time_now = float(now) / o.sim_time
        order_state = np.array([ev[0], ev[1], ev[2], float(delta_time), time_now, 1 - time_now])
        truck_state = np.concatenate([np.array(onehot), np.array([np.log(truck.load + 1), np.log(truck.cycle_time + 1)])])
        road_states = np.concatenate([np.array(distances) * 60 / 25, np.array(truck_counts) / (T + 1e-8),
                                      np.array(oh_dist), np.array(oh_jam)])
        tar = np.concatenate([np.log(sig + 1), np.log(np.array(est_wait) + 1), np.array(queue_lengths) / (T + 1e-8),
                              np.log(np.array(capacities) + 1), np.array(ratio), np.log(np.array(produced) + 1)])
        state = np.concatenate([order_state, truck_state, road_states, tar])
        return state.astype(np.float32), float(delta_time)

    def step(self, kind, truck):
        """RLDispatcher._step (rl_dispatch.py:58-123): observation, reward for the previous decision, action."""
        o = self.o
        obs, dt = self.observation(kind, truck)
        self.current_obs = obs
        # sparse mode (rl_dispatch.py:69-70,186-192): 0 at every decision -- the dense function is not called, so
        # last_production stays 0 until Mine.start_rl calls the dense function once at the end (mine.py:456-459)
        reward = self.reward_dense(final=False) if self.reward_mode == "dense" else 0.0
        sug = o._decide(self.teacher, kind, truck) if self.teacher else -1
        n = o.D if kind == "haul" else o.L
        action = int(self.policy(len(self.log), KINDS[kind], obs, n))
        self.log.append({"obs": obs, "reward": reward, "done": False, "kind": KINDS[kind], "truck": truck.idx,
                         "sug_action": int(sug), "time": float(o.env.now), "delta_time": dt,
                         "produce_tons": float(o.mine_produce_tons), "action": action})
        self.last_truck, self.last_action = truck, action
        self.last_order_time = o.env.now
        return action

    def finish(self):
        """Mine.start_rl (mine.py:453-476): last observation again, final reward, done=True."""
        o = self.o
        reward = self.reward_dense(final=True)
        self.log.append({"obs": self.current_obs, "reward": reward, "done": True, "kind": -1, "truck": -1,
                         "sug_action": -1, "time": float(o.env.now), "delta_time": 0.0,
                         "produce_tons": float(o.mine_produce_tons), "action": -1})


def run_rl_oracle(config, policy, mode="det", seed=0, tape=None, teacher="ShortestTripDispatcher", sim_time=None,
                  reward_mode="dense"):
    cfg = dict(config)
    if sim_time is not None:
        cfg["sim_time"] = sim_time
    o = do.DesOracle(cfg, "external", mode, seed, tape)
    o.rl = RlRecorder(o, policy, teacher, reward_mode)
    res = o.run()
    # env.run(until=T) leaves env.now == T (the stop event is processed at T)
    o.env._now = float(cfg["sim_time"])
    o.rl.finish()
    res.rl_log = o.rl.log
    return res


def keyed_policy(policy_seed):
    """Deterministic pseudo-random policy shared by the oracle, the reference harness and the GPU tests."""
    from . import philox

    def policy(step, kind, obs, n):
        return philox.randint_keyed(policy_seed, step, 0, n, site=13)
    return policy
