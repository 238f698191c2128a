"""CPU tests (no GPU): the oracle restatement against the golden vectors recorded from the
unmodified reference, the SimPy stand-in's ordering semantics, and anchors from SURVEY.md."""
import math
import os
import sys

import numpy as np
import pytest

from helpers import DISPATCHERS, Golden
from oracle import des_oracle as do
from oracle import trace_format as tf


def _same(g, o, check_tape=False):
    d = []
    if not np.array_equal(g.trace, o.trace):
        a, b = g.trace, o.trace
        n = min(len(a), len(b))
        ne = [i for i in range(n) if a[i] != b[i]]
        d.append("trace differs (len %d vs %d, first %s)" % (len(a), len(b), ne[:1]))
    for k in o.tick:
        if not np.array_equal(g.tick[k], o.tick[k], equal_nan=(o.tick[k].dtype.kind == "f")):
            d.append("tick %s" % k)
    for k in o.final:
        if not np.array_equal(g.final[k], o.final[k]):
            d.append("final %s" % k)
    for k in o.orders:
        if not np.array_equal(g.orders[k], o.orders[k]):
            d.append("orders %s" % k)
    if g.reported_tons != o.reported_tons or g.dumped_tons != o.dumped_tons:
        d.append("tons %r/%r vs %r/%r" % (g.reported_tons, g.dumped_tons, o.reported_tons, o.dumped_tons))
    if g.n_truck_events != o.n_truck_events or g.order_count != o.order_count:
        d.append("counts")
    if check_tape and not (np.array_equal(g.tape[0], o.tape[0]) and np.array_equal(g.tape[1], o.tape[1])):
        d.append("tape")
    return d


@pytest.mark.parametrize("d", DISPATCHERS)
def test_oracle_det_matches_reference_golden(d, north_pit):
    g = Golden("det_north_pit_mine_%s_s0" % d)
    o = do.run_oracle(north_pit, d, "det", seed=0)
    assert not _same(g, o)


def test_oracle_det_random_seed7_and_short(north_pit, short_mine):
    assert not _same(Golden("det_north_pit_mine_RandomDispatcher_s7"), do.run_oracle(north_pit, "RandomDispatcher", "det", seed=7))
    for d in DISPATCHERS:
        assert not _same(Golden("det_north_pit_mine_short_%s_s0" % d), do.run_oracle(short_mine, d, "det", seed=0)), d


@pytest.mark.parametrize("d", ["RandomDispatcher", "FixedGroupDispatcher", "SPTFDispatcher", "SQDispatcher"])
def test_oracle_live_same_seed_matches_reference_golden(d, north_pit):
    """Same numpy / random seeds -> the restatement reproduces the reference's stochastic run and its tape."""
    g = Golden("live_north_pit_mine_%s_s11" % d)
    o = do.run_oracle(north_pit, d, "live", seed=11)
    assert not _same(g, o, check_tape=True)


@pytest.mark.parametrize("d", ["NaiveDispatcher", "NearestDispatcher"])
def test_oracle_replay_matches_reference_golden(d, north_pit):
    g = Golden("live_north_pit_mine_%s_s11" % d)
    o = do.run_oracle(north_pit, d, "replay", tape=g.tape)
    assert not _same(g, o, check_tape=True)


def test_survey_anchors():
    """SURVEY.md 8c probe anchors (independent SimPy shim of the surveyor): tons / events / orders."""
    exp = {"NaiveDispatcher": (2313.0, 615, 171), "NearestDispatcher": (10203.0, 1897, 493),
           "FixedGroupDispatcher": (11025.0, 2004, 531), "SPTFDispatcher": (14402.0, 2593, 966),
           "SQDispatcher": (13811.0, 2533, 661)}
    for d, (tons, ev, orders) in exp.items():
        g = Golden("det_north_pit_mine_%s_s0" % d)
        assert (g.reported_tons, g.n_truck_events, g.order_count) == (tons, ev, orders)
    steps = {"NaiveDispatcher": 33489, "SQDispatcher": 37325}
    for d, n in steps.items():
        assert int(Golden("det_north_pit_mine_%s_s0" % d).z["n_steps"]) == n
    g = Golden("det_north_pit_mine_NaiveDispatcher_s0")
    tons = g.tick["produced_tons"]
    assert [tons[k] for k in (30, 60, 120, 180, 239)] == [0, 385, 973, 1658, 2313]
    first_waits = g.trace[g.trace["kind"] == tf.K_WAIT_SHOVEL][:15]
    assert (first_waits["time"] == 7.199999999999999).all()
    gets = g.trace[g.trace["kind"] == tf.K_GET_SHOVEL]
    assert gets["time"][5] == 41.42222222222222 and 75.64444444444445 in gets["time"]
    t10 = gets[gets["truck"] == 10]
    assert t10["time"][0] == 75.64444444444445 and t10["arg0"][0] == 0      # truck 10: third in line at Shovel-1
    sq = Golden("det_north_pit_mine_SQDispatcher_s0")
    assert list(sq.orders["init"][:10]) == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4]
    assert sq.trace[sq.trace["kind"] == tf.K_WAIT_SHOVEL]["time"][0] == 5.04


def test_py_sum_matches_interpreter():
    rng = np.random.RandomState(0)
    for _ in range(200):
        xs = [float(x) for x in rng.uniform(-1e3, 1e3, size=rng.randint(1, 40)) * rng.choice([1e-9, 1, 1e9])]
        want = sum(xs) if sys.version_info >= (3, 12) else math.fsum(xs)
        if sys.version_info >= (3, 12):
            assert do.py_sum(xs) == want
        naive = 0.0
        for x in xs:
            naive += x
        assert do.py_sum(xs, "naive") == naive


def test_simpy_standin_ordering():
    """URGENT before NORMAL at equal time, FIFO by schedule order, process-return hop, until semantics."""
    sys.path.insert(0, os.path.join(os.path.dirname(do.__file__), "simpy_standin"))
    import simpy
    env = simpy.Environment()
    log = []

    def child(name, d):
        log.append(("child-start", name, env.now))
        yield env.timeout(d)
        log.append(("child-end", name, env.now))

    def parent(name, d):
        log.append(("start", name, env.now))
        yield env.process(child(name, d))
        log.append(("resumed", name, env.now))

    env.process(parent("a", 1))
    env.process(parent("b", 1))
    env.run(until=1)            # nothing NORMAL at t == until executes
    assert log == [("start", "a", 0), ("start", "b", 0), ("child-start", "a", 0), ("child-start", "b", 0)]
    env.run(until=2)
    assert log[4:] == [("child-end", "a", 1), ("child-end", "b", 1), ("resumed", "a", 1), ("resumed", "b", 1)]

    # resource: users cleared at release call, hand-over one hop later, head-only trigger
    env = simpy.Environment()
    res = simpy.Resource(env, capacity=1)
    seen = []

    def user(name, hold):
        with res.request() as req:
            seen.append((name, "queued", len(res.queue), res.count, env.now))
            yield req
            seen.append((name, "granted", env.now))
            yield env.timeout(hold)
        seen.append((name, "released", res.count, len(res.queue), env.now))

    env.process(user("x", 2))
    env.process(user("y", 1))
    env.run(until=10)
    assert seen[0] == ("x", "queued", 0, 1, 0)      # granted at once: not in queue, count 1
    assert seen[1] == ("y", "queued", 1, 1, 0)
    assert ("x", "released", 0, 1, 2) in seen        # count drops immediately, y still queued
    assert ("y", "granted", 2) in seen


def test_oracle_sq_keyerror_is_reported(short_mine):
    """An order at an exact integer minute before that minute's monitor block raises KeyError in the
    reference (sq_dispatcher.py:31); the oracle reports it instead of crashing."""
    cfg = dict(short_mine)
    cfg = {**cfg, "road": {**cfg["road"], "charging_to_load_road_matrix": [1.25, 1.25, 1.25]}}   # 3.0 min legs
    o = do.run_oracle(cfg, "SQDispatcher", "det")
    assert o.error is None or "KeyError" in o.error


@pytest.mark.parametrize("name", ["rl_north_pit_mine_det_s0_ShortestTripDispatcher_p1",
                                  "rl_north_pit_mine_live_s21_ShortestTripDispatcher_p2",
                                  "rl_north_pit_mine_live_s22_SQDispatcher_p3",
                                  "rl_north_pit_mine_live_s23_ShortestTripDispatcher_p4_sparse"])
def test_rl_oracle_matches_reference_golden(name, north_pit):
    """The RL restatement (observation float32[194], dense reward, teacher suggestion, final step) is
    bit-identical to the reference's RLDispatcher / start_rl / preprocess_observation."""
    import os
    from helpers import GOLDEN
    from oracle import rl_oracle as ro
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mode, seed, teacher, pseed = str(z["mode"]), int(z["seed"]), str(z["teacher"]), int(z["policy_seed"])
    rmode = "sparse" if name.endswith("_sparse") else "dense"
    if mode == "live":   # replay the reference's own draws
        r = ro.run_rl_oracle(north_pit, ro.keyed_policy(pseed), "replay", tape=(z["tape_sites"], z["tape_values"]),
                             teacher=teacher, reward_mode=rmode)
    else:
        r = ro.run_rl_oracle(north_pit, ro.keyed_policy(pseed), mode, seed=seed, teacher=teacher, reward_mode=rmode)
    log = r.rl_log
    assert len(log) == len(z["reward"])
    assert np.array_equal(np.stack([x["obs"] for x in log]), z["obs"])
    assert np.array_equal(np.array([x["reward"] for x in log]), z["reward"])
    assert np.array_equal(np.array([x["kind"] for x in log]), z["kind"])
    assert np.array_equal(np.array([x["sug_action"] for x in log]), z["sug_action"])
    assert np.array_equal(np.array([x["action"] for x in log]), z["action"])
    assert np.array_equal(np.array([x["done"] for x in log]), z["done"])


# ---- edge-case mines (tests/golden/make_golden_edge.py: goldens from the unmodified reference) -------------------------
@pytest.mark.parametrize("d", DISPATCHERS)
def test_oracle_ragged_mine_det(d, ragged_mine):
    """L != D, 1..9 units per site, 130 trucks of four types.  RandomDispatcher draws haul orders over len(load_sites)
    (random_dispatcher.py:21-22), so with L = 3 > D = 2 the reference dies with IndexError -- the oracle must stop at the
    same record."""
    g = Golden("det_ragged_mine_%s_s0" % d)
    o = do.run_oracle(ragged_mine, d, "det", seed=0)
    assert not _same(g, o)
    assert (g.error is None) == (o.error is None)
    if d == "RandomDispatcher":
        assert "IndexError" in g.error and "IndexError" in o.error


@pytest.mark.parametrize("d", ["SQDispatcher", "SPTFDispatcher", "RandomDispatcher"])
def test_oracle_ragged_mine_live(d, ragged_mine):
    g = Golden("live_ragged_mine_%s_s5" % d)
    o = do.run_oracle(ragged_mine, d, "live", seed=5)
    assert not _same(g, o, check_tape=True)


def test_oracle_one_truck_mine(one_truck_mine):
    """Smallest legal mine: 2 + 8 records per cycle."""
    g = Golden("det_one_truck_mine_NaiveDispatcher_s0")
    o = do.run_oracle(one_truck_mine, "NaiveDispatcher", "det", seed=0)
    assert not _same(g, o)
    assert g.n_truck_events == 30 and (g.n_truck_events - 2) % 4 == 0
    g = Golden("live_one_truck_mine_NaiveDispatcher_s2")
    o = do.run_oracle(one_truck_mine, "NaiveDispatcher", "live", seed=2)
    assert not _same(g, o, check_tape=True)


@pytest.mark.parametrize("d", DISPATCHERS)
def test_oracle_integer_mine_tie_stress(d, integer_mine):
    """All durations are whole minutes, so truck hops, Releases and the per-minute monitors tie constantly."""
    assert not _same(Golden("det_integer_mine_%s_s0" % d), do.run_oracle(integer_mine, d, "det", seed=0))


@pytest.mark.parametrize("cname", ["north_pit_mine", "ragged_mine"])
def test_oracle_shortest_trip_matches_reference_golden(cname, north_pit, ragged_mine):
    """ShortestTripDispatcher run AS A DISPATCHER by the unmodified reference (tests/golden/make_golden_teacher.py): det,
    live with the same seeds (tape included) and tape replay."""
    cfg = north_pit if cname == "north_pit_mine" else ragged_mine
    d = "ShortestTripDispatcher"
    assert not _same(Golden("det_%s_%s_s0" % (cname, d)), do.run_oracle(cfg, d, "det"))
    g = Golden("live_%s_%s_s17" % (cname, d))
    assert not _same(g, do.run_oracle(cfg, d, "live", seed=17), check_tape=True)
    assert not _same(g, do.run_oracle(cfg, d, "replay", tape=g.tape), check_tape=True)


def test_device_philox_trajectories_replayed_by_the_reference(north_pit, ragged_mine):
    """tests/golden/device_tapes.npz holds 12 trajectories the PHILOX-mode CUDA kernels produced on a B200 together with the
    draw tapes they exported (tools/export_device_tapes.py).  Replaying each tape through the CPU restatement -- and, where
    /root/reference exists, through the UNMODIFIED reference itself (oracle/ref_harness.py) -- must give the device's event
    trace (times included), production and waiting-truck curves, tonnage, event / order counts and order lists bit for bit,
    consuming the tape to its last entry (tests/golden/make_golden_device_tapes.py)."""
    import sys
    from helpers import GOLDEN
    sys.path.insert(0, GOLDEN)
    from make_golden_device_tapes import cases, compare
    z = np.load(os.path.join(GOLDEN, "device_tapes.npz"))
    from oracle import ref_harness as rh
    use_ref = rh.reference_available()
    n = 0
    for key, cname, disp in cases(z):
        cfg = north_pit if cname == "north_pit_mine" else ragged_mine
        tape = (z[key + "_tape_sites"], z[key + "_tape_values"])
        o = do.run_oracle(cfg, disp, "replay", tape=tape)
        assert o.error is None and not compare(key, z, o, disp == "SPTFDispatcher"), key
        if use_ref and n % 3 == 0:              # the reference itself on every third one (it is ~5x slower)
            r = rh.run_reference(cfg, disp, "replay", tape=tape)
            assert r.error is None and not compare(key, z, r, disp == "SPTFDispatcher"), key
        n += 1
    assert n == 12
