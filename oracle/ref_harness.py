"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference (370025263/openmines) here.

This module exists only in the authoring container's workflow: it imports the reference's
own Python modules from ``/root/reference`` (which does NOT exist on the GPU box), on top of
the SimPy stand-in (``oracle/simpy_standin``) and ``MagicMock`` stubs for the plotting / solver
packages that are missing from the image, and drives the reference's own wiring function
``openmines.src.cli.run.run_dispatch_sim`` (run.py:48-112).  It is used to

* pin the CPU restatement (``oracle/des_oracle.py``) against the real reference code, and
* generate the committed golden vectors under ``tests/golden`` (``tests/golden/make_golden.py``).

Nothing of the reference is modified: behaviour is observed by wrapping
``EventPool.add_event`` (event.py:35) and ``Mine.start`` (mine.py:340), and randomness is
controlled by replacing the module-level draw functions the reference calls
(``np.random.uniform/exponential/normal/weibull/choice``, ``random.random/expovariate``;
SURVEY.md section 8a-R) for the duration of one run.

RNG modes
  det    no stochastic event can fire: jitter 0.0, every exponential gap +inf, jam coin 1.0
         (SURVEY 8a-R).  RandomDispatcher draws come from keyed Philox (oracle/philox.py).
  live   numpy / random seeded with ``seed``; every draw is recorded on a tape (site id,
         post-transform value) in call order.
  replay draws are served from a previously recorded tape (site ids are checked).
"""
import contextlib
import logging
import math
import os
import random as _pyrandom
import sys
import tempfile
import types
from unittest import mock

import numpy as np

from . import philox
from . import trace_format as tf

REF_ROOT_DEFAULT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_state = {"loaded": False, "mods": None}


def reference_available(ref_root=REF_ROOT_DEFAULT):
    return os.path.isdir(os.path.join(ref_root, "openmines", "src"))


def _install_stubs():
    try:
        import simpy  # noqa: F401  (a real SimPy is preferred if it is ever installed)
    except ImportError:
        sys.path.insert(0, os.path.join(_HERE, "simpy_standin"))
        import simpy  # noqa: F401
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.offsetbox",
                 "matplotlib.legend_handler", "matplotlib.image", "matplotlib.animation",
                 "gurobipy", "gymnasium", "gymnasium.spaces"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = mock.MagicMock(name=name)
    if "inflection" not in sys.modules:
        try:
            import inflection  # noqa: F401
        except ImportError:
            import re
            m = types.ModuleType("inflection")

            def underscore(word):
                word = re.sub(r"([A-Z]+)([A-Z][a-z])", r"\1_\2", word)
                word = re.sub(r"([a-z\d])([A-Z])", r"\1_\2", word)
                return word.replace("-", "_").lower()
            m.underscore = underscore
            sys.modules["inflection"] = m


def load_reference(ref_root=REF_ROOT_DEFAULT):
    """Import the reference's modules (unchanged) and return a namespace of what we use."""
    if _state["loaded"]:
        return _state["mods"]
    if not reference_available(ref_root):
        raise RuntimeError("reference tree not present at %s" % ref_root)
    _install_stubs()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    logging.disable(logging.CRITICAL)
    ns = types.SimpleNamespace()
    import openmines.src.cli.run as run_mod
    import openmines.src.mine as mine_mod
    import openmines.src.utils.event as event_mod
    from openmines.src.dispatch_algorithms.naive_dispatcher import NaiveDispatcher
    from openmines.src.dispatch_algorithms.random_dispatcher import RandomDispatcher
    from openmines.src.dispatch_algorithms.nearest_dispatcher import NearestDispatcher
    from openmines.src.dispatch_algorithms.fixed_group_dispatcher import FixedGroupDispatcher
    from openmines.src.dispatch_algorithms.sptf_dispatcher import SPTFDispatcher
    from openmines.src.dispatch_algorithms.sq_dispatcher import SQDispatcher
    from openmines.src.dispatch_algorithms.shortest_trip_dispatcher import ShortestTripDispatcher
    ns.run_mod, ns.mine_mod, ns.event_mod = run_mod, mine_mod, event_mod
    ns.dispatchers = {c.__name__: c for c in (NaiveDispatcher, RandomDispatcher, NearestDispatcher,
                                              FixedGroupDispatcher, SPTFDispatcher, SQDispatcher, ShortestTripDispatcher)}
    import openmines.src.load_site as ls_mod
    import openmines.src.dump_site as ds_mod
    import openmines.src.charging_site as cs_mod
    ns.LoadSite, ns.DumpSite, ns.ChargingSite = ls_mod.LoadSite, ds_mod.DumpSite, cs_mod.ChargingSite
    _state["loaded"], _state["mods"] = True, ns
    return ns


# ----------------------------------------------------------------------------------------
# RNG control
# ----------------------------------------------------------------------------------------
class TapeExhausted(RuntimeError):
    pass


class RngControl:
    """Replaces the module-level draw functions the reference calls (see module docstring)."""

    def __init__(self, mode, seed=0, tape=None):
        assert mode in ("det", "live", "replay")
        self.mode, self.seed = mode, int(seed)
        self.sites, self.values = [], []
        self._tape = tape
        self._pos = 0
        self._x = np.linspace(0, 1, 1000)
        self.order_epoch = {}       # truck index -> number of orders given so far
        self.current_truck = None   # set by the dispatcher shim around each order

    # -- helpers -------------------------------------------------------------------------
    def _caller(self, depth=2):
        return sys._getframe(depth).f_code.co_name

    def _emit(self, site, value):
        self.sites.append(site)
        self.values.append(float(value))
        return value

    def _next(self, site):
        sites, values = self._tape
        if self._pos >= len(values):
            raise TapeExhausted("tape exhausted at draw %d (site %d)" % (self._pos, site))
        if int(sites[self._pos]) != site:
            raise RuntimeError("tape site mismatch at %d: tape %d, asked %d" % (self._pos, sites[self._pos], site))
        v = float(values[self._pos])
        self._pos += 1
        return self._emit(site, v)

    # -- replacement functions -----------------------------------------------------------
    def uniform(self, low=0.0, high=1.0, size=None):
        assert size is None and (low, high) == (-0.1, 0.1), "unexpected uniform() call site"
        if self.mode == "det":
            return 0.0
        if self.mode == "replay":
            return self._next(tf.S_LOAD_JITTER)
        return self._emit(tf.S_LOAD_JITTER, self._orig["uniform"](low, high))

    def exponential(self, scale=1.0, size=None):
        who = self._caller()
        site = {"run": tf.S_MAINT_GAP, "sample_breakdown": tf.S_BREAK_GAP}[who]
        if self.mode == "det":
            return math.inf
        if self.mode == "replay":
            return self._next(site)
        return self._emit(site, self._orig["exponential"](scale))

    def normal(self, loc=0.0, scale=1.0, size=None):
        who = self._caller()
        site = {"run": tf.S_MAINT_DUR, "sample_breakdown": tf.S_BREAK_DUR,
                "check_availability": tf.S_ROAD_DUR, "get_distance": tf.S_DETOUR}[who]
        if self.mode == "det":
            raise AssertionError("normal() is unreachable in det mode (site %d)" % site)
        if self.mode == "replay":
            return self._next(site)
        return self._emit(site, self._orig["normal"](loc, scale))

    def weibull(self, a, size=None):
        if self.mode == "det":
            raise AssertionError("weibull() is unreachable in det mode")
        if self.mode == "replay":
            return self._next(tf.S_JAM_DUR)
        return self._emit(tf.S_JAM_DUR, self._orig["weibull"](a))

    def choice(self, a, size=None, replace=True, p=None):
        assert len(a) == 1000, "unexpected choice() call site"
        if self.mode == "det":
            raise AssertionError("choice() is unreachable in det mode")
        if self.mode == "replay":
            x = self._next(tf.S_JAM_POS)
            idx = int(np.argmin(np.abs(self._x - x)))
            assert self._x[idx] == x
            return np.int64(idx)
        idx = self._orig["choice"](a, p=p)
        self._emit(tf.S_JAM_POS, self._x[int(idx)])
        return idx

    def expovariate(self, lambd):
        who = self._caller()
        site = {"check_vehicle_availability": tf.S_UNREPAIRABLE, "check_availability": tf.S_ROAD_GAP}[who]
        if self.mode == "det":
            return math.inf
        if self.mode == "replay":
            return self._next(site)
        return self._emit(site, self._orig["expovariate"](lambd))

    def random(self):
        assert self._caller() == "road_jam_sampling", "unexpected random() call site"
        if self.mode == "det":
            return 1.0
        if self.mode == "replay":
            return self._next(tf.S_JAM_COIN)
        return self._emit(tf.S_JAM_COIN, self._orig["random"]())

    # RandomDispatcher.np_random stand-in (random_dispatcher.py:10,16,22,29)
    def randint(self, low, high=None):
        assert low == 0 and high is not None
        if self.mode == "det":
            t = self.current_truck
            epoch = self.order_epoch.get(t, 0)
            return philox.randint_keyed(self.seed, epoch, t, high)
        if self.mode == "replay":
            return int(self._next(tf.S_DISPATCH))
        return int(self._emit(tf.S_DISPATCH, self._disp_rs.randint(low, high)))

    # -- context manager -------------------------------------------------------------------
    def __enter__(self):
        self._orig = {k: getattr(np.random, k) for k in ("uniform", "exponential", "normal", "weibull", "choice")}
        self._orig["expovariate"] = _pyrandom.expovariate
        self._orig["random"] = _pyrandom.random
        if self.mode == "live":
            np.random.seed(self.seed)
            _pyrandom.seed(self.seed)
            self._disp_rs = np.random.RandomState(self.seed)
        for k in ("uniform", "exponential", "normal", "weibull", "choice"):
            setattr(np.random, k, getattr(self, k))
        _pyrandom.expovariate = self.expovariate
        _pyrandom.random = self.random
        return self

    def __exit__(self, *exc):
        for k in ("uniform", "exponential", "normal", "weibull", "choice"):
            setattr(np.random, k, self._orig[k])
        _pyrandom.expovariate = self._orig["expovariate"]
        _pyrandom.random = self._orig["random"]
        return False

    def tape(self):
        return np.asarray(self.sites, dtype=np.uint8), np.asarray(self.values, dtype=np.float64)


# ----------------------------------------------------------------------------------------
# One reference run
# ----------------------------------------------------------------------------------------
class RefResult:
    """Everything the parity tests compare (numpy arrays, JSON-free)."""
    pass


def _loc_code(ns, mine, site):
    if site is None:
        return 0xFF
    if isinstance(site, ns.ChargingSite):
        return 0
    if isinstance(site, ns.LoadSite):
        return 1 + mine.load_sites.index(site)
    return 1 + len(mine.load_sites) + mine.dump_sites.index(site)


def run_reference(config, dispatcher_name, mode="det", seed=0, tape=None, sim_time=None,
                  dump_frames=False, ref_root=REF_ROOT_DEFAULT):
    """Run ``run_dispatch_sim`` of the unmodified reference once and extract the canonical trace.

    ``config`` is a dict (the mine JSON).  Returns a RefResult.
    """
    ns = load_reference(ref_root)
    cfg = dict(config)
    if sim_time is not None:
        cfg["sim_time"] = sim_time
    disp_cls = ns.dispatchers[dispatcher_name]
    rng = RngControl(mode, seed, tape)
    cap = {"mine": None, "records": [], "error": None}

    # dispatcher shim: tells the RNG control which truck is asking (needed only for keyed randint);
    # the subclass adds no behaviour of its own.
    class _Shim(disp_cls):
        def give_init_order(self, truck, mine):
            rng.current_truck = mine.trucks.index(truck)
            r = super().give_init_order(truck, mine)
            rng.order_epoch[rng.current_truck] = rng.order_epoch.get(rng.current_truck, 0) + 1
            return r

        def give_haul_order(self, truck, mine):
            rng.current_truck = mine.trucks.index(truck)
            r = super().give_haul_order(truck, mine)
            rng.order_epoch[rng.current_truck] = rng.order_epoch.get(rng.current_truck, 0) + 1
            return r

        def give_back_order(self, truck, mine):
            rng.current_truck = mine.trucks.index(truck)
            if dispatcher_name == "SPTFDispatcher":
                # SPTF's back order re-enters give_init_order through the wrapper (sptf_dispatcher.py:59-61);
                # count the decision once.
                before = rng.order_epoch.get(rng.current_truck, 0)
                r = super().give_back_order(truck, mine)
                rng.order_epoch[rng.current_truck] = before + 1
                return r
            r = super().give_back_order(truck, mine)
            rng.order_epoch[rng.current_truck] = rng.order_epoch.get(rng.current_truck, 0) + 1
            return r
    _Shim.__name__ = disp_cls.__name__
    dispatcher = _Shim()
    if hasattr(dispatcher, "np_random"):
        dispatcher.np_random = rng

    Mine = ns.mine_mod.Mine
    EventPool = ns.event_mod.EventPool
    orig_start, orig_add, orig_dump = Mine.start, EventPool.add_event, Mine.dump_frames

    def start_wrapper(self, total_time=60 * 8):
        cap["mine"] = self
        return orig_start(self, total_time=total_time)

    def add_wrapper(self, event):
        mine = cap["mine"]
        now = mine.env.now if mine is not None else 0.0
        info = dict(event.info) if event.info else {}
        cap["records"].append((id(self), float(now), event.event_type, float(event.time_stamp), info))
        return orig_add(self, event)

    def dump_wrapper(self, total_time, rl=False):
        cap["n_truck_events"] = sum(len(v) for t in self.trucks for v in t.event_pool.events_by_type.values())
        if dump_frames:
            return orig_dump(self, total_time, rl)
        return {}

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="om_ref_")
    os.chdir(tmp)
    try:
        with rng, contextlib.redirect_stdout(open(os.devnull, "w")):
            Mine.start, EventPool.add_event, Mine.dump_frames = start_wrapper, add_wrapper, dump_wrapper
            try:
                ticks = ns.run_mod.run_dispatch_sim(dispatcher, cfg)
            except TapeExhausted:
                raise
            except Exception as e:  # reference exceptions abort the run; keep what happened
                cap["error"] = "%s: %s" % (type(e).__name__, e)
                ticks = None
    finally:
        Mine.start, EventPool.add_event, Mine.dump_frames = orig_start, orig_add, orig_dump
        os.chdir(cwd)

    mine = cap["mine"]
    res = RefResult()
    res.error = cap["error"]
    res.ticks = ticks
    res.mine = mine
    res.dispatcher = dispatcher
    res.tape = rng.tape()
    res.n_steps = getattr(mine.env, "n_steps", -1)
    res.sim_time = cfg["sim_time"]
    res.python = tuple(sys.version_info[:3])
    _extract(ns, mine, dispatcher, cap, res)
    return res


def _extract(ns, mine, dispatcher, cap, res):
    L, D = len(mine.load_sites), len(mine.dump_sites)
    trucks = mine.trucks
    shovels = [s for ls in mine.load_sites for s in ls.shovel_list]
    dumpers = [d for ds in mine.dump_sites for d in ds.dumper_list]
    truck_pool = {id(t.event_pool): i for i, t in enumerate(trucks)}
    shovel_pool = {id(s.event_pool): i for i, s in enumerate(shovels)}
    shovel_by_name = {}
    for i, s in enumerate(shovels):
        shovel_by_name.setdefault(s.name, i)
    dumper_by_name = {}
    for i, d in enumerate(dumpers):
        dumper_by_name.setdefault(d.name, i)
    site_by_name = {mine.charging_site.name: 0}
    for i, s in enumerate(mine.load_sites):
        site_by_name.setdefault(s.name, 1 + i)
    for j, s in enumerate(mine.dump_sites):
        site_by_name.setdefault(s.name, 1 + L + j)
    random_pool = id(mine.random_event_pool)

    recs = []
    maint_holder = {}
    for pool, now, etype, ts, info in cap["records"]:
        if pool in truck_pool:
            t = truck_pool[pool]
            if etype == "INIT ORDER":
                recs.append((now, info["distance"], 0.0, t, tf.K_INIT_ORDER, 0, info["order_index"], 0))
            elif etype in ("init", "haul", "unhaul"):
                kind = {"init": tf.K_MOVE_INIT, "haul": tf.K_MOVE_HAUL, "unhaul": tf.K_MOVE_UNHAUL}[etype]
                tgt = site_by_name[info["target_location"]]
                idx = tgt - 1 if tgt <= L else tgt - 1 - L
                recs.append((now, info["est_end_time"], info["distance"], t, kind, 0, idx, 0))
            elif etype == "wait shovel":
                recs.append((now, 0.0, 0.0, t, tf.K_WAIT_SHOVEL, 0, shovel_by_name[info["shovel"]], info["queue_index"]))
            elif etype == "get shovel":
                recs.append((now, 0.0, 0.0, t, tf.K_GET_SHOVEL, 0, shovel_by_name[info["shovel"]], 0))
            elif etype == "ORDER":
                tgt = site_by_name[info["target_location"]]
                if tgt > L:
                    recs.append((now, info["distance"], 0.0, t, tf.K_ORDER_HAUL, 0, info["order_index"], 0))
                else:
                    recs.append((now, info["distance"], 0.0, t, tf.K_ORDER_BACK, 0, info["order_index"], 0))
            elif etype == "wait dumper":
                recs.append((now, 0.0, 0.0, t, tf.K_WAIT_DUMPER, 0, dumper_by_name[info["dumper"]], 0))
            elif etype == "get dumper":
                recs.append((now, 0.0, 0.0, t, tf.K_GET_DUMPER, 0, dumper_by_name[info["dumper"]], 0))
            elif etype == "TruckEvent:breakdown":
                recs.append((now, info["start_time"], info["repair_time"], t, tf.K_BREAKDOWN, 0, 0, 0))
            elif etype == "TruckEvent:unrepairable":
                recs.append((now, 0.0, 0.0, t, tf.K_UNREPAIRABLE, 0, 0, 0))
            elif etype == "end":
                pass  # post-run artefact of get_wait_time (truck.py:479); not part of env.run
            else:
                raise RuntimeError("unknown truck record type %r" % etype)
        elif pool == random_pool:
            if etype == "RoadEvent:jam":
                r = tf.road_id(L, D, site_by_name[info["start_location"]], site_by_name[info["end_location"]])
                recs.append((now, info["jam_position"], info["est_end_time"], tf.NO_TRUCK, tf.K_JAM, 0, r, 0))
            elif etype == "RoadEvent:repair":
                r = tf.road_id(L, D, site_by_name[info["start_location"]], site_by_name[info["end_location"]])
                recs.append((now, info["repair_end_time"], info["repair_duration"], tf.NO_TRUCK, tf.K_ROAD_REPAIR, 0, r, 0))
            # Truck events are mirrored into the random pool (truck.py:144,189); already counted above.
        elif pool in shovel_pool:
            s = shovel_pool[pool]
            if etype == "Shovel Maintenance Start":
                # the truck holding the shovel is the one whose record precedes: find by users list later;
                # identify holder through the resource's users at extraction time is not possible, so
                # leave truck = NO_TRUCK here and let comparisons ignore the truck field for kinds 14/15.
                recs.append((now, info["end_time"], 0.0, tf.NO_TRUCK, tf.K_SHOVEL_MAINT_START, 0, s, 0))
            elif etype == "Shovel Maintenance Complete":
                recs.append((now, 0.0, 0.0, tf.NO_TRUCK, tf.K_SHOVEL_MAINT_DONE, 0, s, 0))
    trace = np.zeros(len(recs), dtype=tf.TRACE_DTYPE)
    for i, r in enumerate(recs):
        trace[i] = r
    res.trace = trace
    res.n_truck_events = int(np.isin(trace["kind"], tf.TRUCK_EVENT_KINDS).sum())
    if "n_truck_events" in cap:   # the reference's own count before dump_frames
        assert cap["n_truck_events"] == res.n_truck_events, (cap["n_truck_events"], res.n_truck_events)

    T = int(res.sim_time)
    nt = max([k for k in mine.status.keys() if isinstance(k, int)], default=-1) + 1
    res.n_ticks = nt
    f = lambda key, dt: np.array([mine.status[k][key] for k in range(nt)], dtype=dt)
    res.tick = {
        "produced_tons": f("produced_tons", np.float64),
        "service_count": f("service_count", np.int32),
        "working": f("working_truck_count", np.int32),
        "waiting": f("waiting_truck_count", np.int32),
        "load_unload": f("load_unload_truck_count", np.int32),
        "moving": f("moving_truck_count", np.int32),
        "repairing": f("repairing_truck_count", np.int32),
        "road_jam_count": f("road_jam_count", np.int32),
        "road_repair_count": f("road_repair_count", np.int32),
        "truck_repair": f("truck_repair", np.int32),
        "truck_unrepairable": f("truck_unrepairable", np.int32),
    }
    sites = list(mine.load_sites) + list(mine.dump_sites)
    res.tick["site_queue_total"] = np.array(
        [[s.parking_lot.queue_status["total"].get(k, -1) for s in sites] for k in range(nt)], dtype=np.int32)
    res.tick["site_tons"] = np.array(
        [[s.status[k]["produced_tons"] if k in s.status else np.nan for s in sites] for k in range(nt)], dtype=np.float64)
    res.reported_tons = float(mine.status[nt - 1]["produced_tons"]) if nt > 0 else 0.0
    res.final = {
        "truck_total_load": np.array([t.total_load_count for t in trucks], dtype=np.float64),
        "truck_service_count": np.array([t.service_count for t in trucks], dtype=np.int32),
        "truck_cycle_time": np.array([t.truck_cycle_time for t in trucks], dtype=np.float64),
        "truck_status": np.array([tf.STATUS_NAMES.index(t.status) for t in trucks], dtype=np.uint8),
        "truck_cur": np.array([_loc_code(ns, mine, t.current_location) for t in trucks], dtype=np.uint8),
        "truck_tgt": np.array([_loc_code(ns, mine, t.target_location) for t in trucks], dtype=np.uint8),
        "shovel_tons": np.array([s.produced_tons for s in shovels], dtype=np.float64),
        "shovel_service_count": np.array([s.service_count for s in shovels], dtype=np.int32),
        "dumper_tons": np.array([d.dumper_tons for d in dumpers], dtype=np.float64),
        "dumper_service_count": np.array([d.service_count for d in dumpers], dtype=np.int32),
    }
    res.dumped_tons = float(sum(float(d.dumper_tons) for d in dumpers))
    res.orders = {
        "init": np.array(dispatcher.init_orders, dtype=np.int32),
        "haul": np.array(dispatcher.haul_orders, dtype=np.int32),
        "back": np.array(dispatcher.back_orders, dtype=np.int32),
    }
    res.order_count = int(dispatcher.total_order_count)


def load_config(path):
    import json
    with open(path, "r", encoding="utf-8") as fh:
        return json.load(fh)


# ----------------------------------------------------------------------------------------
# RL surface of the unmodified reference (RLDispatcher + Mine.start_rl + preprocess_observation)
# ----------------------------------------------------------------------------------------
def run_reference_rl(config, policy, mode="det", seed=0, tape=None, teacher="ShortestTripDispatcher",
                     sim_time=None, ref_root=REF_ROOT_DEFAULT, reward_mode="dense"):
    """Drive the reference's own RL loop: ``prepare_env`` (utils/rl_env.py:29-104) runs in a thread exactly
    as ``ThreadMineEnv.reset`` starts it (mine_env.py:302-316); this thread plays the agent.
    policy(step_index, kind_code, obs194, n_choices) -> int.  Returns the per-step log."""
    import queue
    import threading
    load_reference(ref_root)
    for name in ("asyncore", "docutils", "docutils.nodes"):      # stale imports at rl_dispatch.py:6,11
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = mock.MagicMock(name=name)
    import openmines.src.utils.rl_env as rl_env
    from openmines.src.utils.feature_processing import preprocess_observation
    cfg = dict(config)
    if sim_time is not None:
        cfg["sim_time"] = sim_time
    cfg["sug_dispatcher"] = teacher
    rng = RngControl(mode, seed, tape)
    obs_q, act_q = queue.Queue(), queue.Queue()
    err = {}

    def target():
        try:
            rl_env.prepare_env(obs_q, act_q, cfg, reward_mode, cfg["sim_time"], False, False, None)
        except BaseException as e:   # surface crashes of the simulation thread
            err["e"] = e
            obs_q.put(None)

    log = []
    kinds = {"init": 0, "haul": 1, "unhaul": 2}
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="om_ref_rl_")
    os.chdir(tmp)
    try:
        with rng, contextlib.redirect_stdout(open(os.devnull, "w")):
            th = threading.Thread(target=target, daemon=True)
            th.start()
            while True:
                out = obs_q.get(timeout=120)
                if out is None:
                    raise err["e"]
                ob = out["ob"]
                obs = preprocess_observation(ob, cfg["sim_time"])
                done = bool(out["done"])
                kind = kinds[ob["event_name"]]
                n = len(cfg["dump_sites"]) if kind == 1 else len(cfg["load_sites"])
                rec = {"obs": obs, "reward": float(out["reward"]), "done": done,
                       "kind": -1 if done else kind, "sug_action": -1 if done else int(out["info"]["sug_action"]),
                       "time": float(out["info"]["time"]), "produce_tons": float(out["info"]["produce_tons"])}
                if done:
                    rec["action"] = -1
                    log.append(rec)
                    break
                action = int(policy(len(log), kind, obs, n))
                rec["action"] = action
                log.append(rec)
                act_q.put(action)
            th.join(timeout=60)
    finally:
        os.chdir(cwd)
    return log, rng.tape()
