"""Tick recorder: turns one instance's device outputs into the reference's result JSON and the
``states_dict`` that ``Charter.draw`` consumes.

Output contract: file ``results/MINE-{mine}-ALGO-{dispatcher}-TIME-{ts}.json``
(reference mine.py:528-529, utils/ticker.py:455-470) with keys ``"0".."T-1"`` (one packet per simulated
minute: ``truck_states, load_site_states, shovel_states, dump_site_states, dumper_states, mine_states``;
ticker.py:283-395) plus ``"summary"`` (ticker.py:440-453).  Per-minute site / shovel / dumper / mine values
are the monitor samples written by the engine (``OmTickRecord``, ``OmUnitTick``, site tick arrays); truck
states and positions are rebuilt on the host from the event trace exactly the way the reference rebuilds
them from the truck event pools (ticker.py:92-289), including the timestamp-keyed pool with its single
+1e-4 collision bump (utils/event.py:35-38).

Documented deviations: ``avg_time_per_order`` is 0 (decisions are taken on the GPU, there is no per-order
host latency to report); the ``road_jam_count / road_repair_count / truck_repair / truck_unrepairable``
counters are true event counts (the reference counts keys of a timestamp-keyed dict, mine.py:91-101, and
under-counts simultaneous events).
"""
import json
import math
import os
import time

import numpy as np

from . import trace as tr
from .config import py_sum
from .trace import TICK_DTYPE, UNIT_TICK_DTYPE


class _Ev:
    __slots__ = ("time_stamp", "type", "info")

    def __init__(self, ts, typ, info):
        self.time_stamp, self.type, self.info = ts, typ, info


class _Pool:
    """utils/event.py:26-48: dict keyed by timestamp (+1e-4 once on collision, then overwrite) + per-type lists."""

    def __init__(self):
        self.event_set = {}
        self.by_type = {}

    def add(self, ev):
        if ev.time_stamp in self.event_set:
            ev.time_stamp = ev.time_stamp + 0.0001
        self.event_set[ev.time_stamp] = ev
        lst = self.by_type.setdefault(ev.type, [])
        idx = 0
        while idx < len(lst) and lst[idx].time_stamp < ev.time_stamp:      # bisect_left on time_stamp
            idx += 1
        lst.insert(idx, ev)

    def past(self, t):
        return [self.event_set[k] for k in sorted(self.event_set) if k <= t]


class TickRecorder:
    def __init__(self, mine, result, dispatcher_name=None, reference_counters=True, config_index=0):
        """reference_counters: report road_jam_count etc. the way the reference counts them (keys of the
        timestamp-keyed random-event dict, needs the event trace) instead of the engine's true counts."""
        self.reference_counters = reference_counters
        self.mine, self.res = mine, result
        self.pc = mine.pcs[config_index]          # mixed-config batches: the config of the instances to be rendered
        self.cfg = self.pc.raw
        self.dispatcher_name = dispatcher_name
        if result.summary is not None:
            err = result.summary_np()["err"]
            if (err == 5).any():          # OM_IERR_TRACE_OVERFLOW: truck states would be rebuilt from a truncated event trace
                raise RuntimeError("TickRecorder: the event trace of instance %d overflowed its buffer (trace_capacity %d); "
                                   "run again with a larger trace_capacity" % (int(np.nonzero(err == 5)[0][0]), result.trace_capacity))
        cfg, pc = self.cfg, self.pc
        self.charging_pos = list(cfg["charging_site"]["position"])
        self.site_pos, self.park_pos = {}, {}
        self.shovel_pos, self.dumper_pos = [], []
        for ls in cfg["load_sites"]:
            self.site_pos[ls["name"]] = list(ls["position"])
            self.park_pos[ls["name"]] = tuple(a + b for a, b in zip(ls["position"], ls["parkinglot"]["position_offset"]))
            for sc in ls["shovels"]:
                self.shovel_pos.append(tuple(a + b for a, b in zip(ls["position"], sc["position_offset"])))   # load_site.py:198
        for ds in cfg["dump_sites"]:
            self.site_pos[ds["name"]] = list(ds["position"])
            self.park_pos[ds["name"]] = tuple(a + b for a, b in zip(ds["position"], ds["parkinglot"]["position_offset"]))
            for dc in ds["dumpers"]:
                for _ in range(dc["count"]):
                    # dump_site.py:84 multiplies the offset *list* by (count+1); zip keeps its first two entries
                    self.dumper_pos.append(tuple(a + b for a, b in zip(ds["position"], dc["position_offset"])))
        self.loc_name = [cfg["charging_site"]["name"]] + pc.load_site_names + pc.dump_site_names
        self.loc_pos = [self.charging_pos] + [self.site_pos[n] for n in pc.load_site_names + pc.dump_site_names]

    # ------------------------------------------------------------------------------------------
    def _pools(self, inst):
        """Rebuild every truck's event pool (and the jam list) from the engine's event trace."""
        pc = self.pc
        trace, summ = self.res.trace_np(inst)
        n = len(trace)
        # number of valid records: the engine writes them contiguously; find the last non-empty one
        valid = np.nonzero((trace["time"] != 0) | (trace["kind"] != 0) | (trace["truck"] != 0) | (trace["a"] != 0))[0]
        n = int(valid[-1]) + 1 if len(valid) else 0
        pools = [_Pool() for _ in range(pc.T)]
        cur = [0] * pc.T                    # location code at the time of each record
        tgt_of = [None] * pc.T              # target location of the truck's latest order
        rand = []                           # (record time, event) of mine.random_event_pool in creation order
        open_move, open_wait, open_get = {}, {}, {}
        L = pc.L
        for r in trace[:n]:
            k, t, now = int(r["kind"]), int(r["truck"]), float(r["time"])
            if k == tr.K_JAM:
                rid = int(r["arg0"])
                if rid < L:
                    a, b = 0, 1 + rid
                elif rid < L + L * pc.D:
                    a, b = 1 + (rid - L) // pc.D, 1 + L + (rid - L) % pc.D
                else:
                    q = rid - L - L * pc.D
                    a, b = 1 + L + q // L, 1 + q % L
                rand.append((now, _Ev(now, "RoadEvent:jam", {"start": now, "pos": float(r["a"]), "end": float(r["b"]),
                                                             "a": a, "b": b})))
                continue
            if k == tr.K_ROAD_REPAIR:
                rand.append((now, _Ev(now, "RoadEvent:repair", None)))
                continue
            if k == tr.K_SHOVEL_MAINT_START and t < pc.T and t in open_get:
                open_get.pop(t).info["end_time"] = now      # loading ended before the maintenance hold (truck.py:352-370)
            if k > tr.K_ROAD_REPAIR or t >= pc.T:
                continue
            p = pools[t]
            if k == tr.K_INIT_ORDER:
                tgt_of[t] = 1 + int(r["arg0"])
                p.add(_Ev(now, "INIT ORDER", {"start_time": now}))
            elif k in (tr.K_MOVE_INIT, tr.K_MOVE_HAUL, tr.K_MOVE_UNHAUL):
                typ = {tr.K_MOVE_INIT: "init", tr.K_MOVE_HAUL: "haul", tr.K_MOVE_UNHAUL: "unhaul"}[k]
                tgt = 1 + int(r["arg0"]) if k != tr.K_MOVE_HAUL else 1 + L + int(r["arg0"])
                ev = _Ev(now, typ, {"start_time": now, "est_end_time": float(r["a"]), "end_time": None,
                                    "start_location": cur[t], "target_location": tgt})
                p.add(ev)
                open_move[t] = (ev, tgt)
            elif k in (tr.K_WAIT_SHOVEL, tr.K_WAIT_DUMPER):
                if t in open_move:                      # arrival: truck.py:211-217 fills end_time, sets the location
                    ev, tgt = open_move.pop(t)
                    ev.info["end_time"] = now
                    cur[t] = tgt
                typ = "wait shovel" if k == tr.K_WAIT_SHOVEL else "wait dumper"
                ev = _Ev(now, typ, {"start_time": now, "end_time": None, "unit": int(r["arg0"]),
                                    "queue_index": int(r["arg1"])})
                p.add(ev)
                open_wait[t] = ev
            elif k in (tr.K_GET_SHOVEL, tr.K_GET_DUMPER):
                if t in open_wait:
                    open_wait.pop(t).info["end_time"] = now
                typ = "get shovel" if k == tr.K_GET_SHOVEL else "get dumper"
                ev = _Ev(now, typ, {"start_time": now, "end_time": None, "unit": int(r["arg0"])})
                p.add(ev)
                open_get[t] = ev
            elif k in (tr.K_ORDER_HAUL, tr.K_ORDER_BACK):
                if t in open_get:                       # truck.py:352-354 / 431-433: end_time of the get record
                    open_get.pop(t).info["end_time"] = now
                tgt_of[t] = 1 + L + int(r["arg0"]) if k == tr.K_ORDER_HAUL else 1 + int(r["arg0"])
                p.add(_Ev(now, "ORDER", {"start_time": now}))
            elif k == tr.K_BREAKDOWN:
                start, rep = float(r["a"]), float(r["b"])
                ev = _Ev(start, "TruckEvent:breakdown", {"start_time": start, "end_time": start + rep, "repair_time": rep,
                                                         "start_location": cur[t], "target_location": tgt_of[t]})
                p.add(ev)
                rand.append((now, ev))                   # the same object goes into the global pool (truck.py:143-144)
            elif k == tr.K_UNREPAIRABLE:
                ev = _Ev(now, "TruckEvent:unrepairable", {"time": now})
                p.add(ev)
                rand.append((now, ev))
        final_loc = list(cur)       # trucks still on the road at the stop time keep end_time None
        return pools, rand, final_loc, summ

    # ------------------------------------------------------------------------------------------
    def _truck_state(self, pool, t, prev_pos, final_loc):
        """ticker.py:92-289 for one truck at minute t.  Returns (state, position)."""
        past = pool.past(t + 0.1)
        status, pos = None, prev_pos
        if len(past) == 0:
            return -1, np.array(self.charging_pos) + np.array([0.0, 0.0])
        last = past[-1]
        typ = last.type
        if len(past) == 2 and typ == "init":
            status = -2
            ratio = (t - last.info["start_time"]) / (last.info["est_end_time"] - last.info["start_time"])
            direction = np.array(self.loc_pos[last.info["target_location"]]) - np.array(self.charging_pos)
            pos = np.array(self.charging_pos) + ratio * direction
        if typ == "wait shovel":
            status = 1
            sp = self.shovel_pos[last.info["unit"]]
            pos = np.array(sp) - np.array([0.05, 0.0]) - (last.info["queue_index"] + 1) * np.array([0.035, 0.0])
        if typ == "get shovel":
            status = 2
            pos = np.array(self.shovel_pos[last.info["unit"]]) - np.array([0.05, 0.0])
        if typ in ("haul", "unhaul"):
            status = 3 if typ == "haul" else 0
            end = last.info["end_time"] if last.info["end_time"] is not None else last.info["est_end_time"]
            ratio = (t - last.info["start_time"]) / (end - last.info["start_time"])
            a = np.array(self.loc_pos[last.info["start_location"]])
            b = np.array(self.loc_pos[last.info["target_location"]])
            pos = a + ratio * (b - a)
        if typ == "wait dumper":
            status = 4
            pos = np.array(self.park_pos[self.loc_name[final_loc]])      # the truck's FINAL current_location (ticker.py:196)
        if typ == "get dumper":
            status = 5
            pos = np.array(self.dumper_pos[last.info["unit"]])
        if "breakdown" in typ:
            if last.info["start_time"] <= t <= last.info["end_time"]:
                status = 6
            # the reference's "repaired, back on the road" branch (ticker.py:248-266) reads fields of
            # past_events[-2] that only move records carry; reproduce it when they exist
            if last.info["end_time"] < t < past[-2].info["end_time"]:
                prev = past[-2]
                status = {"init": -2, "haul": 3, "unhaul": 0}.get(prev.type, status)
                rep = last.info["repair_time"]
                ratio = (t - rep - prev.info["start_time"]) / (prev.info["end_time"] - rep - prev.info["start_time"])
                a = np.array(self.loc_pos[last.info["start_location"]])
                b = np.array(self.loc_pos[last.info["target_location"]])
                pos = np.array(self.charging_pos) + ratio * (b - a)
        if typ == "ORDER":
            status = 8
            pos = self.loc_pos[final_loc]
        if "unrepairable" in typ:
            status = 7
        return status, pos

    # ------------------------------------------------------------------------------------------
    def ticks(self, inst=0, truck_states=True):
        """The reference's ticks dict for one instance (keys: int minute -> packet, plus "summary")."""
        pc, res = self.pc, self.res
        T = res.tick_capacity
        kp = res.ticks[inst].cpu().numpy().view(TICK_DTYPE).reshape(-1)
        summ = res.summary_np()[inst]
        n_ticks = int(summ["n_ticks"])
        ns = pc.L + pc.D
        sq = res.site_queue_ticks[inst].cpu().numpy() if res.site_queue_ticks is not None else np.zeros((T, ns), np.int16)
        st = res.site_tons_ticks[inst].cpu().numpy() if res.site_tons_ticks is not None else np.zeros((T, ns))
        ut = res.unit_ticks[inst].cpu().numpy().view(UNIT_TICK_DTYPE).reshape(T, -1) if res.unit_ticks is not None else None
        ubeg = pc.site_unit_begin
        pools = rand = None
        if truck_states and res.trace is not None:
            pools, rand, final_loc, _ = self._pools(inst)
        # mine.random_event_pool replayed in creation order: its per-type list gives the reference's jam order
        # (bisect_left on bumped stamps, event.py:40-42) and its timestamp-keyed dict gives the reference's
        # event counters (mine.py:91-101), which under-count simultaneous events
        rpool, rptr = _Pool(), 0
        jam_list_done = False
        if rand is not None:
            full = _Pool()
            for _, ev in rand:
                full.add(_Ev(ev.time_stamp, ev.type, ev.info))
            jams = [e.info for e in full.by_type.get("RoadEvent:jam", [])]
        else:
            jams = None
        prev_pos = [None] * pc.T
        out = {}
        for t in range(min(T, n_ticks)):
            packet = {"tick": t}
            ts = {}
            if pools is not None:
                for i in range(pc.T):
                    try:
                        state, pos = self._truck_state(pools[i], t, prev_pos[i], final_loc[i])
                    except (TypeError, KeyError, IndexError, ZeroDivisionError):
                        state, pos = None, prev_pos[i]
                    prev_pos[i] = pos
                    ts[pc.truck_names[i]] = {"name": pc.truck_names[i], "time": t, "state": state,
                                             "position": [float(x) for x in (pos if pos is not None else (0.5, 0.5))]}
            packet["truck_states"] = ts
            ls_states, sh_states, ds_states, du_states = {}, {}, {}, {}
            for s, name in enumerate(pc.load_site_names):
                units = range(ubeg[s], ubeg[s + 1])
                svc = int(sum(int(ut[t, u]["service_count"]) for u in units)) if ut is not None else 0
                rec = {"name": name, "tons": _num(st[t, s]), "service_count": svc, "time": t,
                       "total_queue_length": int(sq[t, s]), "shovel_queue_length": {}, "position": list(self.site_pos[name])}
                for u in units:
                    uname = pc.shovel_names[u]
                    ql = int(ut[t, u]["queue_length"]) if ut is not None else 0
                    rec["shovel_queue_length"][uname] = ql
                    sh_states[uname] = {"name": uname, "repair": bool(ut[t, u]["repair"]) if ut is not None else False,
                                        "tons": _num(ut[t, u]["tons"]) if ut is not None else 0,
                                        "service_count": int(ut[t, u]["service_count"]) if ut is not None else 0,
                                        "time": t, "queue_length": ql, "position": tuple(self.shovel_pos[u])}
                ls_states[name] = rec
            for j, name in enumerate(pc.dump_site_names):
                s = pc.L + j
                units = range(ubeg[s], ubeg[s + 1])
                svc = int(sum(int(ut[t, u]["service_count"]) for u in units)) if ut is not None else 0
                rec = {"name": name, "tons": _num(st[t, s]), "service_count": svc, "time": t,
                       "total_queue_length": int(sq[t, s]), "dumper_queue_length": {}, "position": list(self.site_pos[name])}
                for u in units:
                    uname = pc.dumper_names[u - pc.S]
                    ql = int(ut[t, u]["queue_length"]) if ut is not None else 0
                    rec["dumper_queue_length"][uname] = ql
                    du_states[uname] = {"name": uname, "tons": _num(ut[t, u]["tons"]) if ut is not None else 0.0,
                                        "service_count": int(ut[t, u]["service_count"]) if ut is not None else 0,
                                        "time": t, "queue_length": ql, "position": list(self.dumper_pos[u - pc.S])}
                ds_states[name] = rec
            packet["load_site_states"], packet["dump_site_states"] = ls_states, ds_states
            packet["shovel_states"], packet["dumper_states"] = sh_states, du_states
            jam_pos, jam_last = [], []
            for jm in (jams or []):
                if jm["start"] <= t < jm["end"]:
                    a, b = np.array(self.loc_pos[jm["a"]]), np.array(self.loc_pos[jm["b"]])
                    jam_pos.append((a + (b - a) * jm["pos"]).tolist())
                    jam_last.append(jm["end"] - t)
            k = kp[t]
            n_j, n_r, n_b, n_u = int(k["n_jams"]), int(k["n_road_repairs"]), int(k["n_truck_repairs"]), int(k["n_unrepairable"])
            if rand is not None and self.reference_counters:
                while rptr < len(rand) and rand[rptr][0] < t:
                    ev = rand[rptr][1]
                    rpool.add(_Ev(ev.time_stamp, ev.type, None))
                    rptr += 1
                types = [e.type for e in rpool.event_set.values()]
                n_j, n_r = types.count("RoadEvent:jam"), types.count("RoadEvent:repair")
                n_b, n_u = types.count("TruckEvent:breakdown"), types.count("TruckEvent:unrepairable")
            packet["mine_states"] = {
                "time": t, "produced_tons": _num(k["produced_tons"]), "service_count": int(k["service_count"]),
                "working_truck_count": int(k["working"]), "waiting_truck_count": int(k["waiting"]),
                "load_unload_truck_count": int(k["load_unload"]), "moving_truck_count": int(k["moving"]),
                "repairing_truck_count": int(k["repairing"]),
                "jams": {"name": "jams", "time": t, "position": jam_pos, "last_times": jam_last},
                "road_jam_count": n_j, "road_repair_count": n_r, "truck_repair": n_b, "truck_unrepairable": n_u,
                "random_event_count": n_j + n_r + n_b + n_u}
            out[t] = packet
        last_t = min(T, n_ticks) - 1
        out["summary"] = self._summary(inst, out[last_t]["mine_states"] if last_t >= 0 else None, pools, summ)
        return out

    def _summary(self, inst, last_mine, pools, summ):
        pc = self.pc
        sa = self.mine.state_arrays(inst)
        cyc = sa["t_cycle"]
        avg = np.mean(cyc)
        # MatchingFactor (ticker.py:403-431; Burt & Caccetta heterogeneous-fleet match factor)
        lt = np.zeros((pc.T, pc.S))
        for i in range(pc.T):
            for j in range(pc.S):
                lt[i, j] = round((self.pc.raw_caps[i] / pc.shovel_tons[j]), 1) * pc.shovel_cycle[j]
        lt = np.unique(lt, axis=0)
        ult = np.unique(lt, axis=1).astype(int)
        cnt = {}
        for i in range(ult.shape[0]):
            row = np.array(lt[i, :]).astype(int)
            for v in set(row):
                cnt["%d_%d" % (i, v)] = list(row).count(v)
        lcm = np.lcm.reduce(ult, axis=1)
        up = 0
        for i in range(ult.shape[0]):
            for j in range(ult.shape[1]):
                up += cnt["%d_%d" % (i, ult[i, j])] * (lcm[i] / ult[i, j])
        mf = (pc.T * np.sum(lcm)) / (up * avg) if avg else 0
        total_wait = 0
        if pools is not None:
            waits = []
            for p in pools:
                # Truck.get_wait_time (truck.py:470-494), including the appended "end" record and bumped stamps
                w = p.by_type.get("wait shovel", [])
                g = list(p.by_type.get("get shovel", []))
                end = _Ev(pc.sim_time, "end", None)
                p.add(end)
                if len(w) > len(g):
                    g.append(p.by_type["end"][0])
                wt = (py_sum([e.time_stamp for e in g]) - py_sum([e.time_stamp for e in w])) / len(w) if len(w) else 0
                if wt == float("inf"):
                    wt = 0
                waits.append(wt)
            total_wait = py_sum(waits) if all(isinstance(x, float) for x in waits) else sum(waits)
        orders = {"init": [], "haul": [], "back": []}
        if self.res.orders is not None:
            row = self.res.orders_np()[inst]
            sptf = self.dispatcher_name == "SPTFDispatcher"
            for code in row[row >= 0]:
                kind, site = int(code) >> 12, int(code) & 0xFFF
                if kind == 0:
                    orders["init"].append(site)
                elif kind == 1:
                    orders["haul"].append(site)
                else:
                    if sptf:
                        orders["init"].append(site)          # sptf_dispatcher.py:59-61 logs the back order twice
                    orders["back"].append(site)
        return {"produced_tons": last_mine["produced_tons"] if last_mine else 0, "MatchingFactor": float(mf),
                "TotalWaitTime": total_wait, "avg_time_per_order": 0, "total_order_count": int(summ["n_wrapper_orders"]),
                "RoadJams": last_mine["road_jam_count"] if last_mine else 0,
                "init_orders": orders["init"], "haul_orders": orders["haul"], "back_orders": orders["back"],
                "load_sites_name": list(pc.load_site_names), "dump_sites_name": list(pc.dump_site_names)}

    # ------------------------------------------------------------------------------------------
    def states_dict(self, ticks):
        """What run_simulation feeds to Charter.draw (cli/run.py:248-267)."""
        keys = [k for k in ticks if k != "summary"]
        ms = [ticks[k]["mine_states"] for k in keys]
        return {"times": [m["time"] for m in ms], "produced_tons_list": [m["produced_tons"] for m in ms],
                "service_count_list": [m["service_count"] for m in ms],
                "waiting_truck_count_list": [m["waiting_truck_count"] for m in ms], "summary": ticks["summary"]}

    def write_to_file(self, ticks, result_dir="results", dispatcher_name=None, file_name=None):
        os.makedirs(result_dir, exist_ok=True)
        if file_name is None:
            ts = time.strftime("%Y-%m-%d %H-%M-%S", time.localtime())
            file_name = "MINE-%s-ALGO-%s-TIME-%s.json" % (self.pc.mine_name, dispatcher_name or self.dispatcher_name, ts)
        path = os.path.join(result_dir, file_name)
        with open(path, "w") as fh:
            json.dump(ticks, fh)
        return path


def _num(x):
    """JSON number as the reference would hold it: Python float (ints stay ints only where the reference has ints)."""
    x = float(x)
    return x
