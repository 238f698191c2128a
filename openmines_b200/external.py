"""User-defined dispatchers on the batched engine (external-decision mode).

The reference's plug-in boundary is ``BaseDispatcher.give_init_order / give_haul_order /
give_back_order(truck, mine) -> int`` (openmines/src/dispatcher.py:91-98, README.md:63-87), called
synchronously from ``Truck.run`` (truck.py:302,389,441) after the wrapper refreshed the mine's derived
state (dispatcher.py:18-33).  The six built-ins are device functions; any OTHER subclass is ordinary
Python that cannot run on the GPU, so it runs here: instances are reset with ``OM_DISP_EXTERNAL``, the
engine pauses each one at its next order (``om_step``), the host builds read-only ``truck`` / ``mine``
views over the instance's state blob -- the attributes README section 6 documents for dispatcher
authors -- calls the user's method, and feeds the returned site index back as the step's action.

The simulation itself (event ordering, queues, stochastic events) still runs on the GPU and is the same
code path as the batched RL step; only the decision is taken on the host.  Error behaviour follows the
reference: exceptions raised by the dispatcher propagate, an out-of-range index raises ``IndexError``
(truck.py:305,390,442).
"""
import copy
import time

import numpy as np

from .config import py_sum
from .engine import BatchedMine, HDR_DTYPE

STATUS_NAMES = ["idle", "waiting for init order", "moving", "waiting for shovel", "loading",
                "waiting for haul order", "waiting for dumper", "unloading", "waiting for back order",
                "repairing", "unrepairable"]                      # truck.py:72-436
_NOLOC = 0xFF
_NONE16 = 0xFFFF


class _View:
    """Plain attribute container with a readable repr."""

    def __repr__(self):
        return "<%s %s>" % (type(self).__name__, getattr(self, "name", ""))


class TruckView(_View):
    pass


class ShovelView(_View):
    pass


class DumperView(_View):
    pass


class ParkingLotView(_View):
    pass


class LoadSiteView(_View):
    pass


class DumpSiteView(_View):
    pass


class ChargingSiteView(_View):
    pass


class EnvView(_View):
    pass


class RoadView(_View):
    def truck_on_road(self, start, end):
        """road.py:66-80: trucks whose (current, target) pair is (start, end) and whose status is "moving"."""
        return [t for t in self.mine.trucks
                if t.current_location is not None and t.target_location is not None
                and t.current_location.name == start.name and t.target_location.name == end.name
                and t.status == "moving"]


class MineView(_View):
    pass


def _queue_of(sa, u):
    """FIFO of waiting trucks of unit u (simpy Resource.put_queue), oldest first."""
    out, t = [], int(sa["u_head"][u])
    for _ in range(int(sa["u_qlen"][u])):
        out.append(t)
        t = int(sa["t_next"][t])
    return out


def build_mine_view(mine, inst, state_host=None, dispatcher=None):
    """Read-only snapshot of instance ``inst`` as the object graph a dispatcher sees (README.md section 6).

    Derived values are recomputed with the reference's arithmetic, as its wrapper does before every order:
    ``update_queue_wait_status`` (load_site.py:21-68), site productivity / service ability (load_site.py:175-187,
    dump_site.py:69-78)."""
    pc = mine.pc
    sa = mine.state_arrays(inst, blob=None if state_host is None else state_host[inst])
    hdr = sa["hdr"]
    now = float(hdr["now"])
    L, D, S = pc.L, pc.D, pc.S
    ub = pc.site_unit_begin
    mv = MineView()
    mv.name = pc.mine_name
    mv.env = EnvView()
    mv.env.now = now
    mv.dispatcher = dispatcher
    mv.produced_tons = float(hdr["reported_tons"])          # mine.produce_tons: the minute-sampled value (mine.py:61)
    mv.produce_tons = mv.produced_tons
    # ---- sites -----------------------------------------------------------------------------------
    load_sites, dump_sites = [], []
    for s in range(L):
        ls = LoadSiteView()
        ls.name = pc.load_site_names[s]
        ls.position = pc.raw["load_sites"][s].get("position")
        ls.shovel_list = []
        for u in range(ub[s], ub[s + 1]):
            sh = ShovelView()
            sh.name, sh.load_site = pc.shovel_names[u], ls
            sh.shovel_tons, sh.shovel_cycle_time = pc.shovel_tons[u], pc.shovel_cycle[u]
            sh.service_count = int(sa["u_svc"][u])
            sh.produced_tons = float(sa["u_tons"][u])
            sh.last_service_time = float(sa["u_last_start"][u])
            sh.last_service_done_time = float(sa["u_last_done"][u])
            sh.last_breakdown_time = float(sa["u_lastbrk"][u])
            sh.repair = bool(sa["u_repair"][u])
            queue = _queue_of(sa, u)
            user = int(sa["u_user"][u])
            sh.queue_length = len(queue)
            sh.users = [] if user == _NONE16 else [user]
            # load_site.py:30-47: max(last start, last done) + sum of load times over queue then users - now, floored at 0
            last = max(sh.last_service_time, sh.last_service_done_time)
            total = 0.0
            for k in queue + sh.users:
                total += float(pc.load_time[pc.truck_type[k] * S + u])
            sh.est_waiting_time = max((last + total) - now, 0)
            ls.shovel_list.append(sh)
        ests = [sh.est_waiting_time for sh in ls.shovel_list]
        ls.estimated_queue_wait_time = min(ests)
        ls.avg_queue_wait_time = sum(ests) / len(ests)
        ls.load_site_productivity = pc.load_site_productivity[s]
        ls.service_ability_ratio = py_sum([pc.shovel_rate[u] * (0.0 if sa["u_repair"][u] else 1.0)
                                           for u in range(ub[s], ub[s + 1])], pc.sum_mode) / ls.load_site_productivity
        ls.produced_tons = py_sum([float(sa["u_tons"][u]) for u in range(ub[s], ub[s + 1])], pc.sum_mode)
        ls.service_count = int(sum(int(sa["u_svc"][u]) for u in range(ub[s], ub[s + 1])))
        pl = ParkingLotView()
        pl.name = ls.name + "-ParkingLot"
        total_q = int(sa["s_qtotal"][s])
        pl.queue_status = {"total": {"cur_value": total_q, int(now): total_q}}       # minute-sampled (load_site.py:81-83)
        pl.queue_lengths = [sh.queue_length for sh in ls.shovel_list]
        ls.parking_lot = ls.parkinglot = pl
        load_sites.append(ls)
    for j in range(D):
        ds = DumpSiteView()
        ds.name = pc.dump_site_names[j]
        ds.position = pc.raw["dump_sites"][j].get("position")
        ds.dumper_list = []
        for u in range(ub[L + j], ub[L + j + 1]):
            du = DumperView()
            du.name, du.dump_site = pc.dumper_names[u - S], ds
            du.dump_time = float(pc.dumper_cycle[u - S])
            du.dumper_tons = float(sa["u_tons"][u])
            du.service_count = int(sa["u_svc"][u])
            du.last_service_time = float(sa["u_last_start"][u])
            du.last_service_done_time = float(sa["u_last_done"][u])
            du.queue_length = int(sa["u_qlen"][u])
            # load_site.py:49-60: dump_time * len(queue) minus the time already used by the current unload
            total = 0.0
            for _ in range(du.queue_length):
                total += du.dump_time
            used = now - du.last_service_time
            du.est_waiting_time = (total - used) if total > used else 0
            ds.dumper_list.append(du)
        ests = [du.est_waiting_time for du in ds.dumper_list]
        ds.estimated_queue_wait_time = min(ests)
        ds.avg_queue_wait_time = sum(ests) / len(ests)
        ds.produce_tons = py_sum([du.dumper_tons for du in ds.dumper_list], pc.sum_mode)
        ds.tons = ds.produce_tons
        ds.service_count = int(sum(du.service_count for du in ds.dumper_list))
        ds.truck_visits = ds.service_count
        ds.service_ability_ratio = 1
        pl = ParkingLotView()
        pl.name = ds.name + "-ParkingLot"
        total_q = int(sa["s_qtotal"][L + j])
        pl.queue_status = {"total": {"cur_value": total_q, int(now): total_q}}
        pl.queue_lengths = [du.queue_length for du in ds.dumper_list]
        ds.parking_lot = ds.parkinglot = pl
        dump_sites.append(ds)
    mv.load_sites, mv.dump_sites = load_sites, dump_sites
    mv.service_count = int(sum(ds.service_count for ds in dump_sites))
    cs = ChargingSiteView()
    cs.name = pc.raw["charging_site"]["name"]
    cs.position = pc.raw["charging_site"].get("position")
    mv.charging_site = cs
    locations = [cs] + load_sites + dump_sites                       # location codes 0, 1..L, L+1..L+D
    # ---- trucks ----------------------------------------------------------------------------------
    trucks = []
    for t in range(pc.T):
        tv = TruckView()
        tv.name = pc.truck_names[t]
        tv.truck_capacity, tv.truck_speed = pc.raw_caps[t], float(pc.truck_speed[t])
        cur, tgt = int(sa["t_cur"][t]), int(sa["t_tgt"][t])
        tv.current_location = None if cur == _NOLOC else locations[cur]
        tv.target_location = None if tgt == _NOLOC else locations[tgt]
        tv.status = STATUS_NAMES[int(sa["t_status"][t])]
        tv.truck_load = float(sa["t_load"][t])
        tv.service_count = int(sa["t_svc"][t])
        tv.total_load_count = float(sa["t_total"][t])
        tv.truck_cycle_time = float(sa["t_cycle"][t])
        tv.first_order_time = 0
        tv.journey_start_time = float(sa["t_jstart"][t])
        tv.pre_jam_time = float(sa["t_prejam"][t])
        tv.last_breakdown_time = float(sa["t_lastbrk"][t])
        tv.repair_time = float(sa["t_repair"][t])
        tv.mine, tv.env, tv.dispatcher = mv, mv.env, dispatcher
        trucks.append(tv)
    mv.trucks = cs.trucks = trucks
    # ---- road ------------------------------------------------------------------------------------
    rv = RoadView()
    rv.mine = mv
    road = pc.raw["road"]
    rv.l2d_road_matrix = np.array(road["l2d_road_matrix"])
    rv.d2l_road_matrix = np.array(road["d2l_road_matrix"])
    rv.charging_to_load_road_matrix = list(road["charging_to_load_road_matrix"])
    rv.charging_to_load = rv.charging_to_load_road_matrix            # attribute name used by road.py:33 / nearest_dispatcher.py:24
    rv.road_matrix = rv.l2d_road_matrix
    rv.load_site_num, rv.dump_site_num = L, D
    repairs = {}
    ends = sa["r_repair_end"]
    for i in range(L):
        for j in range(D):
            for (a, b, rid) in ((load_sites[i], dump_sites[j], L + i * D + j),
                                (dump_sites[j], load_sites[i], L + L * D + j * L + i)):
                e = float(ends[rid])
                repairs[(a, b)] = (now < e, e)                        # road.py:235-282 (in_repair, repair_end)
    rv.road_repairs = repairs
    mv.road = rv
    return mv


_METHODS = ("give_init_order", "give_haul_order", "give_back_order")
_LISTS = ("init_orders", "haul_orders", "back_orders")


def run_external(dispatcher, config, n_instances=1, seeds=None, rng_mode="det", tapes=None, device=0, until=None,
                 jam_capacity=64):
    """Run a user-defined ``BaseDispatcher`` (instance, class or zero-argument factory) on ``n_instances`` mines.

    Returns ``{"mine": BatchedMine, "headers": per-instance InstHdr records, "dispatchers": [...], "produced_tons": [...]}``.
    Each instance gets its own dispatcher object, as the reference constructs a fresh one per run (cli/run.py:158,179)."""
    import torch
    n = int(n_instances)
    if isinstance(dispatcher, type) or (callable(dispatcher) and not hasattr(dispatcher, "give_init_order")):
        dispatchers = [dispatcher() for _ in range(n)]
    else:
        dispatchers = [dispatcher] + [copy.deepcopy(dispatcher) for _ in range(n - 1)]
    mine = BatchedMine(config, n, device=device, jam_capacity=jam_capacity, rl_surface=True)   # om_step needs it
    seeds = np.arange(n, dtype=np.uint64) if seeds is None else np.asarray(seeds, dtype=np.uint64)
    mine.reset("external", seeds=seeds, rng_mode=rng_mode, tapes=tapes)
    dev = mine.device
    io = {"obs": torch.zeros((n, mine.obs_dim()), dtype=torch.float32, device=dev),
          "reward": torch.zeros(n, dtype=torch.float64, device=dev),
          "done": torch.zeros(n, dtype=torch.uint8, device=dev),
          "order_kind": torch.full((n,), -1, dtype=torch.int32, device=dev),
          "truck": torch.full((n,), -1, dtype=torch.int32, device=dev),
          "info": torch.zeros((n, 4), dtype=torch.float64, device=dev)}
    actions = torch.zeros(n, dtype=torch.int32, device=dev)
    mine.step(None, io, until=until)
    while True:
        done = io["done"].cpu().numpy().astype(bool)
        if done.all():
            break
        kinds, trucks = io["order_kind"].cpu().numpy(), io["truck"].cpu().numpy()
        state_host = mine.state.cpu().numpy()
        acts = np.zeros(n, dtype=np.int32)
        for e in range(n):
            if done[e]:
                continue
            d, kind = dispatchers[e], int(kinds[e])
            view = build_mine_view(mine, e, state_host=state_host, dispatcher=d)
            t0 = time.perf_counter()
            a = getattr(d, _METHODS[kind])(view.trucks[int(trucks[e])], view)
            dt = (time.perf_counter() - t0) * 1000.0
            sites = view.dump_sites if kind == 1 else view.load_sites
            a = int(a)
            # the wrapper's bookkeeping (dispatcher.py:70-86) happens before truck.py indexes the site list
            getattr(d, _LISTS[kind]).append(a)
            cnt, tm = _LISTS[kind][:-1] + "_count", _LISTS[kind][:-1] + "_time"
            setattr(d, cnt, getattr(d, cnt, 0) + 1)
            setattr(d, tm, getattr(d, tm, 0) + dt)
            d.total_order_count = getattr(d, "total_order_count", 0) + 1
            d.total_order_time = getattr(d, "total_order_time", 0) + dt
            sites[a]                                                   # IndexError like truck.py:305,390,442
            if a < 0:
                a += len(sites)                                        # Python list indexing accepts negative indices
            acts[e] = a
        actions.copy_(torch.from_numpy(acts))
        mine.step(actions, io, until=until)
    headers = mine.headers()
    return {"mine": mine, "headers": headers, "dispatchers": dispatchers,
            "produced_tons": io["info"][:, 0].cpu().numpy().copy()}
