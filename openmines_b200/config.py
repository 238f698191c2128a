"""Mine JSON -> packed structure-of-arrays config for the device engine.

Mirrors what ``run_dispatch_sim`` builds from the JSON (reference cli/run.py:48-107) and the
static quantities the reference recomputes at run time, evaluated here with the reference's own
arithmetic (same association order, IEEE doubles) so the device reads bit-identical constants:

* ``load_time[type][shovel] = (capacity / shovel_tons) * cycle_time``      (truck.py:223)
* ``load_site_productivity = sum(tons / cycle)``                            (load_site.py:182-184)
* ``sptf_cost[type][site] = capacity / load_site_productivity``             (sptf_dispatcher.py:35)
* ``jam_chance[n] = 1 - exp(-0.05 n)``                                      (road.py:164-167)
* Nearest tables                                                            (nearest_dispatcher.py:24-26,42-45,64-66)
* FixedGroup truck -> load-site table                                       (fixed_group_dispatcher.py:21-56)
* road distances in monitor order, ``d2l`` indexed ``[load][dump]``         (road.py:311-323, mine.py:138-154)

``sum_mode``: CPython >= 3.12 ``sum()`` over floats is Neumaier-compensated; earlier versions add
left to right.  The reference's results depend on it; default "neumaier".
"""
import ctypes
import json
import math

import numpy as np

OM_ABI_VERSION = 9
OM_CFG_RL_SURFACE = 1    # OmConfig.flags: allocate the fields om_step needs (bigger blobs, fewer resident instances)

ROAD_EVENT_DEFAULTS = {            # road.py:37-43
    "lambda_repair": 1 / (60 * 2),
    "mu_repair_duration": 60,
    "sigma_repair_duration": 10,
    "mu_punish_distance": 0.4,
    "sigma_punish_distance": 0.1,
}


def py_sum(items, mode="neumaier"):
    """``sum()`` over Python floats: Neumaier (CPython >= 3.12) or naive left-to-right."""
    if mode == "naive":
        s = 0.0
        for x in items:
            s = s + float(x)
        return s
    s, c = 0.0, 0.0
    for x in items:
        x = float(x)
        t = s + x
        if abs(s) >= abs(x):
            c += (s - t) + x
        else:
            c += (x - t) + s
        s = t
    if c and math.isfinite(c):
        s += c
    return s


class OmConfigStruct(ctypes.Structure):
    """ctypes mirror of ``OmConfig`` (include/openmines_b200.h)."""
    _dp = ctypes.POINTER(ctypes.c_double)
    _ip = ctypes.POINTER(ctypes.c_int32)
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("n_trucks", ctypes.c_int32), ("n_load_sites", ctypes.c_int32),
        ("n_dump_sites", ctypes.c_int32), ("n_shovels", ctypes.c_int32), ("n_dumpers", ctypes.c_int32),
        ("n_truck_types", ctypes.c_int32), ("jam_capacity", ctypes.c_int32), ("sum_mode", ctypes.c_int32),
        ("nearest_init", ctypes.c_int32), ("flags", ctypes.c_int32), ("reserved1", ctypes.c_int32),
        ("truck_capacity", _dp), ("truck_speed", _dp), ("truck_type", _ip), ("site_unit_begin", _ip),
        ("dumper_cycle", _dp), ("road_distance", _dp), ("load_time", _dp), ("sptf_cost", _dp),
        ("jam_chance", _dp), ("nearest_haul", _ip), ("nearest_back", _ip), ("fixed_group", _ip),
        ("shovel_rate", _dp), ("site_productivity", _dp),
        ("road_lambda", ctypes.c_double), ("road_mu", ctypes.c_double), ("road_sigma", ctypes.c_double),
        ("detour_mu", ctypes.c_double), ("detour_sigma", ctypes.c_double), ("breakdown_mean", ctypes.c_double),
        ("repair_mu", ctypes.c_double), ("repair_sigma", ctypes.c_double), ("unrepairable_mean", ctypes.c_double),
        ("maint_mean", ctypes.c_double), ("maint_mu", ctypes.c_double), ("maint_sigma", ctypes.c_double),
        ("jam_sd", ctypes.c_double), ("jam_scale", ctypes.c_double), ("sim_time", ctypes.c_double),
    ]


class PackedConfig:
    """Host-side packed mine config (numpy arrays) + names for reporting."""

    def c_struct(self):
        s = OmConfigStruct()
        s.abi_version = OM_ABI_VERSION
        s.n_trucks, s.n_load_sites, s.n_dump_sites = self.T, self.L, self.D
        s.n_shovels, s.n_dumpers, s.n_truck_types = self.S, self.M, self.n_types
        s.jam_capacity = self.jam_capacity
        s.sum_mode = 1 if self.sum_mode == "neumaier" else 0
        s.nearest_init = int(self.nearest_init)
        s.flags = int(getattr(self, "flags", 0))
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
        for name in ("truck_capacity", "truck_speed", "dumper_cycle", "road_distance", "load_time", "sptf_cost",
                     "jam_chance", "shovel_rate", "site_productivity"):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float64)
            setattr(self, name, a)
            setattr(s, name, a.ctypes.data_as(dp))
        for name in ("truck_type", "site_unit_begin", "nearest_haul", "nearest_back", "fixed_group"):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.int32)
            setattr(self, name, a)
            setattr(s, name, a.ctypes.data_as(ip))
        p = self.road_params
        s.road_lambda, s.road_mu, s.road_sigma = p["lambda_repair"], p["mu_repair_duration"], p["sigma_repair_duration"]
        s.detour_mu, s.detour_sigma = p["mu_punish_distance"], p["sigma_punish_distance"]
        s.breakdown_mean, s.repair_mu, s.repair_sigma = 60 * 6, 10, 3           # truck.py:66,68,69
        s.unrepairable_mean = 60 * 24 * 7 * 7                                    # truck.py:67
        s.maint_mean, s.maint_mu, s.maint_sigma = 60 * 24, 45, 5                 # truck.py:358,362
        s.jam_sd, s.jam_scale = 0.1, 10                                          # road.py:185,212
        s.sim_time = float(self.sim_time)
        return s


CONF_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "conf")


def builtin_config(name="north_pit_mine"):
    """A mine config shipped with the package (``openmines_b200/conf``: the reference's ``conf/north_pit_mine.json`` and
    ``north_pit_mine_short.json``, re-serialised)."""
    import os
    with open(os.path.join(CONF_DIR, name + ".json"), "r", encoding="utf-8") as fh:
        return json.load(fh)


def load_config(path_or_dict):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    with open(path_or_dict, "r", encoding="utf-8") as fh:
        return json.load(fh)


def pack_config(config, jam_capacity=64, sum_mode="neumaier"):
    cfg = load_config(config)
    pc = PackedConfig()
    pc.raw = cfg
    pc.sum_mode = sum_mode
    pc.jam_capacity = int(jam_capacity)
    pc.mine_name = cfg["mine"]["name"]
    pc.sim_time = cfg["sim_time"]

    # trucks (run.py:57-64)
    names, caps, speeds = [], [], []
    for tc in cfg["charging_site"]["trucks"]:
        for i in range(tc["count"]):
            names.append("%s%d" % (tc["type"], i + 1))
            caps.append(tc["capacity"])
            speeds.append(tc["speed"])
    if len(set(names)) != len(names):
        raise NotImplementedError("duplicate truck names: SPTF/SQ exclude trucks by name "
                                  "(sptf_dispatcher.py:23); the device engine assumes unique names")
    pc.truck_names = names
    pc.raw_caps = list(caps)
    pc.T = len(names)
    type_keys = []
    ttype = []
    for cap in caps:
        if cap not in type_keys:
            type_keys.append(cap)
        ttype.append(type_keys.index(cap))
    pc.n_types = len(type_keys)
    pc.truck_capacity = np.array(caps, dtype=np.float64)
    pc.truck_speed = np.array(speeds, dtype=np.float64)
    pc.truck_type = np.array(ttype, dtype=np.int32)

    # load sites / shovels (run.py:67-79), dump sites / dumpers (run.py:82-94)
    pc.load_site_names = [ls["name"] for ls in cfg["load_sites"]]
    pc.dump_site_names = [ds["name"] for ds in cfg["dump_sites"]]
    pc.L, pc.D = len(pc.load_site_names), len(pc.dump_site_names)
    ubeg = [0]
    sh_tons, sh_cycle, sh_names = [], [], []
    for ls in cfg["load_sites"]:
        for sc in ls["shovels"]:
            sh_tons.append(sc["tons"])
            sh_cycle.append(sc["cycle_time"])
            sh_names.append(sc["name"])
        ubeg.append(len(sh_tons))
    pc.S = len(sh_tons)
    du_cycle, du_names = [], []
    for ds in cfg["dump_sites"]:
        for dc in ds["dumpers"]:
            for i in range(dc["count"]):
                du_cycle.append(dc["cycle_time"])
                du_names.append("%s-点位%d" % (ds["name"], i))
        ubeg.append(pc.S + len(du_cycle))
    pc.M = len(du_cycle)
    pc.shovel_names, pc.dumper_names = sh_names, du_names
    pc.shovel_tons, pc.shovel_cycle = sh_tons, sh_cycle
    pc.site_unit_begin = np.array(ubeg, dtype=np.int32)
    pc.dumper_cycle = np.array(du_cycle, dtype=np.float64)

    # road (run.py:97-104, road.py:20-58)
    road = cfg["road"]
    l2d = np.array(road["l2d_road_matrix"])
    d2l = np.array(road["d2l_road_matrix"])
    c2l = list(road["charging_to_load_road_matrix"])
    L, D = pc.L, pc.D
    if l2d.shape != (L, D) or d2l.shape != (L, D) or len(c2l) != L:
        raise ValueError("road matrices must be [n_load_sites][n_dump_sites] (road.py:314-320)")
    rd = [float(c2l[i]) for i in range(L)]
    rd += [float(l2d[i][j]) for i in range(L) for j in range(D)]
    rd += [float(d2l[i][j]) for j in range(D) for i in range(L)]       # dump j -> load i uses d2l[i][j]
    pc.road_distance = np.array(rd, dtype=np.float64)
    pc.R = len(rd)
    params = dict(ROAD_EVENT_DEFAULTS)
    params.update(road.get("road_event_params", {}) or {})
    pc.road_params = params

    # static tables
    lt = np.zeros((pc.n_types, pc.S), dtype=np.float64)
    for k, cap in enumerate(type_keys):
        for u in range(pc.S):
            lt[k, u] = (cap / sh_tons[u]) * sh_cycle[u]
    pc.load_time = lt.reshape(-1)
    prod = []
    for s in range(L):
        prod.append(py_sum([sh_tons[u] / sh_cycle[u] for u in range(ubeg[s], ubeg[s + 1])], sum_mode))
    pc.load_site_productivity = prod
    pc.site_productivity = np.array(prod, dtype=np.float64)
    pc.shovel_rate = np.array([sh_tons[u] / sh_cycle[u] for u in range(pc.S)], dtype=np.float64)
    sc = np.zeros((pc.n_types, L), dtype=np.float64)
    for k, cap in enumerate(type_keys):
        for s in range(L):
            sc[k, s] = cap / prod[s]
    pc.sptf_cost = sc.reshape(-1)
    pc.jam_chance = np.array([1.0 - math.exp(-0.05 * n) for n in range(pc.T + 1)], dtype=np.float64)
    pc.nearest_init = c2l.index(min(c2l))
    pc.nearest_haul = np.array([int(l2d[i, :].argmin()) for i in range(L)], dtype=np.int32)
    pc.nearest_back = np.array([int(d2l[:, j].argmin()) for j in range(D)], dtype=np.int32)

    # FixedGroup (fixed_group_dispatcher.py:21-56); the position-based "dumpsite" of :58-65 is unused
    total_prod = py_sum(prod, sum_mode)
    total_cap = sum(caps)
    quota = [(p / total_prod) * total_cap for p in prod]
    assigned = [0 for _ in range(L)]
    fg = np.zeros(pc.T, dtype=np.int32)
    for t, cap in enumerate(caps):
        for s in range(L):
            if assigned[s] < quota[s]:
                assigned[s] += cap
                fg[t] = s
                break
    pc.fixed_group = fg
    return pc
