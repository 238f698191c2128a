"""``BatchedMine``: host-side driver of the device engine through the C ABI.

It plays the role of ``Mine`` + ``run_dispatch_sim`` (reference mine.py:340-395, cli/run.py:48-112)
for N independent mine instances at once.  PyTorch is used only to own device memory and
streams; every simulation step runs in the hand-written CUDA kernels of
``libopenmines_b200.so`` (there is no CPU or PyTorch fallback).
"""
import ctypes

import numpy as np

from . import _lib
from .config import PackedConfig, pack_config, OM_CFG_RL_SURFACE
from .trace import SUMMARY_DTYPE, TICK_DTYPE, TRACE_DTYPE

RNG_MODES = {"det": 0, "philox": 1, "tape": 2}

DISPATCHER_IDS = {
    "NaiveDispatcher": 0, "RandomDispatcher": 1, "NearestDispatcher": 2, "FixedGroupDispatcher": 3,
    "SPTFDispatcher": 4, "SQDispatcher": 5, "ShortestTripDispatcher": 6, "external": 255,
}


class OmOutputsStruct(ctypes.Structure):
    _fields_ = [("summary", ctypes.c_void_p), ("ticks", ctypes.c_void_p), ("tick_capacity", ctypes.c_int32),
                ("trace_capacity", ctypes.c_int32), ("trace", ctypes.c_void_p), ("site_queue_ticks", ctypes.c_void_p),
                ("site_tons_ticks", ctypes.c_void_p), ("orders", ctypes.c_void_p), ("order_capacity", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("unit_ticks", ctypes.c_void_p), ("draw_values", ctypes.c_void_p),
                ("draw_sites", ctypes.c_void_p), ("draw_count", ctypes.c_void_p), ("draw_capacity", ctypes.c_int64)]


class OmStepIOStruct(ctypes.Structure):
    _fields_ = [("actions", ctypes.c_void_p), ("obs", ctypes.c_void_p), ("reward", ctypes.c_void_p),
                ("done", ctypes.c_void_p), ("order_kind", ctypes.c_void_p), ("truck", ctypes.c_void_p),
                ("sug_action", ctypes.c_void_p), ("info", ctypes.c_void_p), ("teacher", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("final_obs", ctypes.c_void_p), ("seed_stride", ctypes.c_uint64),
                ("episode_totals", ctypes.c_void_p)]


OM_STEP_AUTO_RESET = 1
OM_STEP_SPARSE_REWARD = 2
DRAW_SKIP = 255      # OM_DRAW_SKIP: slot reserved in the draw tape for a draw the reference would not have made


_LAYOUT_FIELDS = ["total_bytes", "hdr", "t_time", "t_load", "t_jstart", "t_prejam", "t_lastbrk", "t_repair",
                  "t_total", "t_cycle", "t_dist", "t_since", "t_key", "t_rng", "t_svc", "t_epoch", "t_next", "t_unit", "t_phase",
                  "t_status", "t_cur", "t_tgt", "u_last_start", "u_last_done", "u_tons", "u_lastbrk",
                  "u_rel_key", "u_svc", "u_user", "u_head", "u_tail", "u_qlen", "u_repair", "r_repair_end",
                  "j_pos", "j_start", "j_end", "j_road", "s_qtotal", "s_coming", "s_tons", "s_cap", "s_ability", "scratch",
                  "t_rec", "u_rec", "t_stride", "u_stride", "e_stride", "hot_bytes", "hdr_bytes", "slot_cache", "t_aux", "r_jamcnt", "t_loc", "t_lastbrk_hot", "r_on", "r_head", "t_rnext", "t_rprev"]
_ENTRY_FIELDS = {"t_time", "t_key", "t_phase"}                # the truck's 16-byte pending entry
_SOA_FIELDS = {"u_rel_key", "u_lastbrk", "t_since"}            # every other t_* / u_* field lives in a record


# numpy mirror of the device-side InstHdr (csrc/om_engine.cuh): the first 192 bytes of every state blob
HDR_DTYPE = np.dtype([("now", "<f8"), ("tick_time", "<f8"), ("reported_tons", "<f8"), ("seed", "<u8"), ("tape_pos", "<u8"),
                      ("seq", "<u4"), ("tick_key", "<u4"), ("tick_k", "<i4"), ("dispatcher", "<i4"),
                      ("n_events", "<u4"), ("n_decisions", "<u4"), ("n_orders", "<u4"), ("n_hops", "<u4"),
                      ("n_jams", "<u4"), ("n_road_rep", "<u4"), ("n_brk", "<u4"), ("n_unrep", "<u4"),
                      ("jam_count", "<u4"), ("err", "<u4"), ("trace_count", "<u4"), ("order_count", "<u4"),
                      ("pending_truck", "<i4"), ("pending_kind", "<i4"), ("n_ticks", "<u4"), ("n_rel", "<u4"),
                      ("rl_last_time", "<f8"), ("rl_last_prod", "<f8"), ("acc_wait", "<f8"), ("acc_shovel", "<f8"),
                      ("acc_dumper", "<f8"), ("acc_move", "<f8"), ("acc_jam", "<f8"), ("finished", "<u4"),
                      ("rl_steps", "<u4"), ("pad", "V8")])
assert HDR_DTYPE.itemsize == 192


class OmStateLayoutStruct(ctypes.Structure):
    _fields_ = [(f, ctypes.c_uint32) for f in _LAYOUT_FIELDS]


def _dispatcher_id(d):
    if isinstance(d, (int, np.integer)):
        return int(d)
    if isinstance(d, str):
        return DISPATCHER_IDS[d]
    dev = getattr(d, "device_id", None)      # openmines_b200.dispatcher.BaseDispatcher subclasses
    if dev is not None:
        return int(dev)
    name = getattr(d, "name", d.__class__.__name__)
    if name in DISPATCHER_IDS:
        return DISPATCHER_IDS[name]
    raise ValueError("dispatcher %r has no device implementation; use the external-decision mode" % (d,))


class RunResult:
    """Device tensors produced by one ``BatchedMine.run`` call (numpy views via the helpers)."""

    def __init__(self):
        self.summary = self.ticks = self.trace = self.site_queue_ticks = self.site_tons_ticks = self.orders = None
        self.unit_ticks = None
        self.draw_values = self.draw_sites = self.draw_count = None
        self.tick_capacity = self.trace_capacity = self.order_capacity = self.draw_capacity = 0
        self.regrown = 0                 # instances re-simulated with a larger jam pool (run(regrow=True))
        self.regrown_instances = []

    def summary_np(self):
        return self.summary.cpu().numpy().view(SUMMARY_DTYPE).reshape(-1)

    def ticks_np(self):
        return self.ticks.cpu().numpy().view(TICK_DTYPE).reshape(-1, self.tick_capacity)

    def trace_np(self, inst):
        s = self.summary_np()
        tr = self.trace[inst].cpu().numpy().view(TRACE_DTYPE).reshape(-1)
        return tr, s[inst]

    def orders_np(self):
        return self.orders.cpu().numpy()

    def draw_tape(self, inst):
        """(sites uint8[], values float64[]) of instance ``inst``: every draw of the Philox run in the reference's call order
        (``run(draw_capacity=...)``), reserved-but-undrawn slots (OM_DRAW_SKIP) removed.  Feeding it back as a tape
        (``reset(rng_mode="tape", tapes=[...])``, or the reference under ``oracle/ref_harness.py``) replays the trajectory."""
        n = int(self.draw_count[inst].item())
        if n > self.draw_capacity:
            raise RuntimeError("draw tape of instance %d overflowed: %d slots needed, capacity %d" % (inst, n, self.draw_capacity))
        sites = self.draw_sites[inst, :n].cpu().numpy()
        values = self.draw_values[inst, :n].cpu().numpy()
        keep = sites != DRAW_SKIP
        return sites[keep].copy(), values[keep].copy()


class BatchedMine:
    """N independent instances of one mine config on one GPU."""

    def __init__(self, config, n_instances, device=0, jam_capacity=64, sum_mode="neumaier", rl_surface=False,
                 config_ids=None):
        """``rl_surface=True`` allocates the fields ``om_step`` needs (reward windows, minute-sampled site values):
        required by ``step()`` / ``BatchedMineEnv`` / ``run_external``; episode runs (``run()``) are faster without it
        because smaller state blobs leave room for more resident instances per SM.

        Mixed-config batch: ``config`` may be a LIST of mine configs (dicts / paths / ``PackedConfig``) with ``config_ids``
        naming, per instance, the one it simulates (``om_create_multi``) -- the reference's fleet-size and scene sweeps
        (cli/run.py:283-472) as one batch.  ``jam_capacity`` may then be a list too.  ``pc_of(inst)`` gives an instance's
        packed config; ``pc`` is the first one."""
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("openmines_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.multi = isinstance(config, (list, tuple))
        configs = list(config) if self.multi else [config]
        jams = list(jam_capacity) if isinstance(jam_capacity, (list, tuple)) else [jam_capacity] * len(configs)
        self.pcs = [c if isinstance(c, PackedConfig) else pack_config(c, j, sum_mode) for c, j in zip(configs, jams)]
        for pc in self.pcs:
            pc.flags = OM_CFG_RL_SURFACE if rl_surface else 0
        self.pc = self.pcs[0]
        self.rl_surface = bool(rl_surface)
        self.n = int(n_instances)
        self.device = torch.device("cuda", device)
        self._cstructs = [pc.c_struct() for pc in self.pcs]
        h = ctypes.c_void_p()
        if self.multi:
            if config_ids is None:
                raise ValueError("a list of configs needs config_ids (one config index per instance)")
            self.config_ids = np.ascontiguousarray(config_ids, dtype=np.int32)
            if self.config_ids.shape != (self.n,):
                raise ValueError("config_ids must have one entry per instance")
            arr = (ctypes.c_void_p * len(self.pcs))(*[ctypes.addressof(c) for c in self._cstructs])
            rc = self.lib.om_create_multi(arr, len(self.pcs), self.config_ids.ctypes.data_as(ctypes.c_void_p), self.n, device,
                                          ctypes.byref(h))
        else:
            self.config_ids = None
            rc = self.lib.om_create(ctypes.byref(self._cstructs[0]), self.n, device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("om_create failed (%d): %s" % (rc, self.lib.om_last_error(None).decode()))
        self.h = h
        # row widths of the per-site / per-unit outputs and the loop bounds of the helpers: maxima over the configs
        self.max_T = max(pc.T for pc in self.pcs)
        self.max_NS = max(pc.L + pc.D for pc in self.pcs)
        self.max_U = max(pc.S + pc.M for pc in self.pcs)
        self.max_R = max(pc.R for pc in self.pcs)
        sb = ctypes.c_size_t()
        self._ck(self.lib.om_state_bytes(self.h, ctypes.byref(sb)))
        self.state_bytes = int(sb.value)
        self.layout = OmStateLayoutStruct()
        self._ck(self.lib.om_state_layout(self.h, ctypes.byref(self.layout)))
        self.state = torch.zeros((self.n, self.state_bytes), dtype=torch.uint8, device=self.device)
        self.dispatcher_ids = None
        self.has_external = False
        self.seeds = None
        self._tape = None
        self.sim_time = max(pc.sim_time for pc in self.pcs)

    def pc_of(self, inst):
        """Packed config of instance ``inst``."""
        return self.pcs[int(self.config_ids[inst])] if self.multi else self.pc

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("openmines_b200 call failed (%d): %s" % (rc, self.lib.om_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.om_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def launch_info(self):
        g, b, s, w = (ctypes.c_int32() for _ in range(4))
        self._ck(self.lib.om_launch_info(self.h, ctypes.byref(g), ctypes.byref(b), ctypes.byref(s), ctypes.byref(w)))
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "warps_per_cta": w.value}

    # ------------------------------------------------------------------------------------------
    def reset(self, dispatchers, seeds=None, rng_mode="philox", tapes=None):
        """Write the t=0 state of every instance.

        dispatchers: one dispatcher (name / id / BaseDispatcher object) or a sequence of N of them.
        seeds:       uint64 per instance (default: instance index).
        tapes:       for rng_mode="tape": list of N (sites uint8[], values float64[]) tuples.
        """
        torch = self.torch
        if isinstance(dispatchers, (str, int, np.integer)) or not hasattr(dispatchers, "__len__"):
            ids = np.full(self.n, _dispatcher_id(dispatchers), dtype=np.int32)
        elif isinstance(dispatchers, np.ndarray) and dispatchers.dtype.kind in "iu":
            ids = dispatchers.astype(np.int32)
        else:
            ids = np.array([_dispatcher_id(d) for d in dispatchers], dtype=np.int32)
        assert ids.shape == (self.n,)
        if seeds is None:
            seeds = np.arange(self.n, dtype=np.uint64)
        seeds = np.asarray(seeds, dtype=np.uint64)
        assert seeds.shape == (self.n,)
        self.dispatcher_ids = torch.from_numpy(ids).to(self.device)
        self._reset_args = (ids.copy(), seeds.copy(), rng_mode, tapes)
        self.has_external = bool((ids == DISPATCHER_IDS["external"]).any())
        self.seeds = torch.from_numpy(seeds.view(np.int64)).to(self.device)
        mode = RNG_MODES[rng_mode]
        tv = ts = to = None
        if mode == 2:
            assert tapes is not None and len(tapes) == self.n
            offs = np.zeros(self.n + 1, dtype=np.int64)
            for i, (s, v) in enumerate(tapes):
                offs[i + 1] = offs[i] + len(v)
            sites = np.concatenate([np.asarray(s, dtype=np.uint8) for s, _ in tapes]) if offs[-1] else np.zeros(1, np.uint8)
            vals = np.concatenate([np.asarray(v, dtype=np.float64) for _, v in tapes]) if offs[-1] else np.zeros(1)
            tv = torch.from_numpy(vals).to(self.device)
            ts = torch.from_numpy(sites).to(self.device)
            to = torch.from_numpy(offs).to(self.device)
            self._tape = (tv, ts, to)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        self._ck(self.lib.om_reset(self.h, ptr(self.seeds), ptr(self.dispatcher_ids), mode, ptr(tv), ptr(ts), ptr(to),
                                   ptr(self.state), self._stream()))
        return self

    def run(self, until=None, ticks=True, trace_capacity=0, site_ticks=False, order_capacity=0, result=None,
            unit_ticks=False, draw_capacity=0, regrow=False):
        """Advance every instance to simulated time ``until`` (default: the config's sim_time).
        ``draw_capacity`` > 0 (rng_mode="philox" only): export every instance's draw tape (``RunResult.draw_tape``);
        ``draw_tape_slots()`` gives a safe size.
        ``regrow=True`` (whole-episode runs after ``reset()``): instances that stopped because more jams were active at once
        than ``jam_capacity`` holds (OM_IERR_JAM_OVERFLOW) are re-simulated from t = 0 in a temporary handle with a four
        times larger pool -- same config, dispatcher, seed and RNG mode, hence the same trajectory, now complete -- and their
        outputs replace the truncated ones (``result.regrown`` / ``regrown_instances``).  Their state blobs in THIS handle
        stay as they stopped (blob sizes differ between pool sizes), so do not resume such a batch.  Costs one
        synchronising read of the summaries."""
        torch = self.torch
        if getattr(self, "has_external", False):
            raise RuntimeError("BatchedMine.run: the batch holds instances of a host-side dispatcher (device id 255); om_run "
                               "would pause them at their first order and return empty results -- use "
                               "openmines_b200.run_external(dispatcher, ...) or BatchedMineEnv / step() for those")
        until = float(self.sim_time if until is None else until)
        r = result if result is not None else RunResult()
        if r.summary is None:
            r.summary = torch.zeros((self.n, 64), dtype=torch.uint8, device=self.device)
        tick_cap = int(np.ceil(self.sim_time)) if ticks else 0
        ns = self.max_NS
        if ticks and r.ticks is None:
            r.ticks = torch.zeros((self.n, tick_cap, 32), dtype=torch.uint8, device=self.device)
            r.tick_capacity = tick_cap
        if site_ticks and r.site_queue_ticks is None:
            r.site_queue_ticks = torch.zeros((self.n, tick_cap, ns), dtype=torch.int16, device=self.device)
            r.site_tons_ticks = torch.zeros((self.n, tick_cap, ns), dtype=torch.float64, device=self.device)
        if unit_ticks and r.unit_ticks is None:
            r.unit_ticks = torch.zeros((self.n, tick_cap, self.max_U, 16), dtype=torch.uint8, device=self.device)
        if trace_capacity and r.trace is None:
            r.trace = torch.zeros((self.n, trace_capacity, 32), dtype=torch.uint8, device=self.device)
            r.trace_capacity = int(trace_capacity)
        if order_capacity and r.orders is None:
            r.orders = torch.full((self.n, order_capacity), -1, dtype=torch.int16, device=self.device)
            r.order_capacity = int(order_capacity)
        if draw_capacity and r.draw_values is None:
            r.draw_values = torch.zeros((self.n, int(draw_capacity)), dtype=torch.float64, device=self.device)
            r.draw_sites = torch.full((self.n, int(draw_capacity)), DRAW_SKIP, dtype=torch.uint8, device=self.device)
            r.draw_count = torch.zeros(self.n, dtype=torch.int32, device=self.device)
            r.draw_capacity = int(draw_capacity)
        o = OmOutputsStruct()
        ptr = lambda t: t.data_ptr() if t is not None else None
        o.draw_values, o.draw_sites, o.draw_count = ptr(r.draw_values), ptr(r.draw_sites), ptr(r.draw_count)
        o.draw_capacity = r.draw_capacity
        o.summary, o.ticks, o.trace = ptr(r.summary), ptr(r.ticks), ptr(r.trace)
        o.tick_capacity, o.trace_capacity = r.tick_capacity, r.trace_capacity
        o.site_queue_ticks, o.site_tons_ticks = ptr(r.site_queue_ticks), ptr(r.site_tons_ticks)
        o.orders, o.order_capacity = ptr(r.orders), r.order_capacity
        o.unit_ticks = ptr(r.unit_ticks)
        self._ck(self.lib.om_run(self.h, until, ctypes.c_void_p(self.state.data_ptr()), ctypes.byref(o), self._stream()))
        if regrow:
            self._regrow(r, until, site_ticks, unit_ticks)
        return r

    def _regrow(self, r, until, site_ticks, unit_ticks):
        torch = self.torch
        OM_IERR_JAM_OVERFLOW = 3
        todo = np.nonzero(r.summary_np()["err"] == OM_IERR_JAM_OVERFLOW)[0]
        r.regrown, r.regrown_instances = len(todo), [int(i) for i in todo]
        ids, seeds, rng_mode, tapes = self._reset_args
        cap = max(pc.jam_capacity for pc in self.pcs)
        while len(todo):
            if cap >= 4096:
                raise RuntimeError("jam pool overflow with the largest supported pool (4096) in %d instances" % len(todo))
            cap = min(4096, cap * 4)
            raws = [pc.raw for pc in self.pcs]
            sub = BatchedMine(raws if self.multi else raws[0], len(todo), device=self.device.index, jam_capacity=cap,
                              sum_mode=self.pc.sum_mode, rl_surface=False,
                              config_ids=self.config_ids[todo] if self.multi else None)
            sub.reset(ids[todo], seeds=seeds[todo], rng_mode=rng_mode, tapes=[tapes[i] for i in todo] if tapes is not None else None)
            q = sub.run(until=until, ticks=r.ticks is not None, trace_capacity=r.trace_capacity, site_ticks=site_ticks,
                        order_capacity=r.order_capacity, unit_ticks=unit_ticks, draw_capacity=r.draw_capacity)
            idx = torch.from_numpy(todo).to(self.device)
            for f in ("summary", "ticks", "trace", "site_queue_ticks", "site_tons_ticks", "orders", "unit_ticks", "draw_values",
                      "draw_sites", "draw_count"):
                dst, src = getattr(r, f), getattr(q, f)
                if dst is not None and src is not None:
                    dst[idx] = src
            still = q.summary_np()["err"] == OM_IERR_JAM_OVERFLOW
            sub.close()
            todo = todo[still]

    def draw_tape_slots(self, until=None):
        """Upper estimate of the draw-tape slots one instance needs up to ``until``: two per truck per minute (breakdown
        block), up to three per road per minute (jam coin / position / duration) and six per truck cycle."""
        until = float(self.sim_time if until is None else until)
        minutes = int(np.ceil(until)) + 1
        return minutes * (2 * self.max_T + 3 * self.max_R) + 6 * self.max_T * (minutes // 4 + 4)

    def obs_dim(self):
        d = ctypes.c_int32()
        self._ck(self.lib.om_obs_dim(self.h, ctypes.byref(d)))
        return d.value

    def step(self, actions, io, until=None, teacher=-1, auto_reset=False, seed_stride=0, sparse_reward=False):
        """One batched env step through ``om_step``; ``io`` is a dict of preallocated device tensors
        (obs, reward, done, order_kind, truck, sug_action, info; with ``auto_reset`` also final_obs and
        episode_totals int64[N, 3]), ``actions`` an int32 device tensor or None."""
        s = OmStepIOStruct()
        ptr = lambda t: t.data_ptr() if t is not None else None
        s.actions = ptr(actions)
        s.obs, s.reward, s.done = ptr(io["obs"]), ptr(io.get("reward")), ptr(io.get("done"))
        s.order_kind, s.truck = ptr(io.get("order_kind")), ptr(io.get("truck"))
        s.sug_action, s.info = ptr(io.get("sug_action")), ptr(io.get("info"))
        s.teacher = int(teacher)
        s.flags = (OM_STEP_AUTO_RESET if auto_reset else 0) | (OM_STEP_SPARSE_REWARD if sparse_reward else 0)
        s.final_obs, s.episode_totals = ptr(io.get("final_obs")), ptr(io.get("episode_totals"))
        s.seed_stride = int(seed_stride)
        until = float(self.sim_time if until is None else until)
        self._ck(self.lib.om_step(self.h, until, ctypes.c_void_p(self.state.data_ptr()), ctypes.byref(s), self._stream()))

    def stats(self, result):
        """Per-dispatcher partial sums [7, 16] (float64 device tensor, integer-valued: exact and order-independent) for the
        cross-GPU all-reduce; ``openmines_b200.parallel.describe`` decodes them."""
        torch = self.torch
        out = torch.zeros((7, 16), dtype=torch.float64, device=self.device)
        self._ck(self.lib.om_stats_reduce(self.h, ctypes.c_void_p(result.summary.data_ptr()),
                                          ctypes.c_void_p(self.dispatcher_ids.data_ptr()),
                                          ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------------------------------
    def headers(self):
        """Per-instance header records (HDR_DTYPE): clock, counters, error code, pending decision."""
        nb = int(self.layout.hdr_bytes)                    # 128 without the RL surface (the reward fields read as 0)
        raw = np.zeros((self.n, HDR_DTYPE.itemsize), dtype=np.uint8)
        raw[:, :nb] = self.state[:, :nb].contiguous().cpu().numpy()
        return raw.view(HDR_DTYPE).reshape(self.n).copy()

    def state_arrays(self, inst, blob=None):
        """Decode one instance's state blob into named numpy arrays (parity tests, tick recorder, dispatcher views).
        ``blob``: the instance's bytes if the caller already copied the state to the host."""
        if blob is None:
            blob = self.state[inst].cpu().numpy()
        ly, pc = self.layout, self.pc_of(inst)
        T, U, S = pc.T, pc.S + pc.M, pc.S

        raw = blob.tobytes()

        def arr(field, dtype, count):
            off = getattr(ly, field)
            item = np.dtype(dtype).itemsize
            stride = item
            if ly.t_loc and field in ("t_status", "t_cur", "t_tgt"):
                stride = 4                                  # split layouts: {status, cur, tgt, flags} per truck in the hot part
            elif ly.t_aux and field in ("t_unit", "t_next", "t_epoch", "t_svc"):
                stride = 8                                  # split layouts: {unit, next, epoch, svc} per truck in the hot part
            elif ly.t_lastbrk_hot and field == "t_lastbrk":
                # split layouts: the minute of the last breakdown as a uint16 (it is always an integer minute)
                return np.ndarray((count,), dtype=np.uint16, buffer=raw, offset=off, strides=(2,)).astype(np.float64)
            elif field in _ENTRY_FIELDS:
                stride = ly.e_stride
            elif field not in _SOA_FIELDS and field[:2] in ("t_", "u_"):
                stride = ly.t_stride if field[0] == "t" else ly.u_stride
            return np.ndarray((count,), dtype=dtype, buffer=raw, offset=off, strides=(stride,)).copy()
        out = {}
        for f in ("t_time", "t_load", "t_jstart", "t_prejam", "t_lastbrk", "t_repair", "t_total", "t_cycle", "t_dist"):
            out[f] = arr(f, np.float64, T)
        out["t_key"] = arr("t_key", np.uint32, T)
        for f in ("t_svc", "t_epoch", "t_next", "t_unit"):
            out[f] = arr(f, np.uint16, T)
        for f in ("t_status", "t_cur", "t_tgt"):
            out[f] = arr(f, np.uint8, T)
        out["t_phase"] = arr("t_phase", np.uint32, T).astype(np.uint8)
        for f in ("u_last_start", "u_last_done", "u_tons"):
            out[f] = arr(f, np.float64, U)
        out["u_lastbrk"] = arr("u_lastbrk", np.float64, S)   # shovels only (dumpers never break down)
        out["u_svc"] = arr("u_svc", np.uint32, U)
        for f in ("u_user", "u_head", "u_tail", "u_qlen"):
            out[f] = arr(f, np.uint16, U)
        out["u_repair"] = arr("u_repair", np.uint8, S)
        out["r_repair_end"] = arr("r_repair_end", np.float64, pc.R)
        for f in ("s_qtotal", "s_coming"):
            out[f] = arr(f, np.uint16, pc.L + pc.D)
        hb = bytearray(HDR_DTYPE.itemsize)
        hb[:ly.hdr_bytes] = raw[ly.hdr:ly.hdr + ly.hdr_bytes]
        out["hdr"] = np.frombuffer(bytes(hb), dtype=HDR_DTYPE, count=1)[0]
        return out
