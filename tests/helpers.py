"""Shared helpers of the parity tests: golden loading and engine-vs-oracle comparison."""
import os

import numpy as np

from oracle import trace_format as tf

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DISPATCHERS = ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher",
               "SPTFDispatcher", "SQDispatcher"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = z
        self.trace = z["trace"]
        self.reported_tons = float(z["reported_tons"])
        self.dumped_tons = float(z["dumped_tons"])
        self.n_truck_events = int(z["n_truck_events"])
        self.order_count = int(z["order_count"])
        self.error = str(z["error"]) or None
        self.tick = {k[5:]: z[k] for k in z.files if k.startswith("tick_")}
        self.final = {k[6:]: z[k] for k in z.files if k.startswith("final_")}
        self.orders = {k[7:]: z[k] for k in z.files if k.startswith("orders_")}
        self.tape = (z["tape_sites"], z["tape_values"]) if "tape_values" in z.files else None
        self.n_ticks = int(z["n_ticks"])


def trace_diffs(ref, got, rel_time_tol=0.0):
    """Compare two traces (structured arrays).  Kinds/args/trucks exact; times and payloads exact
    unless rel_time_tol > 0.  The truck field of shovel-maintenance records is ignored."""
    diffs = []
    a, b = ref.copy(), got.copy()
    for t in (a, b):
        t["truck"][t["kind"] >= 14] = tf.NO_TRUCK
    if len(a) != len(b):
        diffs.append("trace length %d vs %d" % (len(a), len(b)))
    n = min(len(a), len(b))
    for name in ("kind", "truck", "arg0", "arg1"):
        ne = np.nonzero(a[name][:n] != b[name][:n])[0]
        if len(ne):
            i = int(ne[0])
            diffs.append("trace.%s first differs at %d: ref %r got %r (%d diffs)" % (name, i, a[i], b[i], len(ne)))
    for name in ("time", "a", "b"):
        x, y = a[name][:n], b[name][:n]
        if rel_time_tol > 0:
            bad = np.abs(x - y) > rel_time_tol * np.maximum(np.abs(x), 1e-300)
        else:
            bad = x != y
        ne = np.nonzero(bad)[0]
        if len(ne):
            i = int(ne[0])
            diffs.append("trace.%s first differs at %d: ref %r got %r (%d diffs)" % (name, i, a[i], b[i], len(ne)))
    return diffs


class HostResult:
    """Host copy of a RunResult (one device->host transfer per output instead of one per compared instance); quacks like
    RunResult for ``engine_vs_ref``."""

    def __init__(self, mine, result):
        self._r = result
        for f in ("summary", "ticks", "trace", "site_queue_ticks", "site_tons_ticks", "orders", "unit_ticks",
                  "draw_values", "draw_sites", "draw_count"):
            t = getattr(result, f)
            setattr(self, f, t.cpu() if t is not None else None)
        for f in ("tick_capacity", "trace_capacity", "order_capacity", "draw_capacity"):
            setattr(self, f, getattr(result, f))
        self.state = mine.state.cpu().numpy()
        self._summary = self.summary.numpy().view(type(result).summary_np(result).dtype).reshape(-1)
        self._ticks = result.ticks_np() if result.ticks is not None else None

    def summary_np(self):
        return self._summary

    def ticks_np(self):
        return self._ticks

    def trace_np(self, inst):
        from openmines_b200.trace import TRACE_DTYPE
        return self.trace[inst].numpy().view(TRACE_DTYPE).reshape(-1), self._summary[inst]

    def orders_np(self):
        return self.orders.numpy()

    def draw_tape(self, inst):
        n = int(self.draw_count[inst])
        assert n <= self.draw_capacity, "draw tape overflow: %d > %d" % (n, self.draw_capacity)
        sites, values = self.draw_sites[inst, :n].numpy(), self.draw_values[inst, :n].numpy()
        keep = sites != 255
        return sites[keep].copy(), values[keep].copy()


def _replay_job(job):
    """Worker of ``oracle_replays``: one tape replay through the CPU restatement."""
    from oracle import des_oracle as do
    cfg, dispatcher, tape, kw = job
    return do.run_oracle(cfg, dispatcher, "replay", tape=tape, **kw)


def oracle_replays(jobs, procs=None):
    """Run ``(config, dispatcher, tape, kwargs)`` replays of the CPU oracle on the host cores (spawned workers: the parent
    holds a CUDA context).  Returns the OracleResult objects in job order."""
    import multiprocessing as mp
    procs = procs or max(1, min(len(jobs), os.cpu_count() or 1, 32))
    if procs == 1:
        return [_replay_job(j) for j in jobs]
    with mp.get_context("spawn").Pool(procs) as pool:
        return pool.map(_replay_job, jobs, chunksize=max(1, len(jobs) // (procs * 4)))


def engine_vs_ref(ref, mine, result, inst, sptf=False):
    """All differences between a reference/oracle result object and instance `inst` of an engine run."""
    diffs = []
    tr, s = result.trace_np(inst)
    total = len(ref.trace)
    got = tr[:total] if total <= len(tr) else tr
    diffs += trace_diffs(ref.trace, got)
    if total < len(tr) and (tr[total]["time"] != 0 or tr[total]["kind"] != 0 or tr[total]["truck"] != 0):
        diffs.append("engine trace has extra records beyond %d" % total)
    if int(s["n_truck_events"]) != ref.n_truck_events:
        diffs.append("n_truck_events %d vs %d" % (ref.n_truck_events, s["n_truck_events"]))
    if float(s["reported_tons"]) != ref.reported_tons:
        diffs.append("reported_tons %r vs %r" % (ref.reported_tons, float(s["reported_tons"])))
    if float(s["dumped_tons"]) != ref.dumped_tons:
        diffs.append("dumped_tons %r vs %r" % (ref.dumped_tons, float(s["dumped_tons"])))
    if int(s["n_wrapper_orders"]) != ref.order_count:
        diffs.append("order_count %d vs %d" % (ref.order_count, s["n_wrapper_orders"]))
    ticks = result.ticks_np()[inst]
    nt = ref.n_ticks
    for k_ref, k_eng in (("produced_tons", "produced_tons"), ("service_count", "service_count"),
                         ("working", "working"), ("waiting", "waiting"), ("load_unload", "load_unload"),
                         ("moving", "moving"), ("repairing", "repairing")):
        if not np.array_equal(ref.tick[k_ref], ticks[k_eng][:nt].astype(ref.tick[k_ref].dtype)):
            bad = np.nonzero(ref.tick[k_ref] != ticks[k_eng][:nt])[0]
            diffs.append("tick[%s] differs first at minute %d: %r vs %r" % (
                k_ref, bad[0], ref.tick[k_ref][bad[0]], ticks[k_eng][bad[0]]))
    if result.site_queue_ticks is not None:
        ns = mine.pc_of(inst).L + mine.pc_of(inst).D        # rows are max(L+D) wide in a mixed-config batch
        q = result.site_queue_ticks[inst].cpu().numpy()[:nt, :ns]
        if not np.array_equal(ref.tick["site_queue_total"], q.astype(np.int32)):
            diffs.append("tick[site_queue_total] differs")
        st = result.site_tons_ticks[inst].cpu().numpy()[:nt, :ns]
        if not np.array_equal(ref.tick["site_tons"], st):
            diffs.append("tick[site_tons] differs")
    sa = mine.state_arrays(inst, blob=result.state[inst]) if isinstance(result, HostResult) else mine.state_arrays(inst)
    S = mine.pc_of(inst).S
    checks = [("truck_total_load", sa["t_total"]), ("truck_service_count", sa["t_svc"]),
              ("truck_cycle_time", sa["t_cycle"]), ("truck_status", sa["t_status"]), ("truck_cur", sa["t_cur"]),
              ("truck_tgt", sa["t_tgt"]), ("shovel_tons", sa["u_tons"][:S]), ("shovel_service_count", sa["u_svc"][:S]),
              ("dumper_tons", sa["u_tons"][S:]), ("dumper_service_count", sa["u_svc"][S:])]
    for name, got_arr in checks:
        if not np.array_equal(ref.final[name].astype(np.float64), got_arr.astype(np.float64)):
            diffs.append("final[%s] differs" % name)
    if result.orders is not None:
        row = result.orders_np()[inst]
        row = row[row >= 0]
        kinds, sites = row >> 12, row & 0xFFF
        exp_init = list(ref.orders["init"])
        got_haul = list(sites[kinds == 1])
        got_back = list(sites[kinds == 2])
        if sptf:
            got_init = list(sites[(kinds == 0) | (kinds == 2)])
        else:
            got_init = list(sites[kinds == 0])
        if got_init != exp_init:
            diffs.append("orders[init] differs")
        if got_haul != list(ref.orders["haul"]):
            diffs.append("orders[haul] differs")
        if got_back != list(ref.orders["back"]):
            diffs.append("orders[back] differs")
    return diffs
