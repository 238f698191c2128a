#!/usr/bin/env python
"""Headline benchmark: simulated truck-events/s and episodes/s on north_pit_mine.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference ...                      (CPU arm: the oracle port on the host cores)

A "step" is one pass of the hot path over one batch of synthetic instances: om_reset + om_run of
the whole per-GPU batch for the config's sim_time (240 simulated minutes).  Workload (default):
the per-GPU shard of BASELINE.json configs[2] -- north_pit_mine x 8192 seeds x {SQ, SPTF},
breakdown / repair / road / jam events on (Philox), 16384 instances per GPU, weak scaling.
`--workload c2` runs configs[1] (6 dispatchers x 1024 seeds, stochastic off).

Prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DISPATCHERS = ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher",
               "SPTFDispatcher", "SQDispatcher"]
METRIC = "simulated truck-events/sec (north_pit_mine)"   # c5 runs the synthetic scaled mine instead
UNIT = "truck-events/s"


# CPU port (oracle/des_oracle.py) vs the UNMODIFIED reference on the same workload, same core: seconds per episode of the
# reference divided by the port's, measured in the authoring container by tools/measure_reference_factor.py (BASELINE.md
# section 5).  The reference arm's throughput divided by this factor estimates the unmodified reference on the same cores.
REFERENCE_FACTOR = None      # filled from BASELINE measurements below
try:
    with open(os.path.join(ROOT, "profiles", "reference_factor.json")) as _fh:
        REFERENCE_FACTOR = json.load(_fh)
except Exception:
    REFERENCE_FACTOR = None

_REAL_STDOUT = None

COUNTER_METRICS = ["smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                   "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg"]


def live_counters(device_index, prepare, launch, sync):
    """Hardware counters of the kernels `launch()` enqueues, read in-process through CUPTI (tools/counters/om_counters.cpp:
    range profiler, one user range, user replay -- `prepare()` + `launch()` are repeated once per hardware pass the metric
    set needs, `prepare()` outside the range).  Called AFTER the timed region; returns {metric: value} plus "source", or
    {"error": ...}."""
    import ctypes
    path = os.path.join(ROOT, "tools", "counters", "libom_counters.so")
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        return {"error": "libom_counters.so not loadable: %s" % e}
    lib.omc_last_error.restype = ctypes.c_char_p
    names = (ctypes.c_char_p * len(COUNTER_METRICS))(*[m.encode() for m in COUNTER_METRICS])
    sync()
    if lib.omc_setup(int(device_index), names, len(COUNTER_METRICS)) != 0:
        return {"error": lib.omc_last_error().decode()}
    done, passes = ctypes.c_int(0), 0
    try:
        while not done.value and passes < 32:
            prepare()
            sync()
            if lib.omc_pass_begin() != 0:
                raise RuntimeError(lib.omc_last_error().decode())
            launch()
            sync()
            if lib.omc_pass_end(ctypes.byref(done)) != 0:
                raise RuntimeError(lib.omc_last_error().decode())
            passes += 1
        vals = (ctypes.c_double * len(COUNTER_METRICS))()
        nr = ctypes.c_int(0)
        if lib.omc_finish(vals, ctypes.byref(nr)) != 0:
            raise RuntimeError(lib.omc_last_error().decode())
    except RuntimeError as e:
        lib.omc_abort()
        return {"error": str(e)}
    out = dict(zip(COUNTER_METRICS, [float(v) for v in vals]))
    out["passes"] = passes
    out["source"] = "CUPTI range profiler in this process, after the timed region (%d replay pass(es); tools/counters/om_counters.cpp)" % passes
    return out


def _quiet_stdout():
    """The contract is ONE JSON line on stdout: NCCL / torch may print banners (e.g. "NCCL version ...") on fd 1,
    so fd 1 is pointed at stderr for the whole run and the line is written to the saved descriptor at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def load_cfg(name="north_pit_mine"):
    with open(os.path.join(ROOT, "openmines_b200", "conf", name + ".json"), encoding="utf-8") as fh:
        return json.load(fh)


def workload(name, per_gpu, rank):
    """Returns (dispatcher id array, seed array, rng_mode, description) of this rank's shard."""
    import numpy as np
    if name == "c3":
        n = per_gpu or 16384
        half = n // 2
        gid = np.arange(rank * n, (rank + 1) * n, dtype=np.uint64)     # global instance ids
        disp = np.where(np.arange(n) % 2 == 0, 5, 4).astype(np.int32)  # SQ, SPTF interleaved
        seeds = gid // 2                                               # seed s runs once per dispatcher
        if os.environ.get("OM_BENCH_SORTED"):                          # experiment: same set, grouped by dispatcher
            order = np.argsort(disp, kind="stable")
            disp, seeds = disp[order], seeds[order]
        desc = "BASELINE configs[2] per-GPU shard: north_pit_mine x %d seeds x {SQ,SPTF}, stochastic events on " \
               "(Philox4x32-10), %d instances/GPU, sim_time 240" % (half, n)
        return disp, seeds, "philox", desc
    if name == "c2":
        n_seeds = (per_gpu or 6144) // 6
        n = n_seeds * 6
        disp = np.repeat(np.arange(6, dtype=np.int32), n_seeds)
        seeds = np.tile(np.arange(n_seeds, dtype=np.uint64), 6) + np.uint64(rank * n_seeds)
        desc = "BASELINE configs[1]: north_pit_mine x 6 dispatchers x %d seeds, stochastic off, %d instances/GPU" % (n_seeds, n)
        return disp, seeds, "det", desc
    if name == "c5":
        n = per_gpu or 1024
        disp = np.full(n, 5, dtype=np.int32)
        seeds = np.arange(rank * n, (rank + 1) * n, dtype=np.uint64)
        desc = "BASELINE configs[4]: synthetic scaled mine (500 trucks, 20 load sites, 10 dump sites), SQ, " \
               "stochastic events on, %d instances/GPU, sim_time 240" % n
        return disp, seeds, "philox", desc
    raise SystemExit("unknown workload %r" % name)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_episode(args):
    disp, seed, mode = args
    from oracle import des_oracle as do
    t0 = time.perf_counter()
    r = do.DesOracle(load_cfg(), disp, mode, seed, record_trace=False).run()
    return r.n_truck_events, time.perf_counter() - t0


def cpu_sample(workload_name, episodes_per_core, cores=None):
    """Times the CPU oracle port (oracle/des_oracle.py) on `cores` host processes."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    n = max(1, episodes_per_core) * cores
    if workload_name == "c3":
        jobs = [(("SQDispatcher", "SPTFDispatcher")[i % 2], i // 2, "live") for i in range(n)]
    else:
        jobs = [(DISPATCHERS[i % 6], i // 6, "det") for i in range(n)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_episode, jobs[:cores])           # warm the workers (imports, page-in)
        t0 = time.perf_counter()
        res = pool.map(_cpu_episode, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    ev = sum(r[0] for r in res)
    return {"events": ev, "episodes": n, "wall_s": wall, "cores": cores,
            "events_per_s": ev / wall, "episodes_per_s": n / wall}


def _factor_note():
    if not REFERENCE_FACTOR:
        return ""
    return ("; the port runs %.2fx faster per core than the unmodified reference on this workload (measured in the "
            "authoring container, %s), so the unmodified reference would score about value / %.2f" % (
                REFERENCE_FACTOR["port_speedup"], REFERENCE_FACTOR["how"], REFERENCE_FACTOR["port_speedup"]))


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 8
    for _ in range(a.warmup and 1):
        cpu_sample(a.workload, 1, cores)
    vals, eps, walls = [], [], []
    for _ in range(a.steps):
        s = cpu_sample(a.workload, per_core, cores)
        vals.append(s["events_per_s"]); eps.append(s["episodes_per_s"]); walls.append(s["wall_s"])
    v = sum(vals) / len(vals)
    _, _, _, desc = workload(a.workload, a.instances, 0)
    sample = "%d episodes per step (%d per core x %d processes) of the same workload, oracle/des_oracle.py " \
             "(pure-Python port of the reference's SimPy path; the reference itself is Python and needs SimPy, " \
             "which is not installed)%s" % (per_core * cores, per_core, cores, _factor_note())
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000 * sum(walls) / len(walls), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "sim_time": 240},
            "episodes_per_s": sum(eps) / len(eps),
            "cpu_baseline": dict({"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                                 **({"reference_factor": REFERENCE_FACTOR["port_speedup"]} if REFERENCE_FACTOR else {})),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


# ------------------------------------------------------------------------------------------------
# BASELINE configs[3]: batched RL env step, random-init policy feeding the orders
# ------------------------------------------------------------------------------------------------
def run_c4(a, emit=True):
    """Steady-state batched RL rollout: policy forward + om_step per env step, every env restarting by itself when its
    episode ends (per-env auto reset).  One bench step = 500 env steps of all envs (about one episode's decisions)."""
    import numpy as np
    import torch
    import openmines_b200 as om
    from openmines_b200.engine import HDR_DTYPE
    from openmines_b200.rl_env import BatchedMineEnv
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.instances or 16384
    steps_per = 500
    cfg = load_cfg()
    env = BatchedMineEnv(cfg, n, device=local, rng_mode="philox", sug_dispatcher="ShortestTripDispatcher", auto_reset=True)
    dev = env.mine.device
    torch.manual_seed(1)
    torch.set_float32_matmul_precision("high")      # the policy's GEMMs on the tensor cores (TF32); the env itself is fp64
    # policy of ppo_single_net.Agent shape (openmines/test/cleanrl/ppo_single_net.py:124-185): 194-256-256-256-5, ReLU, orthogonal init
    layers, d = [], env.obs_dim
    for width in (256, 256, 256):
        lin = torch.nn.Linear(d, width); torch.nn.init.orthogonal_(lin.weight, 2 ** 0.5); torch.nn.init.zeros_(lin.bias)
        layers += [lin, torch.nn.ReLU()]; d = width
    head = torch.nn.Linear(d, env.n_actions); torch.nn.init.orthogonal_(head.weight, 0.01); torch.nn.init.zeros_(head.bias)
    policy = torch.nn.Sequential(*layers, head).to(dev)
    # one rollout iteration (policy forward + Categorical sample + om_step) captured in a CUDA graph
    replay = env.capture_rollout_step(policy) if not a.no_graph else None
    obs_box = [None]

    def iterate(k):
        with torch.no_grad():
            for _ in range(k):
                if replay is not None:
                    replay()
                else:
                    act = torch.distributions.Categorical(logits=policy(obs_box[0])).sample()
                    obs_box[0] = env.step(act)[0]

    def totals():
        """finished episodes, truck events, decisions so far (device-accumulated totals + the running episodes' counters)"""
        acc = env.io["episode_totals"].sum(dim=0).cpu().numpy().astype(np.float64)
        hdr = env.mine.headers()
        return np.array([acc[0], acc[1] + hdr["n_events"].astype(np.float64).sum(),
                         acc[2] + hdr["n_decisions"].astype(np.float64).sum()])

    obs_box[0] = env.reset(seed=rank * n)[0]
    iterate(steps_per + 100 * max(0, a.warmup - 1))       # warm-up: past the first episode ends, so the envs are out of step
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_before = totals()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    iterate(a.steps * steps_per)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    delta = totals() - t_before
    tot = torch.tensor([ms, delta[0], delta[1], delta[2]], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = tot[:1].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot[1:].clone(); dist.all_reduce(sm)
        ms, delta = float(mx[0]), sm.cpu().numpy()
    n_episodes, n_events, n_dec = float(delta[0]), float(delta[1]), float(delta[2])
    herr = env.mine.headers()["err"]
    failed_t = torch.tensor([float(((herr != 0) & (herr != 5)).sum())], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(failed_t)
    n_failed = float(failed_t.item())
    ms_per_step = ms / a.steps
    sec = ms / 1000.0
    # ---- end to end with the policy on the HOST (the reference's trainers: actions arrive from host memory, observations,
    # rewards and flags go back to it every env step through the public env.step())
    e2e_steps = 200
    h_act = torch.zeros(n, dtype=torch.int32).pin_memory()
    h_obs = torch.empty((n, env.obs_dim), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float64).pin_memory()
    h_done = torch.empty(n, dtype=torch.bool).pin_memory()
    table = torch.randint(0, env.n_actions, (16, n), dtype=torch.int32)      # the host "policy": pre-drawn random actions

    def host_loop(k):
        sink = 0.0
        for q in range(k):
            h_act.copy_(table[q & 15])
            o, r, term, _, _ = env.step(h_act.to(dev, non_blocking=True))
            h_obs.copy_(o, non_blocking=True); h_rew.copy_(r, non_blocking=True); h_done.copy_(term, non_blocking=True)
            torch.cuda.synchronize()
            sink += float(h_rew[0]) + float(h_obs[0, 3]) + float(h_done[0])       # the host reads the step's results
        return sink

    host_loop(20)
    if dist is not None:
        dist.barrier()
    t_before = totals()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    host_loop(e2e_steps)
    h1.record()
    torch.cuda.synchronize()
    e2e_ms = h0.elapsed_time(h1)
    e2e_delta = totals() - t_before
    e2e_tot = torch.tensor([e2e_ms, e2e_delta[1], e2e_delta[2]], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = e2e_tot[:1].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = e2e_tot[1:].clone(); dist.all_reduce(sm)
        e2e_tot = torch.cat([mx, sm])
    e2e_ms, e2e_events, e2e_dec = [float(x) for x in e2e_tot.cpu()]
    # kernel-only timing of om_step for the roofline: the same rollout (the policy keeps choosing the actions, so the envs stay
    # in the regime `value` was measured in), launched eagerly with one event pair around every om_step
    act = torch.zeros(n, dtype=torch.int32, device=dev)

    def policy_action():
        with torch.no_grad():
            act.copy_(torch.distributions.Categorical(logits=policy(env.io["obs"]), validate_args=False).sample())

    for _ in range(20):
        policy_action()
        env._step(act)
    pairs = []
    n_timed = 600            # a whole episode's worth of decisions (~560): the step cost varies with the episode phase
    for _ in range(n_timed):
        policy_action()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        env._step(act)
        k1.record()
        pairs.append((k0, k1))
    torch.cuda.synchronize()
    step_ms = sum(x.elapsed_time(y) for x, y in pairs) / len(pairs)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = n * (2 * env.mine.state_bytes + env.obs_dim * 4 + 4 + 8 + 1 + 3 * 4 + 32)
    achieved = alg / (step_ms / 1000.0) / 1e9
    counters = {"error": "skipped"}
    if rank == 0 and not a.no_counters:
        counters = live_counters(local, policy_action, lambda: env._step(act), torch.cuda.synchronize)
    issue = None
    if "error" not in counters:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        inst = counters["smsp__inst_executed.sum"]
        issue = {"warp_inst_per_launch": inst, "warp_inst_per_env_step": inst / n,
                 "achieved_inst_per_s": inst / (step_ms / 1000.0), "peak_inst_per_s": sms * 4 * sm_mhz * 1e6,
                 "frac": inst / (step_ms / 1000.0) / (sms * 4 * sm_mhz * 1e6),
                 "issue_active_pct": counters.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "dram_bytes": counters["dram__bytes_read.sum"] + counters["dram__bytes_write.sum"],
                 "source": counters["source"]}
    line = {"metric": METRIC, "value": n_events / sec,
            "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: batched RL env step, %d parallel north_pit_mine envs/GPU, random-init "
                                   "policy (194-256-256-256-5 MLP with TF32 GEMMs, Categorical sample) feeding every order; steady state with "
                                   "per-env auto reset; one bench step = %d env steps of every env; rollout iteration %s" %
                                   (n, steps_per, "launched eagerly" if a.no_graph else "replayed from a CUDA graph"),
                       "sim_time": 240, "instances_per_gpu": n, "state_bytes_per_instance": env.mine.state_bytes,
                       "l2_policy": "inputs larger than L2 (state %.0f MB/GPU round-trips HBM every env step)" % (n * env.mine.state_bytes / 1e6),
                       "launch": env.mine.launch_info()},
            "episodes_per_s": n_episodes / sec,
            "instance_errors": n_failed,          # envs stopped by an engine error (e.g. a jam pool that overflowed): they never restart
            "decisions_per_s": n_dec / sec,
            "env_steps_per_s": n * world * a.steps * steps_per / sec,
            "clocks": clocks,
            "e2e": {"value": e2e_events / (e2e_ms / 1000.0), "unit": UNIT,
                    "decisions_per_s": e2e_dec / (e2e_ms / 1000.0),
                    "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": n * (env.obs_dim * 4 + 8 + 1),
                    "note": "policy on the HOST: per env step the actions come from pinned host memory and observations, rewards "
                            "and done flags go back to it (env.step of the public BatchedMineEnv, %d steps, one synchronisation "
                            "per step); `value` is the device-resident rollout (policy and env share the GPU, nothing crosses "
                            "PCIe)" % e2e_steps},
            "gpu_launches": a.steps * steps_per,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": (issue or {}).get("dram_bytes"), "issue": issue if issue else counters,
                         "kernel": "om_step_kernel<philox>", "kernel_ms": step_ms,
                         "algorithmic_bytes_per_launch": alg,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                         "note": "kernel_ms: mean over 600 consecutive om_step launches (a whole episode: the cost of a step varies with the episode phase, 0.29-0.42 ms) of the same policy-driven rollout, one CUDA-event pair around each"}}
    if not emit:
        env.close()
        return line
    if rank == 0:
        emit_line(line)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# extra: the reference's fleet-size sweep (cli/run.py:283-362) as ONE mixed-config launch vs one launch per fleet size
# ------------------------------------------------------------------------------------------------
def run_ablation(a):
    import numpy as np
    import torch
    from openmines_b200 import experiments as ex
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cfg = load_cfg()
    names = ["NaiveDispatcher", "RandomDispatcher", "NearestDispatcher", "FixedGroupDispatcher", "SPTFDispatcher", "SQDispatcher"]
    n_seeds = a.instances or 32
    lo, hi = 10, 160

    def timed(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / a.steps, out

    t_one, r_one = timed(lambda: ex.scene_based_fleet_ablation(cfg, lo, hi, dispatchers=names, n_seeds=n_seeds, device=local))
    t_ser, r_ser = timed(lambda: ex.scene_based_fleet_ablation_serial(cfg, lo, hi, dispatchers=names, n_seeds=n_seeds, device=local))
    same = all(r_one[k]["productions"] == r_ser[k]["productions"] for k in names)
    episodes = 10 * len(names) * n_seeds
    line = {"metric": "fleet-ablation sweep wall time (north_pit_mine, 10 fleet sizes x 6 dispatchers x %d seeds)" % n_seeds,
            "value": t_one, "unit": "s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * t_one,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "scene_based_fleet_ablation: fleet sizes %s, %d instances, stochastic events on; wall time "
                                   "includes config packing, om_create_multi, reset, run and the summary copy" % (
                                       ex.fleet_sizes_for(lo, hi), episodes)},
            "one_launch_s": t_one, "serial_launches_s": t_ser, "speedup": t_ser / t_one, "identical_results": bool(same),
            "episodes_per_s": episodes / t_one}
    emit_line(line)


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs, measured in the same run (short): the driver only runs the default line, so the numbers of
# configs[1] / configs[4] ride along under "extras" instead of existing only in builder-run files
# ------------------------------------------------------------------------------------------------
def measure_extra(name, rank, local, dist, steps=3, warmup=3):
    import numpy as np
    import torch
    import openmines_b200 as om
    if name == "c5":
        from openmines_b200.synthetic import scaled_mine
        cfg = scaled_mine()
    else:
        cfg = load_cfg()
    disp, seeds, mode, desc = workload(name, 0, rank)
    mine = om.BatchedMine(cfg, len(disp), device=local, jam_capacity=512 if name == "c5" else 64)
    mine.reset(disp, seeds=seeds, rng_mode=mode)
    res = mine.run()

    def step():
        mine._ck(mine.lib.om_reset(mine.h, mine.seeds.data_ptr(), mine.dispatcher_ids.data_ptr(),
                                   om.engine.RNG_MODES[mode], None, None, None, mine.state.data_ptr(), mine._stream()))
        mine.run(result=res)
        st = mine.stats(res)
        if dist is not None:
            dist.all_reduce(st)
        return st
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=mine.device)
    evs = []
    for _ in range(steps):
        flush.fill_(1)                                   # L2 flushed before every timed step (outside the events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(x.elapsed_time(y) for x, y in evs)
    t = torch.tensor([ms], dtype=torch.float64, device=mine.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    stats = st.cpu().numpy()
    out = {"workload": desc, "value": float(stats[:, 3].sum()) / (ms_per_step / 1000.0), "unit": UNIT, "ms_per_step": ms_per_step,
           "episodes_per_s": float(stats[:, 0].sum()) / (ms_per_step / 1000.0), "instance_errors": float(stats[:, 9].sum()),
           "steps": steps, "warmup": warmup, "launch": mine.launch_info(), "l2_policy": "L2 flushed (512 MB write) before every timed step"}
    mine.close()
    del flush
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5", "ablation"])
    ap.add_argument("--instances", type=int, default=0, help="instances per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-counters", action="store_true", help="skip the CUPTI counter pass after the timed region")
    ap.add_argument("--no-extras", action="store_true", help="default workload only: skip the short configs[1] / configs[4] measurements")
    ap.add_argument("--no-graph", action="store_true", help="c4: launch the rollout iteration eagerly instead of replaying a CUDA graph")
    a = ap.parse_args()
    _quiet_stdout()
    if a.impl == "reference":
        return run_reference_arm(a)
    if a.workload == "c4":
        return run_c4(a)
    if a.workload == "ablation":
        return run_ablation(a)

    import numpy as np
    import torch
    import openmines_b200 as om

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == a.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % a.gpus

    if a.workload == "c5":
        from openmines_b200.synthetic import scaled_mine
        cfg = scaled_mine()
    else:
        cfg = load_cfg()
    disp, seeds, mode, desc = workload(a.workload, a.instances, rank)
    n = len(disp)
    mine = om.BatchedMine(cfg, n, device=local, jam_capacity=512 if a.workload == "c5" else 64)
    info = mine.launch_info()
    dev = mine.device
    T = int(cfg["sim_time"])

    # ---- device-resident arm: inputs already in HBM ------------------------------------------
    mine.reset(disp, seeds=seeds, rng_mode=mode)          # uploads seeds / dispatcher ids once
    res = mine.run()
    torch.cuda.synchronize()

    def step_device():
        mine._ck(mine.lib.om_reset(mine.h, mine.seeds.data_ptr(), mine.dispatcher_ids.data_ptr(),
                                   om.engine.RNG_MODES[mode], None, None, None, mine.state.data_ptr(), mine._stream()))
        mine.run(result=res)
        st = mine.stats(res)
        if dist is not None:
            dist.all_reduce(st)                           # the only collective: production statistics
        return st

    for _ in range(a.warmup):
        step_device()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # L2 policy: the C3 shard's state (199 MB) exceeds the 126 MB L2, so consecutive steps cannot reuse it;
    # smaller workloads flush L2 (write a 512 MB buffer) before every timed step, outside the timed events.
    need_flush = n * mine.state_bytes < 2 * 126 * 1024 * 1024
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev) if need_flush else None
    run_evs, step_evs = [], []
    torch.cuda.synchronize()
    for _ in range(a.steps):
        if need_flush:
            flush_buf.fill_(1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        mine._ck(mine.lib.om_reset(mine.h, mine.seeds.data_ptr(), mine.dispatcher_ids.data_ptr(),
                                   om.engine.RNG_MODES[mode], None, None, None, mine.state.data_ptr(), mine._stream()))
        # bracket the episode kernel alone for the roofline figure
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        mine.run(result=res)
        k1.record()
        run_evs.append((k0, k1))
        st = mine.stats(res)
        if dist is not None:
            dist.all_reduce(st)
        s1.record()
        step_evs.append((s0, s1))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    elapsed_ms = sum(x.elapsed_time(y) for x, y in step_evs)
    kern_ms = sum(x.elapsed_time(y) for x, y in run_evs) / len(run_evs)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    stats = st.cpu().numpy()                              # all-reduced totals over all ranks
    total_events = float(stats[:, 3].sum())
    total_instances = float(stats[:, 0].sum())
    n_err = float(stats[:, 9].sum())
    ms_per_step = elapsed_ms / a.steps
    value = total_events / (ms_per_step / 1000.0)
    episodes_per_s = total_instances / (ms_per_step / 1000.0)

    # ---- end-to-end arm: host buffers in, host results out, every step ------------------------
    # The public streaming API (openmines_b200.pipeline.EpisodePipeline): every step uploads its (seed, dispatcher)
    # arrays from pinned host memory, resets, simulates, and brings summaries + per-minute KPI records back to pinned
    # host memory, where they are read.  Outputs are double-buffered, so step k's device->host copy overlaps step k+1's
    # kernel; all copies of all K steps are inside the timed region.
    from openmines_b200.pipeline import EpisodePipeline
    pipe = EpisodePipeline(mine, rng_mode=mode)

    def e2e_steps(k):
        tons = 0.0
        for _ in range(k):
            pipe.submit(disp, seeds)
            if pipe.pending == pipe.depth:
                summ, ticks = pipe.collect()
                tons += float(summ["reported_tons"].sum()) + float(ticks["produced_tons"][:, -1].sum())   # the caller reads the curves
        while pipe.pending:
            summ, ticks = pipe.collect()
            tons += float(summ["reported_tons"].sum()) + float(ticks["produced_tons"][:, -1].sum())
        return tons

    e2e_steps(2)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps(a.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = total_events / (e2e_s / a.steps)
    h2d = pipe.h2d_bytes
    d2h = pipe.d2h_bytes

    # ---- roofline of the dominant kernel (om_run_kernel) --------------------------------------
    # The episode kernel keeps each instance's state in shared memory for the whole episode, so its HBM traffic is ~0 by
    # construction; what bounds it is warp-instruction issue.  Primary figure: executed warp
    # instructions per launch (read live from the hardware counters after the timed region) / the kernel's event-timed
    # duration / (148 SMs x 4 schedulers x SM clock).  The HBM figure is reported beside it.
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = n * (2 * mine.state_bytes + T * 32 + 64)  # state in + out, tick KPI records, summary
    achieved_hbm = alg_bytes / (kern_ms / 1000.0) / 1e9
    counters = {"error": "skipped"}
    if rank == 0 and not a.no_counters:
        def _prepare():
            mine._ck(mine.lib.om_reset(mine.h, mine.seeds.data_ptr(), mine.dispatcher_ids.data_ptr(),
                                       om.engine.RNG_MODES[mode], None, None, None, mine.state.data_ptr(), mine._stream()))

        def _launch():
            mine.run(result=res)
        counters = live_counters(local, _prepare, _launch, torch.cuda.synchronize)
    kc = None
    if "error" in counters:
        # fallback: the committed ncu capture of the same kernel / workload, scaled by the instance count
        try:
            with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as fh:
                kc = json.load(fh).get(a.workload)
        except Exception:
            kc = None
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_issue = sms * 4 * sm_mhz * 1e6
    inst = traffic = None
    if "error" not in counters:
        inst = counters["smsp__inst_executed.sum"]
        traffic = counters["dram__bytes_read.sum"] + counters["dram__bytes_write.sum"]
        src = counters["source"]
    elif kc:
        scale = n / float(kc["instances"])
        inst = kc["smsp__inst_executed.sum"] * scale
        traffic = (kc["dram__bytes_read.sum"] + kc["dram__bytes_write.sum"]) * scale
        src = kc["source"] + " (committed ncu capture scaled by the instance count; live counters unavailable: %s)" % counters["error"]
    hbm_part = {"achieved": achieved_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_hbm / hbm_peak,
                "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"}
    if inst is not None:
        ach = inst / (kern_ms / 1000.0)
        roofline = {"bound": "issue", "achieved": ach, "peak": peak_issue, "unit": "warp-inst/s", "frac": ach / peak_issue,
                    "traffic": traffic, "kernel": "om_run_kernel<%s>" % mode, "kernel_ms": kern_ms,
                    "warp_inst_per_launch": inst, "counter_source": src,
                    "peak_source": "%d SMs x 4 schedulers x %.0f MHz (median SM clock sampled during the timed region)" % (sms, sm_mhz),
                    "issue_active_pct": counters.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "hbm": hbm_part,
                    "note": "episode-resident DES: state lives in shared memory for the whole episode, HBM traffic ~0 by "
                            "construction (hbm.frac); the limiter is warp-instruction issue / dependent-chain latency"}
    else:
        roofline = dict(hbm_part, bound="hbm", traffic=None, kernel="om_run_kernel<%s>" % mode, kernel_ms=kern_ms,
                        note="instruction counters unavailable (%s): only the HBM figure, which is ~0 by construction" % counters["error"])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "sim_time": T, "instances_per_gpu": n, "state_bytes_per_instance": mine.state_bytes,
                       "l2_policy": ("L2 flushed (512 MB write) before every timed step; state %.0f MB/GPU" if need_flush else
                                     "inputs larger than L2 (state %.0f MB/GPU per step)") % (n * mine.state_bytes / 1e6),
                       "launch": info, "parallelism": "instances sharded over %d GPU(s); one NCCL all-reduce of "
                                                      "production statistics per step" % world},
            "episodes_per_s": episodes_per_s, "events_per_step": total_events, "instance_errors": n_err,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000 * e2e_s / a.steps},
            "gpu_launches": a.steps * 4,      # om_reset_kernel + om_order_kernel + om_run_kernel + om_stats_kernel per step
            "roofline": roofline}
    if a.workload == "c3" and not a.no_extras:
        # BASELINE configs[1] (det, 6 dispatchers x 1024 seeds) and configs[4] (500-truck mine, 1024 instances per GPU), short
        line["extras"] = {"c2": measure_extra("c2", rank, local, dist), "c5": measure_extra("c5", rank, local, dist)}
        if world == 1:
            # BASELINE configs[3]: the batched RL env step with the policy in the loop (one bench step = 500 env steps)
            import copy
            a4 = copy.copy(a)
            a4.steps, a4.warmup, a4.instances, a4.no_counters = 1, 2, 0, True
            c4 = run_c4(a4, emit=False)
            line["extras"]["c4"] = {"workload": c4["config"]["workload"], "value": c4["value"], "unit": UNIT,
                                    "decisions_per_s": c4["decisions_per_s"], "env_steps_per_s": c4["env_steps_per_s"],
                                    "ms_per_step": c4["ms_per_step"], "om_step_kernel_ms": c4["roofline"]["kernel_ms"],
                                    "hbm_frac": c4["roofline"]["frac"], "instance_errors": c4.get("instance_errors"), "e2e_host_policy": c4["e2e"]}
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            s = cpu_sample(a.workload, 32)
            line["cpu_baseline"] = {"value": s["events_per_s"], "unit": UNIT, "cores": s["cores"], "kind": "port",
                                    "episodes_per_s": s["episodes_per_s"],
                                    "sample": "%d episodes of the same workload (%.1f s wall) on %d host processes, "
                                              "oracle/des_oracle.py%s" % (s["episodes"], s["wall_s"], s["cores"], _factor_note())}
            if REFERENCE_FACTOR:
                line["cpu_baseline"]["reference_factor"] = REFERENCE_FACTOR["port_speedup"]
                line["cpu_baseline"]["unmodified_reference_estimate"] = s["events_per_s"] / REFERENCE_FACTOR["port_speedup"]
        emit_line(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
