#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused reach-env step kernel (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE launch of the fused kernel = one Env.step() for a whole batch of N_envs = 4096 arms (BASELINE
configs[1]: rl_reach_env N_envs=4096, 1xB200; with N ranks every rank owns its own 4096 envs -> weak scaling, and
N = 8 is configs[4], 32768 envs sharded 8 x 4096; no data-path collective).

Timing hygiene: the 4096-env working set (0.5 MB) would sit in L2, so the bench keeps a POOL of independent
4096-env batches whose touched state exceeds 2x the 126 MB L2 and steps them round-robin: every launch reads and
writes HBM-cold state.  Actions: a ring of 61 independent U(-0.7,0.7) sets (coprime with the pool size), so every env
sees a different action on each of its steps and random-walks through the workspace like under an untrained policy
(a constant action per env would pin the arms in workspace corners).  The pool is walked INDEPENDENTLY of K: the timed
unit -- exactly K launches between two event-record nodes -- is laid out INSIDE CUDA graphs that together cover the whole
pool, so consecutive replays never touch the same batches whatever K the caller picks, and the boundary events sit on a
side branch of the graph so that the launches keep their kernel-to-kernel chain across unit boundaries (K = 20 and
K = 2000 measure the same thing to ~1 %; capture_timed_units).  `value` is device-timed: after ~0.3 s of untimed replays
(clocks up, envs in mid-episode states) >= 60 unit intervals are collected, the whole set bracketed by a barrier +
synchronize; per rank the MEDIAN interval counts (p10 / p90 are printed too), max over ranks.  `e2e` is the same
metric through the host-buffer C-ABI call (armsim_step_host on the handle's pinned block: the kernel reads the actions
from / writes the results to host memory over PCIe every step, the host polls per-block doorbells), timed on the host
clock as R repeats of a K-step loop, median -- with the handle's resident step server on (armsim_host_server: one kernel
stays on the GPU and serves a step per command word) and, beside it as e2e.launch_per_step, with one launch per step.
Secondary numbers on the same JSON line (single GPU only): other_configs (push / pick / large-N reach, measured the
same way), rollout_with_td3_actor, e2e.pipelined_depth2 (step_async / step_wait over two env groups).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

N_ENVS = 4096
ALGO_BYTES = {"reach": 118, "push": 242, "pick": 250, "kuka_reach": 106,     # SURVEY 8(d), per env-step
              # torque mode (north star's articulated-body step): reads tau 28 + q 28 + qd 28 + goal 12 + step 4,
              # writes q 28 + qd 28 + step 4 + obs [ee, goal, q, qd] 80 + reward 4 + done 1 + success 1
              "reach_torque": 246}
L2_BYTES = 126 * 1024 * 1024
N_ACT = 61                                                                   # action sets in rotation (prime)
METRIC = "env-steps/sec (reach, N_envs=4096)"
UNIT = "env-steps/s"


def load_traffic(task, n):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None"""
    try:
        for name, key in (("r02_ncu_summary.json", "r02_%s_n%d" % (task, n)), ("r01_ncu_summary.json", "v5_%s_n%d" % (task, n))):
            p = os.path.join(ROOT, "profiles", name)
            if os.path.exists(p) and key in json.load(open(p)):
                d = json.load(open(p))[key]
                break
        return d["dram__bytes_read.sum"]["value"] * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0}[d["dram__bytes_read.sum"]["unit"]] + \
            d["dram__bytes_write.sum"]["value"] * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0}[d["dram__bytes_write.sum"]["unit"]]
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """samples SM clock / throttle reasons of one GPU while the timed regions run (pynvml, 10 ms period)"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v and k not in ("nvmlClocksThrottleReasonAll", "nvmlClocksEventReasonAll"):
                    names[v] = k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit and name not in ("None", "GpuIdle"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- CPU baseline
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _oracle_rate(O, tid, n_envs, acts, nthreads, seconds):
    """env-steps/s of the C fp64 restatement with `nthreads` persistent C worker threads (oracle orc_run_mt: the whole
    K-step loop runs inside ONE C call, so no Python / ctypes dispatch sits between steps)"""
    sim = O.OracleSim(O.default_config(tid, n_envs=n_envs, seed=0, auto_reset=1))
    sim.reset()
    sim.run_mt(acts, 4, nthreads)                                 # threads up, caches warm
    t0 = time.perf_counter()
    sim.run_mt(acts, 8, nthreads)
    per_step = (time.perf_counter() - t0) / 8
    k = max(8, int(seconds / max(per_step, 1e-6)))
    t0 = time.perf_counter()
    sim.run_mt(acts, k, nthreads)
    dt = time.perf_counter() - t0
    sim.close()
    return n_envs * k / dt, k, dt


def cpu_baseline(task="reach", n_envs=N_ENVS, target_seconds=12.0, threads=None):
    """The CPU restatement of the reference step (oracle/, C, fp64) on the host cores: the reference's own step is
    Python on pybullet==3.0.6, which cannot be installed here -> kind = "port"."""
    from oracle import oracle as O
    O.build()
    tid = {"reach": O.TASK_REACH, "push": O.TASK_PUSH, "pick": O.TASK_PICK}[task]
    cores = threads or host_threads()
    rng = np.random.default_rng(1)
    acts = rng.uniform(-0.7, 0.7, (8, n_envs, 3)).astype(np.float32)
    one_core, k1, _ = _oracle_rate(O, tid, n_envs, acts, 1, target_seconds / 3)
    all_core, kc, dt = _oracle_rate(O, tid, n_envs, acts, cores, target_seconds * 2 / 3) if cores > 1 else (one_core, k1, 0.0)
    return {"value": all_core, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d envs x %d steps (%.1f s) of the C fp64 restatement of RLReachEnv.step (oracle/) on %d persistent C "
                      "worker threads, the step loop inside one C call; PyBullet itself is not installable offline" % (n_envs, kc, dt, cores),
            "single_core_value": one_core, "parallel_efficiency": all_core / (cores * one_core)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, max(args.warmup, 1)
    from oracle import oracle as O
    O.build()
    cores = host_threads()
    n = N_ENVS
    # each "step" = one Env.step of all 4096 envs on all host threads; the whole run is bounded to ~60 s whatever K is
    sim = O.OracleSim(O.default_config(O.TASK_REACH, n_envs=n, seed=0, auto_reset=1))
    sim.reset()
    rng = np.random.default_rng(1)
    acts = rng.uniform(-0.7, 0.7, (8, n, 3)).astype(np.float32)
    sim.run_mt(acts, 2, cores)
    t0 = time.perf_counter()
    sim.run_mt(acts, 4, cores)
    per_step = (time.perf_counter() - t0) / 4
    budget = 60.0
    k_eff = max(1, min(steps, int(budget / max(per_step, 1e-6))))
    w_eff = max(1, min(warm, max(1, int(5.0 / max(per_step, 1e-6)))))
    sim.run_mt(acts, w_eff, cores)
    # R repeats of the K-step loop (like the GPU arm: a 20-step loop is ~20 ms, one sample of it is noise), median
    reps = max(1, min(50, int(budget / max(per_step * k_eff, 1e-6))))
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        sim.run_mt(acts, k_eff, cores)
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    val = n * k_eff / dt
    sim.close()
    sample = ("median of %d repeats of %d Env.step batches of %d envs (of the %d requested; bounded to ~%.0f s), C fp64 "
              "restatement of RLReachEnv.step on %d persistent C worker threads" % (reps, k_eff, n, steps, budget, cores))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k_eff,
            "warmup": w_eff, "ms_per_step": 1e3 * dt / k_eff, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "rl_reach_env N_envs=4096, Env.step on the CPU (reference path restated in C; "
                                   "pybullet==3.0.6 unavailable offline)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "timing": {"repeats": reps, "p10_ms_per_step": 1e3 * float(np.percentile(times, 10)) / k_eff,
                       "p90_ms_per_step": 1e3 * float(np.percentile(times, 90)) / k_eff},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
REPEATS = 60          # timed replays per measurement (each with its own CUDA event pair); the median counts


def capture_pool_graphs(torch, envs, actions, steps, stream, first=0):
    """The K-launch unit as CUDA graphs that together walk the WHOLE pool, whatever K is: graph g holds launches
    [g*K, (g+1)*K) of the endless sequence `launch j -> batch j % pool, action set j % len(actions)`; there are
    ceil(pool / K) graphs (at least 1), so one turn through the graphs touches every batch of the pool and consecutive
    replays never hit state that is still in L2."""
    pool = len(envs)
    n_graphs = max(1, -(-pool // steps))
    graphs = []
    for g in range(n_graphs):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            for j in range(first + g * steps, first + (g + 1) * steps):
                envs[j % pool].step(actions[j % len(actions)])
        graphs.append(gr)
    return graphs


def capture_timed_units(torch, envs, actions, steps, stream, first=0, side_events=True):
    """The headline measurement: CUDA graphs made of timed UNITS laid back to back, one unit = exactly K fused-step
    launches between two event-record nodes (torch.cuda.Event(external=True) records become graph nodes; consecutive
    units share their boundary event).  Inside a graph the units follow each other kernel-to-kernel, so a unit's
    interval contains K launches and nothing else -- no graph-launch front-end latency, which at K = 20 is ~7 % of a
    replay timed from outside.  side_events=True: the boundary events are recorded on a SIDE branch of the graph (a
    second captured stream that waits for the unit's last launch), so the record node is a leaf and the launches keep
    their kernel-to-kernel programmatic-dependent-launch chain across unit boundaries -- each interval runs from the
    completion of one unit's last launch to the completion of the next unit's last launch, exactly K launches of steady
    state, whatever K is; with the record node IN the chain every unit pays one un-overlapped launch ramp (0.29 us per
    step at K = 20).  A graph holds U = clamp(6000 // K, 1, 60) units; as many graphs are captured as it takes to walk
    the whole pool."""
    pool, na = len(envs), len(actions)
    units = max(1, min(60, 6000 // steps))
    n_graphs = max(1, -(-pool // (units * steps)))
    graphs = []
    j = first
    side = torch.cuda.Stream(device=stream.device) if side_events else None
    for g in range(n_graphs):
        marks = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(units + 1)]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            def mark(ev):
                if side is None:
                    ev.record(stream)
                else:
                    side.wait_stream(stream)
                    ev.record(side)
            # one untimed launch first, so that the opening mark also sits behind a launch of the chain
            envs[j % pool].step(actions[j % na])
            j += 1
            mark(marks[0])
            for u in range(units):
                for _ in range(steps):
                    envs[j % pool].step(actions[j % na])
                    j += 1
                mark(marks[u + 1])
            if side is not None:
                stream.wait_stream(side)
        graphs.append((gr, list(zip(marks[:-1], marks[1:]))))
    return graphs


def time_units(torch, graphs, stream, repeats=REPEATS, spin_s=0.3, barrier=None, refresh=None):
    """untimed replays for spin_s seconds, then replays until `repeats` unit intervals have been collected (the action
    ring is redrawn before every replay, see time_replays); returns the per-unit seconds"""
    with torch.cuda.stream(stream):
        t_end = time.time() + spin_s
        i = 0
        while time.time() < t_end:
            if refresh:
                refresh()
            graphs[i % len(graphs)][0].replay()
            i += 1
            stream.synchronize()
    if barrier:
        barrier()
    out = []
    with torch.cuda.stream(stream):
        while len(out) < repeats:
            if refresh:
                refresh()
            gr, evs = graphs[i % len(graphs)]
            gr.replay()
            stream.synchronize()
            out.extend(a.elapsed_time(b) * 1e-3 for a, b in evs)
            i += 1
    if barrier:
        barrier()
    return np.array(out[:repeats])


def time_replays(torch, graphs, stream, repeats=REPEATS, spin_s=0.3, barrier=None, refresh=None):
    """untimed replays for spin_s seconds, then `repeats` replays (round-robin over the graphs), each bracketed by its
    own event pair on the launching stream; returns the per-replay seconds.  `refresh()` runs on the stream before
    every replay, OUTSIDE the event pairs: it redraws the action ring, so that no env ever sees the same action twice
    (a captured graph would otherwise hand batch b the same few action sets forever, and arms fed a periodic action
    sequence drift into a workspace corner where the IK is cheap -- the K = 20 vs K = 2000 gap of the first r02 run)."""
    with torch.cuda.stream(stream):
        t_end = time.time() + spin_s
        i = 0
        while time.time() < t_end:
            if refresh:
                refresh()
            graphs[i % len(graphs)].replay()
            i += 1
            if i % 8 == 0:
                stream.synchronize()
        stream.synchronize()
    if barrier:
        barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(repeats)]
    with torch.cuda.stream(stream):
        for r in range(repeats):
            if refresh:
                refresh()
            ev[r][0].record(stream)
            graphs[(i + r) % len(graphs)].replay()
            ev[r][1].record(stream)
    if barrier:
        barrier()
    else:
        stream.synchronize()
    return np.array([a.elapsed_time(b) * 1e-3 for a, b in ev])


def run_ours(args):
    import torch
    import torch.distributed as dist
    import drl_on_robot_arm_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = args.steps, max(args.warmup, 3)
    task, n = args.task, args.n_envs
    abytes = ALGO_BYTES[task]
    peak_gbs, peak_src = load_peaks()

    # pool of independent batches: touched state > 2 x L2 so each launch is HBM-cold
    pool = args.pool or int(np.ceil(2.0 * L2_BYTES / (abytes * n)))
    envs = [pkg.BatchedArmEnv(task, n_envs=n, device=dev, seed=0, auto_reset=True,
                              env_id_offset=(rank * pool + b) * n) for b in range(pool)]
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    # a ring of N_ACT independent U(-0.7,0.7) action sets (pre-generated [T,N,3]); launch k uses set k % N_ACT,
    # N_ACT coprime with the pool size so every env sees a different action at each of its steps (a constant action
    # would walk every arm into a workspace corner, where the IK needs its full 20 iterations)
    actions = (torch.rand((N_ACT, n, 3), device=dev, generator=gen) * 1.4 - 0.7).contiguous()
    stream = torch.cuda.Stream(device=dev)
    launches0 = sum(e.launch_count for e in envs)

    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        for k in range(max(warmup, 3)):                         # eager warm-up (also first-use init)
            envs[k % pool].step(actions[k % len(actions)])
    stream.synchronize()
    timed_units, side_events = True, os.environ.get("BENCH_INLINE_EVENTS", "0") != "1"
    try:
        graphs = capture_timed_units(torch, envs, actions, steps, stream, first=max(warmup, 3), side_events=side_events)
    except Exception:
        stream.synchronize()
        launches0 = sum(e.launch_count for e in envs) - max(warmup, 3)
        try:                   # boundary events in the launch chain instead of on a side branch
            side_events = False
            graphs = capture_timed_units(torch, envs, actions, steps, stream, first=max(warmup, 3), side_events=False)
        except Exception:      # no external-event support: time whole replays from outside instead
            launches0 = sum(e.launch_count for e in envs) - max(warmup, 3)
            graphs = capture_pool_graphs(torch, envs, actions, steps, stream, first=max(warmup, 3))
            timed_units = False
    cap_launches = sum(e.launch_count for e in envs) - launches0 - max(warmup, 3)
    assert (cap_launches - (len(graphs) if timed_units else 0)) % steps == 0      # (+ one untimed lead-in launch per unit graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler.start()
    # ---- timed region 1: device-resident inputs.  ~0.3 s of untimed replays (clock ramp, mid-episode states), then
    # REPEATS units of exactly K launches each, every unit timed by its own event pair, the set bracketed by a
    # barrier + synchronize; this rank's number is the MEDIAN unit (one ~60 us interval is at the mercy of a single
    # scheduling hiccup, which is what made round 1's scaling curve read 0.78).
    def redraw():
        actions.uniform_(-0.7, 0.7, generator=gen)              # fresh U(-0.7,0.7) actions for every replay (untimed)
    times = (time_units if timed_units else time_replays)(torch, graphs, stream, REPEATS, 0.3, barrier, refresh=redraw)
    sec = float(np.median(times))
    p10, p90 = float(np.percentile(times, 10)), float(np.percentile(times, 90))
    # ---- timed region 2: end to end through the host-buffer C-ABI call.  Every step: the [N,3] f32 actions sit in
    # pinned host memory, cross PCIe to the GPU, the fused kernel runs, obs/reward/done/success cross back into pinned
    # host memory and the call returns only when they are readable there (then one value of the result is read).
    e2e_steps = min(steps, args.e2e_steps)
    # the host owns a ring of 509 pre-drawn action sets (prime, ~25 MB) and WRITES the step's actions into the pinned block
    # every step, like a host-side policy would (the 48 KB memcpy is inside the timed region)
    host_ring = np.random.default_rng(7 + rank).uniform(-0.7, 0.7, (509, n, 3)).astype(np.float32)
    e2e_reps = max(5, min(REPEATS, int(2.0 / max(e2e_steps * 2.5e-5, 1e-6))))     # ~2 s of host stepping at most
    clock = time.perf_counter
    kk = 0
    acc = 0.0

    def run_e2e(group):
        """R repeats of a K-step host loop over the handles of `group` (round-robin); returns the per-repeat seconds spent
        inside the public call + the read of its result, and the per-repeat seconds including the host's action write"""
        nonlocal kk, acc
        bufs = [e.host_buffers() for e in group]
        for k in range(max(warmup, 3) + len(group)):
            bufs[kk % len(group)][0][:] = host_ring[kk % 509]
            group[kk % len(group)].step_pinned()
            kk += 1
        barrier()
        inside, incl = [], []
        for r in range(e2e_reps):
            t_rep = clock()
            in_call = 0.0
            for k in range(e2e_steps):
                b = kk % len(group)
                bufs[b][0][:] = host_ring[kk % 509]           # the host-side "policy" puts this step's actions into pinned memory
                t0 = clock()
                rew = group[b].step_pinned()[1]               # the public call: actions cross PCIe, kernel, results cross back
                acc += float(rew[0])                          # the step's result is consumed on the host
                in_call += clock() - t0
                kk += 1
            inside.append(in_call)
            incl.append(clock() - t_rep)
        torch.cuda.synchronize(dev)
        barrier()
        return inside, incl

    # (a) launch per step: every call replays the one-kernel graph; 64 handles in rotation
    e2e_pool = min(pool, 64)
    launch_times, launch_incl = run_e2e(envs[:e2e_pool])
    # (b) resident step server (armsim_host_server, a public switch of the handle): ONE kernel per handle stays on the
    # GPU and serves a step per command word, so the call is a release store + a doorbell poll.  4 handles in rotation
    # (each resident kernel holds 32 blocks; their 2 MB of state stay in L2, as the 64 x 0.5 MB of (a) do).
    server_group = envs[:min(pool, 4)]
    e2e_mode = "launch_per_step"
    e2e_times, e2e_incl = launch_times, launch_incl
    if os.environ.get("BENCH_NO_SERVER", "0") != "1" and task in ("reach", "push", "pick", "kuka_reach"):
        srv_times = None
        ok = 1.0
        try:
            for e in server_group:
                e.host_server(20000)
        except Exception as ex:      # the launch path stays the number
            ok = 0.0
            print("resident step server unavailable: %r" % (ex,), file=sys.stderr)
        if world > 1:                # every rank takes the same legs (run_e2e contains barriers)
            w = torch.tensor([ok], device=dev, dtype=torch.float64)
            dist.all_reduce(w, op=dist.ReduceOp.MIN)
            ok = float(w)
        if ok > 0.0:
            srv_times, srv_incl = run_e2e(server_group)
        worse = 1.0 if (srv_times is None or np.median(srv_times) >= np.median(launch_times)) else 0.0
        if world > 1:                # ... and reports the same mode
            w = torch.tensor([worse], device=dev, dtype=torch.float64)
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            worse = float(w)
        if worse == 0.0:
            e2e_mode, e2e_times, e2e_incl = "resident_server", srv_times, srv_incl
    e2e_sec = float(np.median(e2e_times))
    bufs = [e.host_buffers() for e in envs[:2]]
    # ---- secondary: the same end-to-end step as a depth-2 pipeline over two independent 4096-env groups
    # (armsim_step_host_async / _wait, gym.vector's step_async / step_wait): group B's launch + PCIe round trip is in
    # flight while the host consumes group A's results.  NOT the headline: twice the envs are live at any time.
    pipe_sec = None
    if pool >= 2:
        ea, eb = envs[0], envs[1]
        for k in range(6):
            ea.step_async(); eb.step_async(); ea.step_wait(); eb.step_wait()
        pipe_steps = max(e2e_steps, 200)
        t0 = time.perf_counter()
        bufs[0][0][:] = host_ring[kk % 509]; kk += 1
        ea.step_async()
        for k in range(pipe_steps // 2):
            bufs[1][0][:] = host_ring[kk % 509]; kk += 1
            eb.step_async()
            acc += float(ea.step_wait()[1][0])
            bufs[0][0][:] = host_ring[kk % 509]; kk += 1
            ea.step_async()
            acc += float(eb.step_wait()[1][0])
        ea.step_wait()
        pipe_sec = (time.perf_counter() - t0) / (2 * (pipe_steps // 2) + 1)
    for e in server_group:
        e.host_server(0)
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([sec, e2e_sec, p10, p90], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec, e2e_sec, p10, p90 = (float(x) for x in t)
    value = world * n * steps / sec
    e2e_value = world * n * e2e_steps / e2e_sec
    launch_us = 1e6 * sec / steps
    achieved = abytes * n / (sec / steps) / 1e9
    h2d = n * 3 * 4
    d2h = n * envs[0].obs_dim * 4 + n * 4 + n + n

    extra = {}
    if rank == 0 and world == 1 and not args.quick:          # single-GPU secondary numbers; the scaling runs skip them
        extra = side_measurements(torch, pkg, dev, peak_gbs)
    cpu = cpu_baseline() if (rank == 0 and world == 1 and not args.no_cpu) else None
    train_update = None
    if not args.quick:
        try:
            train_update = measure_train_updates(torch, dist, dev, world, rank)      # every rank takes part (collectives)
        except Exception as e:
            train_update = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "rl_%s_env N_envs=%d fused-step kernel, %d x B200 (%d envs total)" % (task, n, world, world * n),
                       "task": task, "n_envs_per_gpu": n, "robot": "kuka_iiwa", "mode": "ik_teleport", "mapping": envs[0].mapping,
                       "actions": "U(-0.7,0.7) [%d,N,3] f32 ring on device, launch k uses set k mod %d, the ring is REDRAWN before every "
                                  "timed replay (outside the event pair): a random walk, no env ever repeats an action; auto-reset in kernel" % (N_ACT, N_ACT),
                       "l2": "inputs larger than L2: round-robin over a pool of %d independent %d-env batches "
                             "(%.0f MB touched state, L2 = 126 MB), every launch HBM-cold" % (pool, n, pool * abytes * n / 1e6),
                       "launch": ("CUDA graphs of timed units, one unit = exactly K fused-step launches between two event-record nodes "
                                  "(%s); %d graph(s) x %d units covering the pool"
                                  % ("records on a side branch of the graph: the launches keep their kernel-to-kernel chain across unit "
                                     "boundaries" if side_events else "records in the launch chain", len(graphs), len(graphs[0][1]))) if timed_units else
                                 ("CUDA graphs of exactly K fused-step launches each (%d graphs covering the pool, replayed round-robin), "
                                  "CUDA event pair per replay on the launching stream" % len(graphs))},
            "timing": {"repeats": REPEATS, "statistic": "median over the timed replays, max over ranks",
                       "p10_ms_per_step": 1e3 * p10 / steps, "p90_ms_per_step": 1e3 * p90 / steps,
                       "e2e_repeats": e2e_reps, "e2e_p10_us_per_step": 1e6 * float(np.percentile(e2e_times, 10)) / e2e_steps,
                       "e2e_p90_us_per_step": 1e6 * float(np.percentile(e2e_times, 90)) / e2e_steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": load_traffic(task, n), "traffic_unit": "bytes per launch (dram read + write, ncu --set full, "
                         "profiles/r02_ncu_summary.json; writes still in L2 at kernel end are not counted by ncu)",
                         "peak_source": peak_src, "algorithmic_bytes_per_env_step": abytes,
                         "kernel": "step_%s_kernel<%s>" % (envs[0].mapping, task), "avg_launch_us": launch_us,
                         "note": "kernel is fp32-issue / dependent-latency bound, not HBM bound (SURVEY 7): ~6 kFLOP of "
                                 "dependent fp32 per 118 B; at N=4096 every warp sits alone on an SM sub-partition (one FP32 "
                                 "instruction per ~2 cycles, tools/micro/ffma_rate.cu) and the launch waits for the slowest "
                                 "arm's IK (3 DLS iterations typical); see other_configs for the multi-wave sizes"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "incl_host_action_write": world * n * e2e_steps / float(np.median(e2e_incl)),
                    "timed": "time inside the public call + the host read of its result, summed over the K steps of a repeat (median of "
                             "the repeats); the host-side policy writing the NEXT 48 KB of actions into the pinned block sits between "
                             "calls and is reported separately (incl_host_action_write)",
                    "mode": e2e_mode,
                    "api": ("ArmSimHandle.host_server(20000) once, then ArmSimHandle.step_pinned -> armsim_step_host on the handle's "
                            "pinned host block (armsim_host_buffers): a RESIDENT kernel (armsim_host_server) polls the step's command "
                            "word, reads the actions / writes the results over PCIe and rings per-block doorbells; the call itself is "
                            "a release store + a doorbell poll, no CUDA API call" if e2e_mode == "resident_server" else
                            "ArmSimHandle.step_pinned -> armsim_step_host on the handle's pinned host block "
                            "(armsim_host_buffers): graph-replayed kernel reads actions / writes results over PCIe, per-block doorbells"),
                    "launch_per_step": {"value": world * n * e2e_steps / float(np.median(launch_times)), "unit": UNIT,
                                        "note": "the same call without the resident server: one graph-replayed launch per step"},
                    "pipelined_depth2": None if pipe_sec is None else
                    {"value": world * n / pipe_sec, "unit": UNIT, "note": "secondary: step_async/step_wait over two independent "
                     "%d-env groups per GPU (same per-step H2D/D2H bytes); the headline e2e above is the synchronous call" % n}},
            "gpu_launches": steps,
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if train_update is not None:
            line["train_update"] = train_update
        line.update(extra)
        print(json.dumps(line), flush=True)
    for e in envs:
        e.close()
    if world > 1:
        from drl_on_robot_arm_b200 import distributed as D
        D.shutdown()


def measure_train_updates(torch, dist, dev, world, rank):
    """BASELINE config 5 in front of the driver: microseconds per learning update (replay.sample + agent.train) at N
    ranks INCLUDING the flat-bucket gradient all-reduce (distributed.GradBucket: one NCCL all-reduce per network per
    optimizer step, the only collective of the training path; algo/DARC/DARC_mlp.py:181-203, TD3_mlp.py:147-157),
    issued eagerly and replayed from a CUDA graph that contains the all-reduce, plus a replica-identity bit: after the
    same updates every rank must hold bit-identical networks (checksums all-gathered)."""
    from drl_on_robot_arm_b200 import train
    out = {}
    for algo in ("TD3_MLP", "DARC_MLP"):
        entry = {}
        for mode, graphed in (("eager", False), ("graphed", True)):
            tr = train.make_trainer(task="reach", algo=algo, n_envs=1024, device=dev, seed=0, window=1024, sync_every=10 ** 9,
                                    minimal_episodes=10 ** 12, graph_updates=graphed)
            for _ in range(540):
                tr.rollout_step()                  # every env has committed its first (<= 501-step) episode
            tr._stream.synchronize()
            cyc = tr.agent.update_cycle()[0]
            tr.train_updates(12 * cyc)             # warm-up (+ graph capture)
            tr._stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            k = 40 * cyc
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(tr._stream)
            tr.train_updates(k)
            e1.record(tr._stream)
            tr._stream.synchronize()
            wall = time.perf_counter() - t0
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / k, wall * 1e6 / k], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            entry[mode + "_us_per_update"] = float(t[0])
            entry[mode + "_us_per_update_wall"] = float(t[1])
            flat = torch.cat([p.detach().reshape(-1) for l, _ in tr.agent._learners() for p in l.net.parameters()])
            chk = flat.view(torch.int32).to(torch.int64).sum().reshape(1)          # bit-level checksum of every weight
            same = True
            if world > 1:
                got = [torch.empty_like(chk) for _ in range(world)]
                dist.all_gather(got, chk)
                same = all(int(g) == int(got[0]) for g in got)
            entry[mode + "_replicas_identical"] = bool(same)
            entry["allreduce_bytes_per_update"] = int(sum(l.bucket.nbytes for l, _ in tr.agent._learners() if l.bucket is not None))
            tr.release_graphs()             # recorded NCCL work must be gone before the process group is (distributed.shutdown)
            tr.env.close(); tr.replay.close()
            del tr
        out[algo] = entry
    out["what"] = ("us per {replay.sample(256) + agent.train} at %d rank(s), 1024 reach envs per rank, device-timed on the trainer's "
                   "stream (max over ranks); multi-rank updates include one NCCL all-reduce of the flat gradient bucket per "
                   "network per optimizer step; graphed = CUDA-graph replays of one control-flow cycle, all-reduce recorded "
                   "inside the graph" % world)
    return out


def side_measurements(torch, pkg, dev, peak_gbs):
    """secondary numbers (not the headline): the DRAM-honest large-N sweep and the push / pick kernels, measured the
    same way as the headline (CUDA graph of K launches over a pool of batches whose state exceeds 2 x L2)"""
    out = {}
    try:
        res = {}
        stream = torch.cuda.Stream(device=dev)
        for task, n in (("reach", 1 << 20), ("reach", 1 << 22), ("push", N_ENVS), ("pick", 2048), ("reach", 32768),
                        ("push", 1 << 20), ("pick", 1 << 20), ("reach_torque", N_ENVS), ("reach_torque", 1 << 20)):
            torque = task.endswith("_torque")
            pool = max(1, int(np.ceil(2.0 * L2_BYTES / (ALGO_BYTES[task] * n))))
            k = 20 * pool if n >= (1 << 20) else 600
            envs = [pkg.BatchedArmEnv(task.split("_")[0], n_envs=n, device=dev, seed=0, auto_reset=True, env_id_offset=b * n,
                                      mode="torque" if torque else "ik_teleport") for b in range(pool)]
            na = 7 if n >= (1 << 20) else N_ACT
            if pool % na == 0:
                na -= 1                                          # keep the ring out of step with the pool
            k = max(na, k // na * na)                            # whole turns of the action ring per replay
            if torque:
                a = (torch.rand((na, n, 7), device=dev) * 2.0 - 1.0) * 30.0      # joint torques, N m (effort limit 300)
            else:
                a = (torch.rand((na, n, 3), device=dev) * 1.4 - 0.7)
                if task != "reach":
                    a *= 0.4 / 0.7                              # action_bound 0.4 for push / pick (main.py:457,526)
            with torch.cuda.stream(stream):
                for j in range(max(3, min(pool, 8))):
                    envs[j % pool].step(a[j % na])
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for j in range(k):
                    envs[j % pool].step(a[j % na])
            # one graph already spans the whole pool here (k >= pool): median of 15 timed replays after 0.25 s untimed
            lim = 30.0 if torque else (0.7 if task == "reach" else 0.4)

            def redraw(a=a, lim=lim):
                a.uniform_(-lim, lim)
            tt = time_replays(torch, [g], stream, repeats=15, spin_s=0.25, refresh=redraw)
            sec = float(np.median(tt)) / k
            gbs = ALGO_BYTES[task] * n / sec / 1e9
            res["%s_n%d" % (task, n)] = {"env_steps_per_s": n / sec, "us_per_launch": sec * 1e6, "achieved_gbs": gbs,
                                         "hbm_frac": gbs / peak_gbs, "pool": pool, "launches": k, "repeats": 15,
                                         "p10_us": float(np.percentile(tt, 10)) / k * 1e6, "p90_us": float(np.percentile(tt, 90)) / k * 1e6,
                                         "l2": "pool state %.0f MB > 2 x L2, CUDA graph" % (pool * ALGO_BYTES[task] * n / 1e6)}
            del g
            for e in envs:
                e.close()
        out["other_configs"] = res
        # "a second number with the real TD3 actor in the loop (CUDA-graph captured)": one lockstep rollout
        # step = actor forward + exploration noise + fused env step + replay store + episode statistics, no learning
        from drl_on_robot_arm_b200 import train
        tr = train.make_trainer(task="reach", algo="TD3_MLP", n_envs=N_ENVS, device=dev, seed=0, minimal_episodes=10 ** 12,
                                sync_every=10 ** 9)
        tr.sync_every = 50                              # chunk graphs of 50 lockstep steps (VectorTrainer.rollout_chunk)
        for _ in range(50):
            tr.rollout_step()
        done_steps = 0
        while done_steps < 200:                         # warm-up incl. the chunk-graph capture
            done_steps += tr.rollout_chunk(10 ** 9)
        tr._stream.synchronize()
        k = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tr._stream)
        while k < 2000:
            k += tr.rollout_chunk(10 ** 9)
        e1.record(tr._stream)
        tr._stream.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / k
        out["rollout_with_td3_actor"] = {"env_steps_per_s": N_ENVS / sec, "us_per_rollout_step": sec * 1e6, "n_envs": N_ENVS,
                                         "what": "CUDA-graph replays (50 lockstep steps per graph) of {TD3 actor forward + N(0,0.98) exploration noise as ONE "
                                                 "launch (armsim_policy_act on the PyTorch parameters), fused reach step, trajectory-replay store, "
                                                 "armsim_track_episodes}, device-timed"}
        tr.env.close(); tr.replay.close()
    except Exception as e:  # secondary numbers must never kill the headline line
        out.setdefault("other_configs", {})["error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default="reach", choices=list(ALGO_BYTES))
    ap.add_argument("--n-envs", type=int, default=N_ENVS)
    ap.add_argument("--pool", type=int, default=0, help="independent batches in rotation (0 = enough for 2x L2)")
    ap.add_argument("--e2e-steps", type=int, default=500)
    ap.add_argument("--quick", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
