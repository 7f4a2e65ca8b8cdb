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
(a constant action per env would pin the arms in workspace corners).  `value` is device-timed (CUDA events on the
launching stream around a CUDA-graph replay of exactly K launches, after ~0.3 s of untimed replays that bring the
clocks up and the envs into mid-episode states; max over ranks); `e2e` is the same metric through the host-buffer
C-ABI call (armsim_step_host on the handle's pinned block: the kernel reads the actions from / writes the results
to host memory over PCIe every step, the host polls per-block doorbells), timed on the host clock.
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
    p = os.path.join(ROOT, "profiles", "r01_ncu_summary.json")
    try:
        d = json.load(open(p))["v5_%s_n%d" % (task, n)]
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
def cpu_baseline(task="reach", n_envs=N_ENVS, target_seconds=6.0, threads=None):
    """The CPU restatement of the reference step (oracle/, C, fp64) on the host cores: the reference's own step is
    Python on pybullet==3.0.6, which cannot be installed here -> kind = "port"."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.build()
    tid = {"reach": O.TASK_REACH, "push": O.TASK_PUSH, "pick": O.TASK_PICK}[task]
    cores = threads or os.cpu_count() or 1
    rng = np.random.default_rng(1)
    acts = rng.uniform(-0.7, 0.7, (8, n_envs, 3)).astype(np.float32)

    def run(nthreads, seconds):
        sim = O.OracleSim(O.default_config(tid, n_envs=n_envs, seed=0, auto_reset=1))
        sim.reset()
        bounds = np.linspace(0, n_envs, nthreads + 1).astype(int)
        lib = O.lib()
        pool = ThreadPoolExecutor(nthreads) if nthreads > 1 else None

        def one(k):
            a = acts[k % len(acts)]
            if pool is None:
                sim.step(a)
            else:
                futs = [pool.submit(lib.orc_step_range, sim.h, int(bounds[i]), int(bounds[i + 1]), a.ctypes.data,
                                    sim.obs.ctypes.data, sim.reward.ctypes.data, sim.done.ctypes.data,
                                    sim.success.ctypes.data) for i in range(nthreads)]
                for f in futs:
                    f.result()
        one(0)
        t0 = time.perf_counter()
        k = 0
        while True:
            one(k)
            k += 1
            dt = time.perf_counter() - t0
            if dt >= seconds:
                break
        if pool:
            pool.shutdown()
        sim.close()
        return n_envs * k / dt, k

    one_core, k1 = run(1, target_seconds / 3)
    all_core, kc = run(cores, target_seconds * 2 / 3) if cores > 1 else (one_core, k1)
    return {"value": all_core, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d envs x %d steps of the C fp64 restatement of RLReachEnv.step (oracle/), %d threads; "
                      "PyBullet itself is not installable offline" % (n_envs, kc, cores),
            "single_core_value": one_core}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    n = N_ENVS
    # each "step" = one Env.step of all 4096 envs; bound the run to ~60 s whatever K the driver passes
    sim = O.OracleSim(O.default_config(O.TASK_REACH, n_envs=n, seed=0, auto_reset=1))
    sim.reset()
    rng = np.random.default_rng(1)
    acts = rng.uniform(-0.7, 0.7, (8, n, 3)).astype(np.float32)
    bounds = np.linspace(0, n, cores + 1).astype(int)
    lib = O.lib()
    pool = ThreadPoolExecutor(cores)

    def one(k, lo_hi=None):
        a = acts[k % len(acts)]
        futs = [pool.submit(lib.orc_step_range, sim.h, int(bounds[i]), int(bounds[i + 1]), a.ctypes.data,
                            sim.obs.ctypes.data, sim.reward.ctypes.data, sim.done.ctypes.data, sim.success.ctypes.data)
                for i in range(cores)]
        for f in futs:
            f.result()
    t0 = time.perf_counter()
    one(0)
    per_step = time.perf_counter() - t0
    budget = 60.0
    k_eff = max(1, min(steps, int(budget / max(per_step, 1e-6))))
    w_eff = max(1, min(warm, max(1, k_eff // 10)))
    for k in range(w_eff):
        one(k)
    t0 = time.perf_counter()
    for k in range(k_eff):
        one(k)
    dt = time.perf_counter() - t0
    val = n * k_eff / dt
    sample = ("%d Env.step batches of %d envs (of the %d requested; bounded to ~%.0f s), C fp64 restatement of "
              "RLReachEnv.step on %d host threads" % (k_eff, n, steps, budget, cores))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k_eff,
            "warmup": w_eff, "ms_per_step": 1e3 * dt / k_eff, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "rl_reach_env N_envs=4096, Env.step on the CPU (reference path restated in C; "
                                   "pybullet==3.0.6 unavailable offline)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def time_graph(torch, envs, actions, steps, warmup, stream):
    """CUDA-graph capture of `steps` back-to-back fused-step launches over the pool; returns seconds (device time)."""
    pool = len(envs)

    def launch_range(lo, hi):
        for k in range(lo, hi):
            envs[k % pool].step(actions[k % len(actions)])
    with torch.cuda.stream(stream):
        launch_range(0, max(warmup, 3))                       # eager warm-up (also first-use init)
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        launch_range(0, steps)
    return g


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
    graph = time_graph(torch, envs, actions, steps, warmup, stream)
    cap_launches = sum(e.launch_count for e in envs) - launches0 - max(warmup, 3)
    assert cap_launches == steps
    with torch.cuda.stream(stream):
        graph.replay()                                          # untimed replay: graph upload, caches
    stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler.start()
    # keep the GPU busy long enough for the clock sampler to see the load (not timed)
    t_end = time.time() + 0.3
    with torch.cuda.stream(stream):
        while time.time() < t_end:
            graph.replay()
            stream.synchronize()
    # ---- timed region 1: device-resident inputs, exactly K launches
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        graph.replay()
        ev1.record(stream)
    barrier()
    sec = ev0.elapsed_time(ev1) * 1e-3
    # ---- timed region 2: end to end through the host-buffer C-ABI call.  Every step: the [N,3] f32 actions sit in
    # pinned host memory, cross PCIe to the GPU, the fused kernel runs, obs/reward/done/success cross back into pinned
    # host memory and the call returns only when they are readable there (then one value of the result is read).
    e2e_steps = min(steps, args.e2e_steps)
    e2e_pool = min(pool, 64)
    bufs = []
    for b in range(e2e_pool):
        hb = envs[b].host_buffers()
        hb[0][:] = actions[b % len(actions)].cpu().numpy()
        bufs.append(hb)
    for k in range(max(warmup, 3) + e2e_pool):
        envs[k % e2e_pool].step_pinned()
    barrier()
    acc = 0.0
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        rew = envs[k % e2e_pool].step_pinned()[1]
        acc += float(rew[0])                                  # the step's result is consumed on the host
    torch.cuda.synchronize(dev)
    e2e_sec = time.perf_counter() - t0
    barrier()
    # ---- secondary: the same end-to-end step as a depth-2 pipeline over two independent 4096-env groups
    # (armsim_step_host_async / _wait, gym.vector's step_async / step_wait): group B's launch + PCIe round trip is in
    # flight while the host consumes group A's results.  NOT the headline: twice the envs are live at any time.
    pipe_sec = None
    if e2e_pool >= 2:
        ea, eb = envs[0], envs[1]
        for k in range(6):
            ea.step_async(); eb.step_async(); ea.step_wait(); eb.step_wait()
        t0 = time.perf_counter()
        ea.step_async()
        for k in range(e2e_steps // 2):
            eb.step_async()
            acc += float(ea.step_wait()[1][0])
            ea.step_async()
            acc += float(eb.step_wait()[1][0])
        ea.step_wait()
        pipe_sec = (time.perf_counter() - t0) / (2 * (e2e_steps // 2) + 1)
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([sec, e2e_sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec, e2e_sec = float(t[0]), float(t[1])
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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "rl_%s_env N_envs=%d fused-step kernel, %d x B200 (%d envs total)" % (task, n, world, world * n),
                       "task": task, "n_envs_per_gpu": n, "robot": "kuka_iiwa", "mode": "ik_teleport", "mapping": "lane",
                       "actions": "pre-generated U(-0.7,0.7) [%d,N,3] f32 on device (launch k uses set k mod %d), auto-reset in kernel" % (N_ACT, N_ACT),
                       "l2": "inputs larger than L2: round-robin over a pool of %d independent %d-env batches "
                             "(%.0f MB touched state, L2 = 126 MB), every launch HBM-cold" % (pool, n, pool * abytes * n / 1e6),
                       "launch": "CUDA graph of exactly K fused-step launches, CUDA events on the launching stream"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": load_traffic(task, n), "traffic_unit": "bytes per launch (dram read + write, ncu --set full, "
                         "profiles/r01_ncu_summary.json; writes still in L2 at kernel end are not counted by ncu)",
                         "peak_source": peak_src, "algorithmic_bytes_per_env_step": abytes,
                         "kernel": "step_lane_kernel<%s>" % task, "avg_launch_us": launch_us,
                         "note": "kernel is fp32-issue / dependent-latency bound, not HBM bound (SURVEY 7): ~6 kFLOP of "
                                 "dependent fp32 per 118 B; at N=4096 128 warps on 592 SM sub-partitions wait for the "
                                 "slowest arm's IK (3 DLS iterations typical); see other_configs for the multi-wave sizes"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "ArmSimHandle.step_pinned -> armsim_step_host on the handle's pinned host block "
                           "(armsim_host_buffers): graph-replayed kernel reads actions / writes results over PCIe, per-block doorbells",
                    "pipelined_depth2": None if pipe_sec is None else
                    {"value": world * n / pipe_sec, "unit": UNIT, "note": "secondary: step_async/step_wait over two independent "
                     "%d-env groups per GPU (same per-step H2D/D2H bytes); the headline e2e above is the synchronous call" % n}},
            "gpu_launches": steps,
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        line.update(extra)
        print(json.dumps(line), flush=True)
    for e in envs:
        e.close()
    if world > 1:
        dist.destroy_process_group()


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
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                t_end = time.time() + 0.25                       # untimed: past the reset transient, clocks ramped up
                while time.time() < t_end:
                    g.replay()
                    stream.synchronize()
                e0.record(stream)
                g.replay()
                e1.record(stream)
            stream.synchronize()
            sec = e0.elapsed_time(e1) * 1e-3 / k
            gbs = ALGO_BYTES[task] * n / sec / 1e9
            res["%s_n%d" % (task, n)] = {"env_steps_per_s": n / sec, "us_per_launch": sec * 1e6, "achieved_gbs": gbs,
                                         "hbm_frac": gbs / peak_gbs, "pool": pool, "launches": k,
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
        for _ in range(20):
            tr.rollout_step()
        tr._stream.synchronize()
        k = 1000
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tr._stream)
        for _ in range(k):
            tr.rollout_step()
        e1.record(tr._stream)
        tr._stream.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / k
        out["rollout_with_td3_actor"] = {"env_steps_per_s": N_ENVS / sec, "us_per_rollout_step": sec * 1e6, "n_envs": N_ENVS,
                                         "what": "CUDA-graph replay of {TD3 actor forward (PyTorch), armsim_explore N(0,0.98) noise, fused reach "
                                                 "step, trajectory-replay store, armsim_track_episodes} per lockstep step, device-timed"}
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
