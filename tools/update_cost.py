"""time per learning update (replay.sample + agent.train) graphed vs eager, and per rollout step, on one GPU"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drl_on_robot_arm_b200 import train
algo = sys.argv[1] if len(sys.argv) > 1 else "TD3_MLP"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
for graphed in (True, False):
    tr = train.make_trainer(task="reach", algo=algo, n_envs=n, device="cuda:0", seed=0, window=1024, sync_every=10 ** 9,
                            minimal_episodes=10 ** 12, graph_updates=graphed)
    for _ in range(600):
        tr.rollout_step()                  # fill the replay with > 1 episode per env
    tr._stream.synchronize()
    tr.train_updates(30)                   # warm-up (+ graph capture)
    tr._stream.synchronize()
    k = 300
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tr._stream)
    tr.train_updates(k)
    e1.record(tr._stream)
    tr._stream.synchronize()
    wall = time.perf_counter() - t0
    print("%s n=%d graphed=%s: %.1f us per update on the device, %.1f us wall" % (algo, n, graphed, e0.elapsed_time(e1) * 1e3 / k, wall * 1e6 / k), flush=True)
    t0 = time.perf_counter()
    for _ in range(500):
        tr.rollout_step()
    tr._stream.synchronize()
    print("   rollout step: %.1f us wall" % ((time.perf_counter() - t0) * 1e6 / 500), flush=True)
    tr.env.close(); tr.replay.close()
