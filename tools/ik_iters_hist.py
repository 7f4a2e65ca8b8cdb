"""histogram of DLS iteration counts per env-step on the device vs the oracle, free-running from reset"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drl_on_robot_arm_b200 as pkg
from drl_on_robot_arm_b200 import _lib as L
from oracle import oracle as O
task = sys.argv[1] if len(sys.argv) > 1 else "pick"
tid = {"reach": O.TASK_REACH, "push": O.TASK_PUSH, "pick": O.TASK_PICK}[task]
n = 2048
env = pkg.ArmSimHandle(task, n_envs=n, seed=0, auto_reset=True)
ora = O.OracleSim(O.default_config(tid, n_envs=n, seed=0, auto_reset=1))
env.reset_host(); ora.reset()
rng = np.random.default_rng(0)
hg = np.zeros(21, int); ho = np.zeros(21, int)
for k in range(60):
    a = rng.uniform(-0.4, 0.4, (n, 3)).astype(np.float32)
    env.step_host(a); ora.step(a)
    ig = env.get_state(L.F_IK_ITERS).astype(int).ravel(); io = ora.get_state(O.F_IK_ITERS).astype(int).ravel()
    if k >= 10:
        hg += np.bincount(ig, minlength=21); ho += np.bincount(io, minlength=21)
    if k in (0, 1, 5, 20, 40):
        print("step", k, "device mean %.2f max %d n20 %d | oracle mean %.2f max %d n20 %d" % (ig.mean(), ig.max(), (ig == 20).sum(), io.mean(), io.max(), (io == 20).sum()))
print("device", {i: int(h) for i, h in enumerate(hg) if h})
print("oracle", {i: int(h) for i, h in enumerate(ho) if h})
q = env.get_state(L.F_Q)
print("q6 range", q[:, 6].min(), q[:, 6].max(), "q abs max", np.abs(q).max())
