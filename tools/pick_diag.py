import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drl_on_robot_arm_b200 as pkg
from drl_on_robot_arm_b200 import _lib as L
task = sys.argv[1] if len(sys.argv) > 1 else "pick"
n = 2048
dev = torch.device("cuda:0")
env = pkg.BatchedArmEnv(task, n_envs=n, device=dev, seed=0, auto_reset=True)
acts = (torch.rand((64, n, 3), device=dev) * 1.4 - 0.7) * (0.4 / 0.7)
for k in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); env.step(acts[k % 64]); e1.record(); torch.cuda.synchronize()
    it = env.get_state(L.F_IK_ITERS).astype(int).ravel()
    grip = env.get_state(L.F_GRIP).ravel() if task == "pick" else np.zeros(1)
    cz = env.get_state(L.F_CUBE_POS)[:, 2]
    wmax = it.reshape(-1, 32).max(axis=1)
    if k < 12 or k % 5 == 0:
        print("step %2d  %.1f us  iters mean %.2f max %d n20 %d  warp-max mean %.2f | grip>0: %d  cube z [%.4f, %.4f]" %
              (k, e0.elapsed_time(e1) * 1e3, it.mean(), it.max(), (it == 20).sum(), wmax.mean(), (grip > 0).sum(), cz.min(), cz.max()))
