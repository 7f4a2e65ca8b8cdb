import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drl_on_robot_arm_b200 as pkg
from drl_on_robot_arm_b200 import _lib as L
task = sys.argv[1] if len(sys.argv) > 1 else "reach"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
env = pkg.BatchedArmEnv(task, n_envs=n, device=dev, seed=0, auto_reset=True)
sc = 0.7 if task == "reach" else 0.4
acts = (torch.rand((61, n, 3), device=dev) * 2 - 1) * sc
hist = np.zeros(21, int)
for k in range(1100):
    env.step(acts[k % 61])
    if k % 50 == 49 or k in (0, 1, 2):
        it = env.get_state(L.F_IK_ITERS).astype(int).ravel()
        wmax = it.reshape(-1, 32).max(axis=1)
        hist += np.bincount(it, minlength=21)
        ee = env.obs[:, :3].cpu().numpy()
        print("step %4d iters mean %.2f max %2d  n>=4: %3d n20 %3d  warp-max mean %.2f  | ee z<0.02: %d  x>0.68: %d" %
              (k, it.mean(), it.max(), (it >= 4).sum(), (it == 20).sum(), wmax.mean(), (ee[:, 2] < 0.02).sum(), (ee[:, 0] > 0.68).sum()))
print({i: int(h) for i, h in enumerate(hist) if h})
