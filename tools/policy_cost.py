"""python tools/policy_cost.py [n_envs] : device time of the acting policy per lockstep step, one fused launch
(armsim_policy_act: MLP forward + exploration noise) against the PyTorch module + armsim_explore, both replayed from a
CUDA graph of 200 calls."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import drl_on_robot_arm_b200 as pkg
from drl_on_robot_arm_b200.algo.nets import PolicyNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = pkg.BatchedArmEnv("reach", n_envs=n, device="cuda:0", seed=0)
net = PolicyNet(6, 256, 3, 0.7).to("cuda:0")
obs = torch.rand((n, 6), device="cuda")
out = torch.empty((n, 3), device="cuda")
st = torch.cuda.Stream()


def fused():
    env.policy_act(net, obs, noise_std=0.98, out=out)


@torch.no_grad()
def eager():
    env.explore(net(obs), 0.98, out=out)


for name, fn in (("fused armsim_policy_act", fused), ("torch PolicyNet + armsim_explore", eager)):
    with torch.cuda.stream(st):
        for _ in range(5):
            fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(200):
                fn()
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(st); g.replay(); e1.record(st); st.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / 200)
    print("%-36s n=%d  %.2f us per call" % (name, n, best), flush=True)
