#!/usr/bin/env python
"""Generate tests/golden/algo_<NAME>.pt by running the REFERENCE's MLP agents (/root/reference/algo/*/_mlp.py) on the
CPU for a few updates on fixed batches: initial weights, the batches, the per-call losses, the actions chosen and the
final weights of every network (incl. targets).  Small nets (hidden_dim 32, batch 16) keep the fixtures at ~100 KB.
Run in the BUILD container only; tests/test_algo_parity.py loads the fixtures, so nothing reads the reference later."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ARMSIM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import algo as RA  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NETS = {"DDPG_MLP": ["actor", "critic", "target_actor", "target_critic"],
        "TD3_MLP": ["actor", "critic", "target_actor", "target_critic"],
        "DADDPG_MLP": ["actor1", "actor2", "critic", "target_actor1", "target_actor2", "target_critic"],
        "DATD3_MLP": ["actor1", "actor2", "critic1", "critic2", "target_actor1", "target_actor2", "target_critic1", "target_critic2"],
        "DARC_MLP": ["actor1", "actor2", "critic1", "critic2", "target_actor1", "target_actor2", "target_critic1", "target_critic2"]}


def main():
    S, A, H, B, K = 6, 3, 32, 16, 7
    for name, nets in NETS.items():
        torch.manual_seed(7)
        agent = getattr(RA, name)(state_dim=S, action_dim=A, action_bound=0.7, hidden_dim=H, device=torch.device("cpu"))
        init = {n: {k: v.clone() for k, v in getattr(agent, n).state_dict().items()} for n in nets}
        rng = np.random.default_rng(3)
        batches, losses = [], []
        for k in range(K):
            b = dict(states=rng.uniform(-1, 1, (B, S)).astype(np.float32), actions=rng.uniform(-0.7, 0.7, (B, A)).astype(np.float32),
                     rewards=rng.uniform(-5, 1, B).astype(np.float32), next_states=rng.uniform(-1, 1, (B, S)).astype(np.float32),
                     dones=(rng.uniform(size=B) < 0.2).astype(np.float32))
            batches.append(b)
            torch.manual_seed(100 + k)                          # pins the target-smoothing noise of this call
            out = agent.train({kk: vv.copy() for kk, vv in b.items()})
            losses.append(None if out is None else float(out))
        probe = rng.uniform(-1, 1, (5, S)).astype(np.float32)
        acts = np.stack([agent.take_action(p) for p in probe])
        final = {n: {k: v.clone() for k, v in getattr(agent, n).state_dict().items()} for n in nets}
        path = os.path.join(ROOT, "tests", "golden", "algo_%s.pt" % name)
        torch.save(dict(dims=(S, A, H, B, K), init=init, final=final, batches=batches, losses=losses, probe=probe, actions=acts,
                        total_it=agent.total_it if hasattr(agent, "total_it") else None), path)
        print(name, "->", path, os.path.getsize(path), "bytes; losses", losses[:3])


if __name__ == "__main__":
    main()
