"""The reference's five MLP agents (algo/{TD3,DDPG,DADDPG,DATD3,DARC}/*_mlp.py) -- same class names, constructor
arguments, take_action / train / save / load surface and update arithmetic -- restructured for the batched engine:

  * train() accepts the reference's dict of numpy arrays OR a dict of device tensors (TrajectoryReplay.sample) and
    never leaves the device unless asked (sync=True mirrors the reference's `loss.cpu().numpy()` return);
  * act(states [N,S]) picks actions for a whole batch of envs on the device (take_action is its N = 1 host wrapper);
  * every network's gradients can live in one flat bucket that is all-reduced once per optimizer step
    (distributed.GradBucket), the single collective of the multi-GPU training path (between
    `loss.backward()` and `optimizer.step()`, TD3_mlp.py:147-148 and :156-157).
"""
import copy

import numpy as np
import torch
import torch.nn.functional as F

from ..config import opt
from ..distributed import GradBucket
from .nets import PolicyNet, QValueNet, TwinQValueNet


class _Learner:
    """one network + its target + Adam (+ optional flat gradient bucket)"""

    def __init__(self, net, lr, device, bucket):
        self.net = net.to(device)
        self.target = copy.deepcopy(self.net)
        for p in self.target.parameters():
            p.requires_grad_(False)
        # capturable: the step counter lives on the device, so a whole update can be recorded in a CUDA graph (train.py)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr, capturable=torch.device(device).type == "cuda")
        self.bucket = GradBucket(self.net.parameters()) if bucket else None

    def step(self, loss):
        if self.bucket is not None:
            self.bucket.zero()
        else:
            self.opt.zero_grad()
        loss.backward()
        if self.bucket is not None:
            self.bucket.allreduce_mean()
        self.opt.step()

    def soft_update(self, tau):
        """target <- target * (1 - tau) + net * tau   (TD3_mlp.py:99-112), all tensors in two fused calls"""
        tp = [p.data for p in self.target.parameters()]
        sp = [p.data for p in self.net.parameters()]
        torch._foreach_mul_(tp, 1.0 - tau)
        torch._foreach_add_(tp, sp, alpha=tau)

    def sync_target(self):
        self.target.load_state_dict(self.net.state_dict())


class _AgentBase:
    _nets = ()          # (attribute name, file suffix) pairs written by save()

    def update_cycle(self):
        """(trains per cycle, phase): the host-side control flow of train() repeats every `cycle` calls and depends only
        on `phase` -- what a CUDA graph of `cycle` consecutive updates is keyed by (VectorTrainer.train_updates)"""
        return 1, 0

    def __init__(self, state_dim, action_dim, action_bound, hidden_dim, sigma, tau, gamma, device, distributed):
        self.state_dim, self.action_dim, self.action_bound = state_dim, action_dim, action_bound
        self.hidden_dim, self.sigma, self.tau, self.gamma = hidden_dim, sigma, tau, gamma
        self.device = torch.device(device)
        self.distributed = bool(distributed)
        self.total_it = 0

    # ---- batches
    def _batch(self, d):
        """(states, actions, rewards [B,1], next_states, dones [B,1]) as float32 tensors on self.device"""
        def t(x, col=False):
            if not torch.is_tensor(x):
                x = torch.as_tensor(np.asarray(x), dtype=torch.float)
            x = x.to(device=self.device, dtype=torch.float)
            return x.view(-1, 1) if col else x
        return t(d['states']), t(d['actions']), t(d['rewards'], True), t(d['next_states']), t(d['dones'], True)

    def _ret(self, loss, sync):
        return loss.detach().cpu().numpy() if sync else loss.detach()

    # ---- acting
    @torch.no_grad()
    def act(self, states):
        raise NotImplementedError

    def acting_policy(self):
        """the single PolicyNet act() evaluates, or None when acting is more than one MLP forward (double-actor agents
        pick between two actors by Q value) -- lets the rollout run the forward as one fused launch (env.policy_act)"""
        return None

    def take_action(self, state):
        """reference signature: one state (array [S]) -> action (array [A]) on the host"""
        s = torch.as_tensor(np.asarray(state), dtype=torch.float, device=self.device).view(1, -1)
        return self.act(s)[0].cpu().numpy()

    # ---- persistence (file names of the reference: TD3_mlp.py:163-168, DARC_mlp.py:221-230)
    def _learners(self):
        return [(getattr(self, "_" + name), suffix) for name, suffix in self._nets]

    def save(self, filename):
        for lr, suffix in self._learners():
            torch.save(lr.net.state_dict(), filename + suffix)

    def load(self, filename):
        """loads what save() wrote and re-synchronises the targets (the reference's load() assigns misspelt attributes,
        TD3_mlp.py:170-177, so its targets silently keep their old weights; here they follow the loaded nets)"""
        for lr, suffix in self._learners():
            lr.net.load_state_dict(torch.load(filename + suffix, map_location=self.device))
            lr.sync_target()

    def state_dict(self):
        """everything needed to resume: nets, targets, optimizers, update counter (checkpoint.py)"""
        out = {"total_it": self.total_it}
        for lr, suffix in self._learners():
            out[suffix] = {"net": lr.net.state_dict(), "target": lr.target.state_dict(), "opt": lr.opt.state_dict()}
        return out

    def load_state_dict(self, sd):
        self.total_it = int(sd["total_it"])
        for lr, suffix in self._learners():
            lr.net.load_state_dict(sd[suffix]["net"])
            lr.target.load_state_dict(sd[suffix]["target"])
            lr.opt.load_state_dict(sd[suffix]["opt"])

    def broadcast_parameters(self, src=0):
        from ..distributed import broadcast_module
        for lr, _ in self._learners():
            broadcast_module(lr.net, src)
            lr.sync_target()


class DDPG_MLP(_AgentBase):
    """algo/DDPG/DDPG_mlp.py:33-160."""
    _nets = (("critic_l", "_critic.pt"), ("actor_l", "_actor.pt"))

    def __init__(self, state_dim, action_dim, action_bound, hidden_dim=opt.hidden_dim, actor_lr=opt.actor_lr,
                 critic_lr=opt.critic_lr, sigma=opt.sigma, tau=opt.tau, gamma=opt.gamma, device=None, distributed=False):
        super().__init__(state_dim, action_dim, action_bound, hidden_dim, sigma, tau, gamma, device or opt.device, distributed)
        self._actor_l = _Learner(PolicyNet(state_dim, hidden_dim, action_dim, action_bound), actor_lr, self.device, distributed)
        self._critic_l = _Learner(QValueNet(state_dim, hidden_dim, action_dim), critic_lr, self.device, distributed)
        self.actor, self.target_actor = self._actor_l.net, self._actor_l.target
        self.critic, self.target_critic = self._critic_l.net, self._critic_l.target
        self.actor_optimizer, self.critic_optimizer = self._actor_l.opt, self._critic_l.opt

    @torch.no_grad()
    def act(self, states):
        return self.actor(states)

    def acting_policy(self):
        return self.actor

    def train(self, transition_dict, sync=True):
        s, a, r, s2, d = self._batch(transition_dict)
        with torch.no_grad():
            y = r + (1 - d) * self.gamma * self.target_critic(s2, self.target_actor(s2))      # DDPG_mlp.py:121-124
        critic_loss = F.mse_loss(self.critic(s, a), y)                                        # :127-130
        self._critic_l.step(critic_loss)
        self._actor_l.step(-self.critic(s, self.actor(s)).mean())                             # :137-140
        self._actor_l.soft_update(self.tau)
        self._critic_l.soft_update(self.tau)
        self.total_it += 1
        return self._ret(critic_loss, sync)


class TD3_MLP(_AgentBase):
    """algo/TD3/TD3_mlp.py:33-177: twin critic in one module, target-policy smoothing, delayed actor."""
    _nets = (("critic_l", "_critic.pt"), ("actor_l", "_actor.pt"))

    def __init__(self, state_dim, action_dim, action_bound, hidden_dim=opt.hidden_dim, actor_lr=opt.actor_lr,
                 critic_lr=opt.critic_lr, sigma=opt.sigma, tau=opt.tau, gamma=opt.gamma, policy_noise=opt.policy_noise,
                 noise_clip=opt.noise_clip, policy_freq=opt.policy_freq, device=None, distributed=False):
        super().__init__(state_dim, action_dim, action_bound, hidden_dim, sigma, tau, gamma, device or opt.device, distributed)
        self.policy_noise, self.noise_clip, self.policy_freq = policy_noise, noise_clip, policy_freq
        self._actor_l = _Learner(PolicyNet(state_dim, hidden_dim, action_dim, action_bound), actor_lr, self.device, distributed)
        self._critic_l = _Learner(TwinQValueNet(state_dim, hidden_dim, action_dim), critic_lr, self.device, distributed)
        self.actor, self.target_actor = self._actor_l.net, self._actor_l.target
        self.critic, self.target_critic = self._critic_l.net, self._critic_l.target
        self.actor_optimizer, self.critic_optimizer = self._actor_l.opt, self._critic_l.opt

    @torch.no_grad()
    def act(self, states):
        return self.actor(states)

    def acting_policy(self):
        return self.actor

    def update_cycle(self):
        return int(self.policy_freq), self.total_it % int(self.policy_freq)      # delayed actor update, TD3_mlp.py:151

    def train(self, transition_dict, sync=True):
        s, a, r, s2, d = self._batch(transition_dict)
        self.total_it += 1
        with torch.no_grad():
            noise = (torch.randn_like(a) * self.policy_noise).clamp(-self.noise_clip, self.noise_clip)   # TD3_mlp.py:127-129
            a2 = (self.target_actor(s2) + noise).clamp(-self.action_bound, self.action_bound)            # :131-133
            q1, q2 = self.target_critic(s2, a2)
            y = r + (1 - d) * self.gamma * torch.min(q1, q2)                                             # :135-137
        c1, c2 = self.critic(s, a)
        critic_loss = F.mse_loss(c1, y) + F.mse_loss(c2, y)                                              # :143
        self._critic_l.step(critic_loss)
        if self.total_it % self.policy_freq == 0:                                                        # :151
            self._actor_l.step(-self.critic.Q1(s, self.actor(s)).mean())
            self._actor_l.soft_update(self.tau)
            self._critic_l.soft_update(self.tau)
        return self._ret(critic_loss, sync)


class _DoubleActorBase(_AgentBase):
    """two actors; act with the one whose critic rates its action higher (DADDPG_mlp.py:84-97, DATD3_mlp.py:95-107)"""

    def _q_pair(self, states, a1, a2):
        raise NotImplementedError

    @torch.no_grad()
    def act(self, states):
        a1, a2 = self.actor1(states), self.actor2(states)
        q1, q2 = self._q_pair(states, a1, a2)
        return torch.where(q1 >= q2, a1, a2)


class DADDPG_MLP(_DoubleActorBase):
    """algo/DADDPG/DADDPG_mlp.py:33-190: two actors, one critic; the actors alternate by the parity of total_it."""
    _nets = (("critic_l", "_critic.pt"), ("actor1_l", "_actor1.pt"), ("actor2_l", "_actor2.pt"))

    def __init__(self, state_dim, action_dim, action_bound, hidden_dim=opt.hidden_dim, actor_lr=opt.actor_lr,
                 critic_lr=opt.critic_lr, sigma=opt.sigma, tau=opt.tau, gamma=opt.gamma, device=None, distributed=False):
        super().__init__(state_dim, action_dim, action_bound, hidden_dim, sigma, tau, gamma, device or opt.device, distributed)
        self._actor1_l = _Learner(PolicyNet(state_dim, hidden_dim, action_dim, action_bound), actor_lr, self.device, distributed)
        self._actor2_l = _Learner(PolicyNet(state_dim, hidden_dim, action_dim, action_bound), actor_lr, self.device, distributed)
        self._critic_l = _Learner(QValueNet(state_dim, hidden_dim, action_dim), critic_lr, self.device, distributed)
        self.actor1, self.target_actor1 = self._actor1_l.net, self._actor1_l.target
        self.actor2, self.target_actor2 = self._actor2_l.net, self._actor2_l.target
        self.critic, self.target_critic = self._critic_l.net, self._critic_l.target

    def _q_pair(self, states, a1, a2):
        return self.critic(states, a1), self.critic(states, a2)

    def update_cycle(self):
        return 2, self.total_it % 2                                              # the actors alternate, DADDPG_mlp.py:119

    def train(self, transition_dict, batch_size=opt.batch_size, sync=True):
        return self.update(transition_dict, batch_size, sync)

    def update(self, transition_dict, batch_size=opt.batch_size, sync=True):
        self.total_it += 1
        first = self.total_it % 2 == 0                                                   # DADDPG_mlp.py:119
        s, a, r, s2, d = self._batch(transition_dict)
        with torch.no_grad():
            q = torch.min(self.target_critic(s2, self.target_actor1(s2)), self.target_critic(s2, self.target_actor2(s2)))
            y = r + (1 - d) * self.gamma * q                                             # :132-140
        critic_loss = F.mse_loss(self.critic(s, a), y)
        self._critic_l.step(critic_loss)
        if first:
            self._actor1_l.step(-self.critic(s, self.actor1(s)).mean())                  # :152-159
            self._actor1_l.soft_update(self.tau)
        else:
            self._actor2_l.step(-self.critic(s, self.actor2(s)).mean())                  # :160-167
            self._actor2_l.soft_update(self.tau)
            self._critic_l.soft_update(self.tau)
        return self._ret(critic_loss, sync)


class DATD3_MLP(_DoubleActorBase):
    """algo/DATD3/DATD3_mlp.py:33-235: two actors + two critics, cross-update (train = update(a1) then update(a2)),
    target = max(min(Q1', Q2'), min(Q1', Q2')) as written in the reference (:155-163)."""
    _nets = (("critic1_l", "_critic1.pt"), ("actor1_l", "_actor1.pt"), ("critic2_l", "_critic2.pt"), ("actor2_l", "_actor2.pt"))
    _q_weight = None
    _reg = 0.0

    def __init__(self, state_dim, action_dim, action_bound, hidden_dim=opt.hidden_dim, actor_lr=opt.actor_lr,
                 critic_lr=opt.critic_lr, sigma=opt.sigma, tau=opt.tau, gamma=opt.gamma, policy_noise=opt.policy_noise,
                 noise_clip=opt.noise_clip, policy_freq=opt.policy_freq, device=None, distributed=False):
        super().__init__(state_dim, action_dim, action_bound, hidden_dim, sigma, tau, gamma, device or opt.device, distributed)
        self.policy_noise, self.noise_clip, self.policy_freq = policy_noise, noise_clip, policy_freq
        mk_a = lambda: _Learner(PolicyNet(state_dim, hidden_dim, action_dim, action_bound), actor_lr, self.device, distributed)
        mk_c = lambda: _Learner(QValueNet(state_dim, hidden_dim, action_dim), critic_lr, self.device, distributed)
        self._actor1_l, self._actor2_l, self._critic1_l, self._critic2_l = mk_a(), mk_a(), mk_c(), mk_c()
        self.actor1, self.target_actor1 = self._actor1_l.net, self._actor1_l.target
        self.actor2, self.target_actor2 = self._actor2_l.net, self._actor2_l.target
        self.critic1, self.target_critic1 = self._critic1_l.net, self._critic1_l.target
        self.critic2, self.target_critic2 = self._critic2_l.net, self._critic2_l.target

    def _q_pair(self, states, a1, a2):
        return self.critic1(states, a1), self.critic2(states, a2)

    def train(self, transition_dict, batch_size=opt.batch_size, sync=True):
        l1 = self.update(transition_dict, True, batch_size, sync)                        # DATD3_mlp.py:126-129
        self.update(transition_dict, False, batch_size, sync)
        return l1

    def _target(self, r, s2, d, a):
        noise = (torch.randn_like(a) * self.policy_noise).clamp(-self.noise_clip, self.noise_clip)
        n1 = (self.target_actor1(s2) + noise).clamp(-self.action_bound, self.action_bound)
        n2 = (self.target_actor2(s2) + noise).clamp(-self.action_bound, self.action_bound)
        # both "a1" and "a2" estimates pair critic1 with actor1's action and critic2 with actor2's (:152-156)
        tq = torch.min(self.target_critic1(s2, n1), self.target_critic2(s2, n2))
        if self._q_weight is None:
            q = torch.max(tq, tq)                                                        # :158-161
        else:
            q = self._q_weight * torch.min(tq, tq) + (1.0 - self._q_weight) * torch.max(tq, tq)   # DARC_mlp.py:172
        return r + (1 - d) * self.gamma * q

    def update(self, transition_dict, update_a1=True, batch_size=100, sync=True):
        s, a, r, s2, d = self._batch(transition_dict)
        self.total_it += 1
        with torch.no_grad():
            y = self._target(r, s2, d, a)
        mine, other = (self._critic1_l, self._critic2_l) if update_a1 else (self._critic2_l, self._critic1_l)
        actor = self._actor1_l if update_a1 else self._actor2_l
        q_mine = mine.net(s, a)
        loss = F.mse_loss(q_mine, y)
        if self._reg:
            loss = loss + self._reg * F.mse_loss(q_mine, other.net(s, a))                # DARC_mlp.py:181 / :203
        mine.step(loss)
        actor.step(-mine.net(s, actor.net(s)).mean())
        actor.soft_update(self.tau)
        mine.soft_update(self.tau)
        return self._ret(loss, sync)


class DARC_MLP(DATD3_MLP):
    """algo/DARC/DARC_mlp.py:33-240: DATD3 + soft target nu*min + (1-nu)*max (q_weight, :172) + critic regulariser
    lambda * MSE(Q1, Q2) (regularization_weight, :181, :203)."""

    def __init__(self, state_dim, action_dim, action_bound, hidden_dim=opt.hidden_dim, actor_lr=opt.actor_lr,
                 critic_lr=opt.critic_lr, sigma=opt.sigma, tau=opt.tau, gamma=opt.gamma, policy_noise=opt.policy_noise,
                 noise_clip=opt.noise_clip, policy_freq=opt.policy_freq, q_weight=opt.q_weight,
                 regularization_weight=opt.regularization_weight, device=None, distributed=False):
        super().__init__(state_dim, action_dim, action_bound, hidden_dim, actor_lr, critic_lr, sigma, tau, gamma, policy_noise,
                         noise_clip, policy_freq, device, distributed)
        self.q_weight, self.regularization_weight = q_weight, regularization_weight
        self._q_weight, self._reg = q_weight, regularization_weight
