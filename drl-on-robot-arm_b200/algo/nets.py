"""Actor / critic MLPs of the reference agents (algo/*/net_mlp.py): [S] -> 256 -> 256 -> A with tanh * action_bound,
and [S+A] -> 256 -> 256 -> 1 critics.  Layer names follow the reference (fc1..fc3, twin critic fc1..fc6) so that
state_dicts saved by either side load into the other (TD3_mlp.py:163-168)."""
import torch
from torch import nn


def _mlp3(i, h, o):
    return nn.Linear(i, h), nn.Linear(h, h), nn.Linear(h, o)


class PolicyNet(nn.Module):
    """algo/TD3/net_mlp.py:29-40 (identical in DDPG / DADDPG / DATD3 / DARC)."""

    def __init__(self, state_dim, hidden_dim, action_dim, action_bound):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = _mlp3(state_dim, hidden_dim, action_dim)
        self.action_bound = action_bound

    def forward(self, x):
        h = torch.relu(self.fc2(torch.relu(self.fc1(x))))
        return torch.tanh(self.fc3(h)) * self.action_bound


class QValueNet(nn.Module):
    """single Q(s, a): algo/DDPG/net_mlp.py:43-56 (also DADDPG, DATD3, DARC)."""

    def __init__(self, state_dim, hidden_dim, action_dim):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = _mlp3(state_dim + action_dim, hidden_dim, 1)

    def forward(self, state, action):
        x = torch.cat([state, action], dim=1)
        return self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))


class TwinQValueNet(nn.Module):
    """two Q heads in one module: algo/TD3/net_mlp.py:43-71 (fc1-3 = Q1, fc4-6 = Q2)."""

    def __init__(self, state_dim, hidden_dim, action_dim):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = _mlp3(state_dim + action_dim, hidden_dim, 1)
        self.fc4, self.fc5, self.fc6 = _mlp3(state_dim + action_dim, hidden_dim, 1)

    def forward(self, state, action):
        x = torch.cat([state, action], dim=1)
        q1 = self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))
        q2 = self.fc6(torch.relu(self.fc5(torch.relu(self.fc4(x)))))
        return q1, q2

    def Q1(self, state, action):
        x = torch.cat([state, action], dim=1)
        return self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))
