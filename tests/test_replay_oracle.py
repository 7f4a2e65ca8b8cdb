"""CPU: the numpy replay restatement reproduces the REFERENCE's own batches (fixtures made by tools/gen_her_golden.py
from /root/reference/utils/rl_utils.py) on the recorded picks."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("kind", ["reach", "push"])
def test_relabel_matches_reference_batches(kind):
    from oracle import replay_oracle as RO
    g = np.load(os.path.join(GOLD, "her_%s.npz" % kind))
    out = RO.relabel(g["states"], g["actions"], g["rewards"], g["dones"], g["lengths"], g["picks"], kind,
                     float(g["dis_threshold"]))
    assert np.array_equal(out["states"], g["out_states"])
    assert np.array_equal(out["next_states"], g["out_next_states"])
    assert np.array_equal(out["actions"], g["out_actions"])
    assert np.array_equal(out["rewards"], g["out_rewards"])
    assert np.array_equal(out["dones"].astype(np.uint8), g["out_dones"])
    her = g["picks"][:, 2] >= 0
    assert 0.6 < her.mean() < 0.95                                   # her_ratio 0.8 of a 96-sample batch
    assert set(np.unique(g["out_rewards"][her])) == {-0.1, 1.0}      # both relabel outcomes are exercised
