"""Pins the greedy-policy oracle to traces recorded from the reference's GreedyAgents."""
import json
import os

import numpy as np

from golden_util import GOLDEN
from oracle import oracle_env as orc
from oracle.greedy_oracle import GreedyOracle

MAP1 = "\n".join(" ".join("1" if (r in (0, 9) or c in (0, 9)) else "0" for c in range(10)) for r in range(10))


def _run(grid, T, R, P, seed):
    e = orc.OracleEnv(grid, T, R, P, seed=seed)
    e.reset()
    ag = GreedyOracle(e)
    acts, rews = [], []
    done = False
    while not done:
        a = ag.get_actions()
        acts.append(a)
        r, done = e.step(a)
        rews.append(r)
    return e, np.array(acts, np.uint8), np.array(rews)


def test_greedy_traces_and_readme_row():
    z = np.load(os.path.join(GOLDEN, "kats_greedy.npz"))
    for k, m in enumerate(json.loads(str(z["greedy_meta"]))):
        e, acts, rews = _run(z[f"greedy_grid_{k}"], m["T"], m["R"], m["P"], m["seed"])
        assert np.array_equal(acts, z[f"greedy_actions_{k}"]), m
        assert np.array_equal(rews, z[f"greedy_rewards_{k}"]) and e.total_reward == m["total"]
    grid = orc.parse_map_text(MAP1)
    rewards, delivered = [], []
    for ep in range(100):
        e, acts, _ = _run(grid, 1000, 5, 100, 10 + ep)
        if ep < 3:
            assert np.array_equal(acts, z[f"readme_greedy_actions_ep{ep}"]), ep
        rewards.append(e.total_reward)
        delivered.append(sum(s == orc.DELIVERED for s in e.p_status))
    assert np.array_equal(np.array(rewards), z["readme_greedy_rewards"])
    assert np.array_equal(np.array(delivered), z["readme_greedy_delivered"])
    assert f"{np.mean(rewards):.2f} {np.std(rewards):.2f}" == "34.04 14.83"  # README.md:119
    assert f"{np.mean(delivered):.2f} {np.std(delivered):.2f}" == "3.79 1.55"
