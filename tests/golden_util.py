"""Shared loaders for the fixtures in tests/golden (see tests/golden/make_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TRAJ_NAMES = ["map1_c3", "map1_mix", "map7_odd", "map7_dense", "map7_done", "map5_c4", "map5_c5",
              "map1_home", "map7_home", "map3_small",
              # synthetic maps (make_golden.py shapes): non-square, planes that are not whole float4s, planes beyond 512 cells
              "rect_12x17", "rect_9x7", "rect_5x24", "big_40x28", "big_23x23",
              # move_cost / delivery_reward / delay_reward other than the constructor defaults
              "map1_rewards", "map7_rewards"]


def reward_kw(meta):
    """Constructor keywords for the fixture's reward constants ({} = the reference's defaults)."""
    rw = meta.get("rewards")
    return dict(move_cost=rw[0], delivery_reward=rw[1], delay_reward=rw[2]) if rw else {}


def load_traj(name):
    z = np.load(os.path.join(GOLDEN, f"traj_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def load_reset_blocks():
    z = np.load(os.path.join(GOLDEN, "reset_blocks.npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def load_kats():
    return np.load(os.path.join(GOLDEN, "kats.npz"))


def readme_random_actions():
    """Re-draw the README Random-row actions the way evaluation.py + randomagent.py do: global numpy
    RNG seeded with 10, per robot `np.random.choice(['U','D','L','R','S'])` then
    `np.random.choice(['0','1','2'])` (randomagent.py:16-22).  Returned in the action-code encoding
    of SURVEY.md section 3A (move 'DLRSU'[a % 5], op a // 5)."""
    rs = np.random.RandomState(10)  # np.random.seed(10) seeds the same legacy generator
    mv_map = np.array([4, 0, 1, 2, 3], np.uint8)  # U D L R S -> codes in 'DLRSU'
    draws = rs.randint(0, 1, size=0)  # no-op, keeps the generator type explicit
    del draws
    out = np.zeros((100, 1000, 5), np.uint8)
    for ep in range(100):
        for s in range(1000):
            for i in range(5):
                m = rs.choice(5)
                o = rs.choice(3)
                out[ep, s, i] = mv_map[m] + 5 * o
    return out
