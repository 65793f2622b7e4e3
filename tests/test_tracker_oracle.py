"""Pins oracle/tracker_oracle.py (the MAPPO trainer's un-cleared package tracker) to fixtures recorded from the
unmodified reference with the tracker driven exactly like MAPPO/trainer.py:95-130, 229-234 drives it."""
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN, load_kats, load_traj
from oracle import oracle_env as orc
from oracle import tracker_oracle as trk

STALE_NAMES = ["map1_stale", "map7_stale", "map5_stale", "rect_stale"]  # rect_stale: a non-square 10x15 map


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


@pytest.mark.parametrize("name", STALE_NAMES)
def test_stale_trajectory(name):
    meta, z = load_traj(name)
    assert meta["stale"] and meta["n_shadowed"] > 0
    R, P, T = meta["R"], meta["P"], meta["T"]
    K_o, K_p, MR, MP = meta["K_o"], meta["K_p"], meta["MR"], meta["MP"]
    e = orc.OracleEnv(z["grid"], T, R, P, seed=meta["seed"])
    e.reset()
    tk = trk.Tracker()
    tk.update(e)
    feat_every, big_every = meta["feat_every"], meta["big_vec_every"]
    fi = bi = 0
    for s in range(meta["n_steps"]):
        a = z["actions"][s]
        before, pp_prev = e.snapshot(), tk.snapshot()
        r, d = e.step(a)
        after = e.snapshot()
        assert r == z["reward"][s] and bool(d) == bool(z["done"][s]), s
        sm = trk.shaped_reward(e, r, before, after, a, pp_prev, orc.MAPPO_SHAPING)
        sq = trk.shaped_reward(e, r, before, after, a, pp_prev, orc.QMIX_SHAPING)
        assert sm == z["shaped_mappo"][s] and sq == z["shaped_qmix"][s], (s, sm, z["shaped_mappo"][s])
        if d:
            e.reset()
        tk.update(e)
        ts, snap = tk.snapshot(), e.snapshot()
        assert [x[0] for x in ts] == z["trk_ids"][s][:len(ts)].tolist() and not z["trk_ids"][s][len(ts):].any(), s
        assert [x[5] for x in ts] == z["trk_status"][s][:len(ts)].tolist(), s
        if s % feat_every == 0:
            o = np.stack([trk.observation(e, snap, ts, i) for i in range(R)])
            v = np.stack([trk.vector_features(e, snap, ts, i, K_o, K_p) for i in range(R)])
            g, gv = trk.global_state(e, snap, ts, MR, MP)
            assert np.array_equal(o, z["obs"][fi].astype(np.float32)), s
            assert np.array_equal(_bits(v), _bits(z["vec"][fi])), s
            assert np.array_equal(g, z["gstate"][fi].astype(np.float32)), s
            assert np.array_equal(_bits(gv), _bits(z["gvec"][fi])), s
            fi += 1
        if big_every and s % big_every == 0:
            vb = np.stack([trk.vector_features(e, snap, ts, i) for i in range(R)])
            assert np.array_equal(_bits(vb), _bits(z["vec_big"][bi])), s
            gq = trk.global_state(e, snap, ts, 10, 20)[1]
            assert np.array_equal(_bits(gq), _bits(z["gvec_qmix"][bi])), s
            bi += 1


def test_kat_b_stale_tracker_sum_1243_89():
    """KAT-B with the tracker the MAPPO trainer really keeps: sum of shaped rewards 1243.89 (cleared: 1627.37)."""
    k = np.load(os.path.join(GOLDEN, "kats_stale.npz"))
    grid = load_traj("map1_c3")[1]["grid"]
    ve = orc.OracleVecEnv(5, grid, seed=42, max_time_steps=500, n_robots=5, n_packages=50)
    ve.reset()
    tks = [trk.Tracker() for _ in range(5)]
    for t_, e in zip(tks, ve.envs):
        t_.update(e)
    rs = np.random.RandomState(0)
    total = 0.0
    for s in range(1200):
        a = rs.randint(0, 15, size=(5, 5))
        shaped, dones, obs, vec, gst, gvec = trk.mappo_rollout_step(ve, tks, a, 4, 5)
        assert np.array_equal(shaped, k["kat_b_stale_shaped"][s]), s
        assert np.array_equal(dones, k["kat_b_stale_done"][s].astype(bool)), s
        fs = vec[:, 0].astype(np.float64).sum(axis=1) + gvec.astype(np.float64).sum(axis=1)
        assert np.array_equal(fs, k["kat_b_stale_featsum"][s]), s
        total += float(shaped.astype(np.float64).sum())
    assert abs(total - 1243.89) < 5e-3
    assert not np.array_equal(k["kat_b_stale_shaped"], load_kats()["kat_b_shaped"])


def test_cleared_tracker_equals_env_truth():
    """With clear-on-reset (QMIX / IDQ / legacy trainers) the tracker path gives exactly the truth-based oracle."""
    grid = load_traj("map7_stale")[1]["grid"]
    a_env = orc.OracleVecEnv(3, grid, seed=9, max_time_steps=20, n_robots=4, n_packages=12)
    b_env = orc.OracleVecEnv(3, grid, seed=9, max_time_steps=20, n_robots=4, n_packages=12)
    a_env.reset()
    b_env.reset()
    tks = [trk.Tracker() for _ in range(3)]
    for t_, e in zip(tks, b_env.envs):
        t_.update(e)
    rs = np.random.RandomState(1)
    for s in range(90):
        a = rs.randint(0, 15, size=(3, 4))
        want = orc.mappo_rollout_step(a_env, a, 3, 5, 10, 20)
        got = trk.mappo_rollout_step(b_env, tks, a, 3, 5, 10, 20, clear_on_reset=True)
        for x, y in zip(want, got):
            assert np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8)), s
