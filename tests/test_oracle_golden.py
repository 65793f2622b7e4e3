"""Pins the Python oracle (oracle/oracle_env.py) to the reference's golden vectors."""
import hashlib

import numpy as np
import pytest

from golden_util import TRAJ_NAMES, load_kats, load_reset_blocks, load_traj, readme_random_actions, reward_kw
from oracle import oracle_env as orc


def test_reset_blocks():
    meta, z = load_reset_blocks()
    for k, m in enumerate(meta):
        e = orc.OracleEnv(z[f"grid_{k}"], m["T"], m["R"], m["P"], seed=m["seed"])
        for b in range(3):
            if b:
                e.reset()
            assert e.rcell == z[f"robots_{k}"][b].tolist(), (m, b)
            pk = z[f"pkgs_{k}"][b]
            assert e.p_start == pk[:, 0].tolist() and e.p_target == pk[:, 1].tolist(), (m, b)
            assert e.p_t0 == pk[:, 2].tolist() and e.p_deadline == pk[:, 3].tolist(), (m, b)


@pytest.mark.parametrize("name", TRAJ_NAMES)
def test_trajectory(name):
    meta, z = load_traj(name)
    R, P, T = meta["R"], meta["P"], meta["T"]
    K_o, K_p, MR, MP = meta["K_o"], meta["K_p"], meta["MR"], meta["MP"]
    e = orc.OracleEnv(z["grid"], T, R, P, seed=meta["seed"], **reward_kw(meta))
    e.reset()
    W = e.W
    s0 = e.snapshot()
    assert np.array_equal(np.stack([orc.observation(e, s0, i) for i in range(R)]), z["obs0"].astype(np.float32))
    assert np.array_equal(np.stack([orc.vector_features(e, s0, i, K_o, K_p) for i in range(R)]), z["vec0"])
    g, gv = orc.global_state(e, s0, MR, MP)
    assert np.array_equal(g, z["gstate0"].astype(np.float32)) and np.array_equal(gv, z["gvec0"])
    feat_every, big_every = meta["feat_every"], meta["big_vec_every"]
    fi = bi = 0
    n_reset = 1
    for s in range(meta["n_steps"]):
        a = z["actions"][s]
        before = e.snapshot()
        r, d = e.step(a)
        after = e.snapshot()
        assert [(c // W, c % W, k) for c, k in zip(e.rcell, e.carry)] == [tuple(x) for x in z["robots"][s].tolist()], s
        assert r == z["reward"][s] and bool(d) == bool(z["done"][s]) and e.t == z["t_after"][s], s
        assert e.p_status == z["status"][s].tolist(), s
        assert e.total_reward == z["total_reward"][s], s
        sm = orc.shaped_reward(e, r, before, after, a, orc.MAPPO_SHAPING)
        sq = orc.shaped_reward(e, r, before, after, a, orc.QMIX_SHAPING)
        assert type(sm) is np.float32
        assert sm == z["shaped_mappo"][s] and sq == z["shaped_qmix"][s], (s, sm, z["shaped_mappo"][s])
        if d:
            e.reset()
            assert e.rcell == z["reset_robots"][n_reset].tolist()
            assert e.p_start == z["reset_pkgs"][n_reset][:, 0].tolist()
            n_reset += 1
        snap = e.snapshot()
        if s % feat_every == 0:
            o = np.stack([orc.observation(e, snap, i) for i in range(R)])
            v = np.stack([orc.vector_features(e, snap, i, K_o, K_p) for i in range(R)])
            g, gv = orc.global_state(e, snap, MR, MP)
            assert np.array_equal(o, z["obs"][fi].astype(np.float32)), s
            assert np.array_equal(v.view(np.uint32), z["vec"][fi].view(np.uint32)), s
            assert np.array_equal(g, z["gstate"][fi].astype(np.float32)), s
            assert np.array_equal(gv.view(np.uint32), z["gvec"][fi].view(np.uint32)), s
            fi += 1
        if big_every and s % big_every == 0:
            vb = np.stack([orc.vector_features(e, snap, i) for i in range(R)])
            assert vb.shape[1] == 1007
            assert np.array_equal(vb.view(np.uint32), z["vec_big"][bi].view(np.uint32)), s
            gq = orc.global_state(e, snap, 10, 20)[1]
            assert np.array_equal(gq.view(np.uint32), z["gvec_qmix"][bi].view(np.uint32)), s
            bi += 1
    assert n_reset == meta["n_resets"]


def test_kat_a():
    k = load_kats()
    e = orc.OracleEnv(orc.parse_map_text(MAP1), 500, 5, 50, seed=2025)
    e.reset()
    assert e.robots_rc() == [(2, 5, 0), (1, 4, 0), (3, 3, 0), (7, 8, 0), (1, 6, 0)]  # SURVEY.md KAT-A (0-indexed)
    rs = np.random.RandomState(7)
    for s in range(500):
        a = rs.randint(0, 15, size=5)
        assert np.array_equal(a, k["kat_a_actions"][s])
        r, d = e.step(a)
        assert r == k["kat_a_rewards"][s]
    assert d and e.total_reward == -12.839999999999886 == float(k["kat_a_total"])
    assert sum(s == orc.DELIVERED for s in e.p_status) == 4


def test_kat_b_vectorised_autoreset():
    k = load_kats()
    ve = orc.OracleVecEnv(5, orc.parse_map_text(MAP1), seed=42, max_time_steps=500, n_robots=5, n_packages=50)
    ve.reset()
    rs = np.random.RandomState(0)
    total = 0.0
    for s in range(1200):
        a = rs.randint(0, 15, size=(5, 5))
        shaped, dones, obs, vec, gst, gvec = orc.mappo_rollout_step(ve, a, 4, 5)
        assert np.array_equal(shaped, k["kat_b_shaped"][s]), s
        assert np.array_equal(dones, k["kat_b_done"][s].astype(bool)), s
        fs = vec[:, 0].astype(np.float64).sum(axis=1) + gvec.astype(np.float64).sum(axis=1)
        assert np.array_equal(fs, k["kat_b_featsum"][s]), s
        total += float(shaped.astype(np.float64).sum())
    assert abs(total - 1627.37) < 5e-3  # SURVEY.md KAT-B (cleared tracker)


def test_readme_random_row():
    """README.md:117-121 Random row reproduced by the oracle: -16.50 +- 6.24 / 13.66 +- 3.63."""
    k = load_kats()
    acts = readme_random_actions()
    assert np.array_equal(acts[0], k["readme_actions_ep0"])
    assert hashlib.sha256(acts.tobytes()).hexdigest() == str(k["readme_actions_sha"])
    grid = orc.parse_map_text(MAP1)
    rewards, delivered = [], []
    for ep in range(100):
        e = orc.OracleEnv(grid, 1000, 5, 100, seed=10 + ep)
        e.reset()
        for s in range(1000):
            r, d = e.step(acts[ep, s])
        assert d
        rewards.append(e.total_reward)
        delivered.append(sum(x == orc.DELIVERED for x in e.p_status))
    assert np.array_equal(np.array(rewards), k["readme_rewards"])
    assert np.array_equal(np.array(delivered), k["readme_delivered"])
    assert f"{np.mean(rewards):.2f} {np.std(rewards):.2f}" == "-16.50 6.24"
    assert f"{np.mean(delivered):.2f} {np.std(delivered):.2f}" == "13.66 3.63"


def test_notebook_dummy_shaping():
    """MAPPO notebook cell 19 -> 215.04000854492188; QMIX notebook cell 20 -> 12.539999961853027.
    Inputs re-typed from the notebook cell (0-indexed here)."""
    W = 10
    prev = [(1, 1, 0), (4, 4, 1)]
    cur = [(2, 1, 2), (4, 3, 0)]
    acts = [0 + 5 * 1, 1 + 5 * 2]  # ('D','1'), ('L','2')
    pk = {1: (0 * W + 0, 4 * W + 3, 0, 10, orc.IN_TRANSIT), 2: (2 * W + 1, 0 * W + 4, 3, 15, orc.WAITING)}
    m = orc.shaped_reward_raw(W, 10, prev, cur, 4, 5, acts, pk, orc.MAPPO_SHAPING)
    q = orc.shaped_reward_raw(W, 10, prev, cur, 4, 5, acts, pk, orc.QMIX_SHAPING)
    assert float(m) == 215.04000854492188
    assert float(q) == 12.539999961853027


def test_feature_dims():
    """Notebook cell 25: obs (6,10,10), state (4,10,10), vec 1007, global vec 1301; trainer call site 52."""
    e = orc.OracleEnv(orc.parse_map_text(MAP1), 500, 5, 50, seed=1)
    s = e.snapshot()
    assert orc.observation(e, s, 0).shape == (6, 10, 10)
    assert orc.vector_features(e, s, 0).shape == (1007,)
    assert orc.vector_features(e, s, 0, 4, 5).shape == (52,)
    g, gv = orc.global_state(e, s)
    assert g.shape == (4, 10, 10) and gv.shape == (1301,)


MAP1 = "\n".join(
    " ".join("1" if (r in (0, 9) or c in (0, 9)) else "0" for c in range(10)) for r in range(10)
) + "\n"
