"""Pins the C oracle (oracle/oracle.c) to the reference's golden trajectories and to the Python oracle."""
import numpy as np
import pytest

from golden_util import TRAJ_NAMES, load_traj, reward_kw
from oracle import oracle_env as orc
from oracle_c import COracleEnv, COracleVec

MAPPO = np.array(orc.MAPPO_SHAPING, np.float32)
QMIX = np.array(orc.QMIX_SHAPING, np.float32)


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


@pytest.mark.parametrize("name", TRAJ_NAMES)
def test_c_oracle_golden_trajectory(name):
    meta, z = load_traj(name)
    R, P, T, W = meta["R"], meta["P"], meta["T"], meta["W"]
    K = (meta["K_o"], meta["K_p"], meta["MR"], meta["MP"])
    e = COracleEnv(z["grid"], T, R, P, meta["seed"], **reward_kw(meta))
    e.reset()
    assert np.array_equal(e.packages(), z["reset_pkgs"][0])
    o, v, g, gv = e.features(*K)
    assert np.array_equal(o, z["obs0"].astype(np.float32)) and np.array_equal(_bits(v), _bits(z["vec0"]))
    assert np.array_equal(g, z["gstate0"].astype(np.float32)) and np.array_equal(_bits(gv), _bits(z["gvec0"]))
    fi = bi = 0
    n_reset = 1
    for s in range(meta["n_steps"]):
        a = z["actions"][s]
        r, d = e.step(a)
        rb, st, t, tot = e.state()
        assert np.array_equal(rb, z["robots"][s]) and np.array_equal(st, z["status"][s]), s
        assert r == z["reward"][s] and d == bool(z["done"][s]) and t == z["t_after"][s] and tot == z["total_reward"][s], s
        assert e.shaped(a, MAPPO, r) == z["shaped_mappo"][s] and e.shaped(a, QMIX, r) == z["shaped_qmix"][s], s
        if d:
            e.reset()
            assert np.array_equal(e.packages(), z["reset_pkgs"][n_reset])
            n_reset += 1
        if s % meta["feat_every"] == 0:
            o, v, g, gv = e.features(*K)
            assert np.array_equal(o, z["obs"][fi].astype(np.float32)), s
            assert np.array_equal(_bits(v), _bits(z["vec"][fi])), s
            assert np.array_equal(g, z["gstate"][fi].astype(np.float32)), s
            assert np.array_equal(_bits(gv), _bits(z["gvec"][fi])), s
            fi += 1
        if meta["big_vec_every"] and s % meta["big_vec_every"] == 0:
            _, vb, _, gq = e.features(100, 100, 10, 20)
            assert np.array_equal(_bits(vb), _bits(z["vec_big"][bi])) and np.array_equal(_bits(gq), _bits(z["gvec_qmix"][bi])), s
            bi += 1


@pytest.mark.parametrize("trial", range(6))
def test_c_oracle_vs_python_oracle_random(trial):
    import marl_delivery_b200 as mdb

    rs = np.random.RandomState(50 + trial)
    grid = mdb.maps.stock_map(["map.txt", "map1.txt", "map5.txt"][trial % 3])
    nfree = int((grid == 0).sum())
    R = int(rs.randint(1, min(32, nfree) + 1))
    P, T = int(rs.randint(1, 80)), int(rs.randint(6, 40))
    K = (int(rs.randint(0, R + 2)), int(rs.randint(0, 8)), int(rs.randint(1, 40)), int(rs.randint(0, 40)))
    consts = orc.QMIX_SHAPING if trial % 2 else orc.MAPPO_SHAPING
    seed = int(rs.randint(0, 2 ** 30))
    N = 3
    pv = orc.OracleVecEnv(N, grid, seed=seed, max_time_steps=T, n_robots=R, n_packages=P)
    cv = COracleVec(N, grid, T, R, P, seed, *K, consts)
    pv.reset()
    cv.reset()
    for s in range(2 * T + 3):
        a = rs.randint(0, 18, size=(N, R)).astype(np.uint8)
        want = orc.mappo_rollout_step(pv, a, *K, consts=consts)
        r, sh, d = cv.rollout_step(a)
        assert np.array_equal(_bits(sh), _bits(want[0])) and np.array_equal(d.astype(bool), want[1]), s
        assert np.array_equal(cv.obs, want[2]) and np.array_equal(_bits(cv.vec), _bits(want[3])), s
        assert np.array_equal(cv.gstate, want[4]) and np.array_equal(_bits(cv.gvec), _bits(want[5])), s
