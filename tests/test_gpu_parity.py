"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libmdx.so via ctypes): the CUDA path is compared with the reference's golden fixtures and, on fresh
seeded inputs, with the Python oracle.  Integers, flags and every feature tensor must be BIT-EXACT;
the env reward is compared as float64 equality (tolerance of the north star: 1e-6) and the shaped
reward as float32 equality."""
import numpy as np
import pytest
import torch

import marl_delivery_b200 as mdb
from golden_util import TRAJ_NAMES, load_kats, load_traj, readme_random_actions, reward_kw
from oracle import oracle_env as orc

pytestmark = pytest.mark.gpu

REWARD_TOL = 1e-6  # north-star tolerance; the tests below actually demand equality


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


def _batch_for(meta, z, n_envs=1, shaping=mdb.MAPPO_SHAPING, K=None, **kw):
    K_o, K_p, MR, MP = K if K else (meta["K_o"], meta["K_p"], meta["MR"], meta["MP"])
    return mdb.DeliveryBatch([z["grid"]], n_envs, meta["R"], meta["P"], meta["T"], seeds=[meta["seed"]] * n_envs,
                             k_others=K_o, k_packages=K_p, max_robots_in_state=MR, max_packages_in_state=MP,
                             shaping=shaping, **reward_kw(meta), **kw)


def _state_tuple(b, i=0):
    s = b.get_state(i)
    robots = np.stack([s["robot_row"], s["robot_col"], s["robot_carry"]], 1)
    return s, robots


@pytest.mark.parametrize("name", TRAJ_NAMES)
def test_golden_trajectory_unfused(name):
    """step -> compare state with the reference -> explicit reset on done -> featurize."""
    meta, z = load_traj(name)
    b = _batch_for(meta, z)
    bq = _batch_for(meta, z, shaping=mdb.QMIX_SHAPING)
    W = meta["W"]
    for bb in (b, bq):
        bb.reset()
    s, robots = _state_tuple(b)
    assert np.array_equal(robots[:, 0] * W + robots[:, 1], z["reset_robots"][0])
    pk = np.stack([s["pkg_start_row"] * W + s["pkg_start_col"], s["pkg_target_row"] * W + s["pkg_target_col"],
                   s["pkg_start_time"], s["pkg_deadline"]], 1)
    assert np.array_equal(pk, z["reset_pkgs"][0])
    o, v, g, gv = [t.cpu().numpy() for t in b.featurize()]
    assert np.array_equal(o[0], z["obs0"].astype(np.float32))
    assert np.array_equal(_bits(v[0]), _bits(z["vec0"]))
    assert np.array_equal(g[0], z["gstate0"].astype(np.float32))
    assert np.array_equal(_bits(gv[0]), _bits(z["gvec0"]))
    fi = 0
    n_reset = 1
    for step in range(meta["n_steps"]):
        a = torch.from_numpy(z["actions"][step][None].copy())
        r, sh, d = b.step(a)
        _, shq, _ = bq.step(a)
        r, sh, d, shq = float(r.cpu()[0]), sh.cpu().numpy()[0], int(d.cpu()[0]), shq.cpu().numpy()[0]
        s, robots = _state_tuple(b)
        assert np.array_equal(robots, z["robots"][step]), step
        assert abs(r - z["reward"][step]) <= REWARD_TOL and r == z["reward"][step], step
        assert d == z["done"][step] and s["t"] == z["t_after"][step], step
        assert np.array_equal(s["pkg_status"], z["status"][step]), step
        assert s["total_reward"] == z["total_reward"][step], step
        assert sh == z["shaped_mappo"][step], (step, sh, z["shaped_mappo"][step])
        assert shq == z["shaped_qmix"][step], (step, shq, z["shaped_qmix"][step])
        if d:
            b.reset()
            bq.reset()
            s, robots = _state_tuple(b)
            assert np.array_equal(robots[:, 0] * W + robots[:, 1], z["reset_robots"][n_reset]), step
            assert s["t"] == 0 and s["total_reward"] == 0 and s["episode"] == n_reset + 1
            n_reset += 1
        if step % meta["feat_every"] == 0:
            o, v, g, gv = [t.cpu().numpy() for t in b.featurize()]
            assert np.array_equal(o[0], z["obs"][fi].astype(np.float32)), step
            assert np.array_equal(_bits(v[0]), _bits(z["vec"][fi])), step
            assert np.array_equal(g[0], z["gstate"][fi].astype(np.float32)), step
            assert np.array_equal(_bits(gv[0]), _bits(z["gvec"][fi])), step
            fi += 1
    assert n_reset == meta["n_resets"]


@pytest.mark.parametrize("name", TRAJ_NAMES)
def test_golden_trajectory_fused_autoreset(name):
    """The one-launch-per-iteration path (step + shaped + on-device auto-reset + featurise), the way
    MAPPO/trainer.py:172-286 sequences it.  Four copies of the env run side by side in one batch."""
    meta, z = load_traj(name)
    N = 4
    b = _batch_for(meta, z, n_envs=N)
    b.reset()
    fi = 0
    for step in range(meta["n_steps"]):
        a = torch.from_numpy(np.repeat(z["actions"][step][None], N, 0).copy())
        r, sh, d, o, v, g, gv = b.step_featurize(a, auto_reset=True)
        r, sh, d = r.cpu().numpy(), sh.cpu().numpy(), d.cpu().numpy()
        assert np.all(r == z["reward"][step]) and np.all(d == z["done"][step]), step
        assert np.all(sh == z["shaped_mappo"][step]), step
        if step % meta["feat_every"] == 0:
            o, v, g, gv = o.cpu().numpy(), v.cpu().numpy(), g.cpu().numpy(), gv.cpu().numpy()
            for i in range(N):
                assert np.array_equal(o[i], z["obs"][fi].astype(np.float32)), (step, i)
                assert np.array_equal(_bits(v[i]), _bits(z["vec"][fi])), (step, i)
                assert np.array_equal(g[i], z["gstate"][fi].astype(np.float32)), (step, i)
                assert np.array_equal(_bits(gv[i]), _bits(z["gvec"][fi])), (step, i)
            fi += 1
    st = b.episode_stats().cpu().numpy()
    assert st[0] == N * (meta["n_resets"] - 1) and st[5] == N * meta["n_steps"] and st[7] == 0
    carry = z["robots"][:, :, 2]
    prev = np.concatenate([np.zeros((1, meta["R"]), carry.dtype), carry[:-1]], 0)
    prev[1:][z["done"][:-1].astype(bool)] = 0  # a reset clears what the robots carried
    n_deliveries = int(((prev != 0) & (carry == 0)).sum())  # the env only drops on delivery (env.py:282-292)
    n_pickups = int(((prev == 0) & (carry != 0)).sum())
    assert st[2] + st[3] == N * n_deliveries and st[4] == N * n_pickups


@pytest.mark.parametrize("name", ["map1_c3", "map1_mix", "map7_odd", "map7_done", "map1_home"])
def test_golden_default_k_vec_and_qmix_gvec(name):
    """generate_vector_features with its defaults (K=100/100 -> dim 1007, agent.py / main.py / the
    notebook run) and convert_global_state with the QMIX sizes 10/20 (QMIX/config.yaml:57-58)."""
    meta, z = load_traj(name)
    b = _batch_for(meta, z, K=(100, 100, 10, 20))
    assert b.vec_dim == 1007 and b.gvec_dim == 201
    b.reset()
    bi = 0
    for step in range(meta["n_steps"]):
        a = torch.from_numpy(z["actions"][step][None].copy())
        _, _, d = b.step(a, want_shaped=False)
        if int(d.cpu()[0]):
            b.reset()
        if step % meta["big_vec_every"] == 0:
            _, v, _, gv = b.featurize()
            assert np.array_equal(_bits(v.cpu().numpy()[0]), _bits(z["vec_big"][bi])), step
            assert np.array_equal(_bits(gv.cpu().numpy()[0]), _bits(z["gvec_qmix"][bi])), step
            bi += 1


def test_kat_a_through_environment_api():
    """SURVEY.md KAT-A through the reference-shaped API: strings in, dicts out."""
    k = load_kats()
    e = mdb.Environment("map1.txt", 500, 5, 50, seed=2025)
    st = e.reset()
    assert st["robots"] == [(3, 6, 0), (2, 5, 0), (4, 4, 0), (8, 9, 0), (2, 7, 0)] and st["time_step"] == 0
    assert st["map"] == mdb.maps.stock_map("map1.txt").tolist()
    moves, ops = "DLRSU", "012"
    for s in range(500):
        acts = [(moves[a % 5], ops[a // 5]) for a in k["kat_a_actions"][s]]
        st, r, done, info = e.step(acts)
        assert r == k["kat_a_rewards"][s]
    assert done and info["total_reward"] == -12.839999999999886 and info["total_time_steps"] == 500
    assert sum(p.status == "delivered" for p in e.packages) == 4
    with pytest.raises(ValueError):
        e.step([("S", "0")] * 4)


def test_kat_b_vectorised_autoreset():
    """SURVEY.md KAT-B: 5 envs, seed 42, 1200 steps, sum of shaped reward 1627.37 (cleared tracker)."""
    k = load_kats()
    ve = mdb.VectorizedEnv(mdb.Environment, 5, map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=500,
                           seed=42, k_others=4, k_packages=5)
    ve.reset()
    rs = np.random.RandomState(0)
    total = 0.0
    for s in range(1200):
        a = rs.randint(0, 15, size=(5, 5)).astype(np.uint8)
        r, sh, d, o, v, g, gv = ve.rollout_step(torch.from_numpy(a))
        assert np.array_equal(sh.cpu().numpy(), k["kat_b_shaped"][s]), s
        assert np.array_equal(r.cpu().numpy(), k["kat_b_reward"][s]), s
        assert np.array_equal(d.cpu().numpy(), k["kat_b_done"][s]), s
        fs = v[:, 0].double().sum(1) + gv.double().sum(1)
        assert np.array_equal(fs.cpu().numpy(), k["kat_b_featsum"][s]), s
        total += float(sh.double().sum())
    assert abs(total - 1627.37) < 5e-3


def test_readme_random_row():
    """README.md:117-121 Random row (-16.50 +- 6.24 reward, 13.66 +- 3.63 delivered) on the GPU: the 100
    evaluation episodes are 100 envs of one batch (seed 10 + episode, evaluation.py:12-19)."""
    k = load_kats()
    acts = readme_random_actions()  # [100, 1000, 5]
    ve = mdb.VectorizedEnv(mdb.Environment, 100, map_file="map1.txt", n_robots=5, n_packages=100, max_time_steps=1000,
                           seed=10)
    ve.reset()
    a_dev = torch.from_numpy(np.ascontiguousarray(acts.transpose(1, 0, 2))).cuda()
    total = torch.zeros(100, dtype=torch.float64, device="cuda")
    for s in range(1000):
        r, _, d = ve.step_tensor(a_dev[s])
        total += r
    assert bool(d.all())
    rewards = np.array([ve.envs[i].total_reward for i in range(100)])
    delivered = np.array([sum(p.status == "delivered" for p in ve.envs[i].packages) for i in range(100)])
    assert np.array_equal(rewards, k["readme_rewards"])
    assert np.array_equal(delivered, k["readme_delivered"])
    assert f"{rewards.mean():.2f} {rewards.std():.2f}" == "-16.50 6.24"
    assert f"{delivered.mean():.2f} {delivered.std():.2f}" == "13.66 3.63"
    assert np.allclose(total.cpu().numpy(), rewards, atol=1e-9)


def test_notebook_dummy_shaping_known_answers():
    prev = {"robots": [(2, 2, 0), (5, 5, 1)], "time_step": 4}
    cur = {"robots": [(3, 2, 2), (5, 4, 0)], "time_step": 5}
    acts = [("D", "1"), ("L", "2")]
    pp = {1: {"id": 1, "start_pos": (0, 0), "target_pos": (4, 3), "start_time": 0, "deadline": 10, "status": "in_transit"},
          2: {"id": 2, "start_pos": (2, 1), "target_pos": (0, 4), "start_time": 3, "deadline": 15, "status": "waiting"}}
    m = mdb.helper.compute_shaped_rewards(10, prev, cur, acts, pp, 2)
    q = mdb.helper.compute_shaped_rewards(10, prev, cur, acts, pp, 2, shaping=mdb.QMIX_SHAPING)
    assert type(m) is np.float32 and float(m) == 215.04000854492188
    assert float(q) == 12.539999961853027


def _random_case(rs, trial):
    name = ["map.txt", "map1.txt", "map2.txt", "map3.txt", "map4.txt", "map5.txt"][trial % 6]
    grid = mdb.maps.stock_map(name)
    nfree = int((grid == 0).sum())
    dense = trial % 3 == 0
    R = int(rs.randint(max(1, min(20, nfree // 2)), min(32, nfree) + 1)) if dense else int(rs.randint(1, 9))
    P = int(rs.randint(1, 70))
    T = int(rs.randint(8, 60))
    K_o = int(rs.randint(0, R + 2))
    K_p = int(rs.randint(0, 9))
    MR = int(rs.randint(1, 40))
    MP = int(rs.randint(0, 40))
    return grid, R, P, T, K_o, K_p, MR, MP


@pytest.mark.parametrize("trial", range(18))
def test_random_configs_against_oracle(trial):
    """Hypothesis-style sweep: random map / robots (incl. dense 20-32 robot cases with swaps, chains and
    cycles) / packages / feature sizes, 6 envs each, random actions with a few illegal codes, fused
    auto-reset path vs the Python oracle, every step, every tensor."""
    rs = np.random.RandomState(1000 + trial)
    grid, R, P, T, K_o, K_p, MR, MP = _random_case(rs, trial)
    N = 6
    seed = int(rs.randint(0, 2 ** 30))
    consts = orc.QMIX_SHAPING if trial % 2 else orc.MAPPO_SHAPING
    ve = orc.OracleVecEnv(N, grid, seed=seed, max_time_steps=T, n_robots=R, n_packages=P)
    ve.reset()
    b = mdb.DeliveryBatch([grid], N, R, P, T, base_seed=seed, k_others=K_o, k_packages=K_p, max_robots_in_state=MR,
                          max_packages_in_state=MP, shaping=tuple(float(x) for x in consts))
    b.reset()
    steps = 2 * T + 5
    for s in range(steps):
        a = rs.randint(0, 15, size=(N, R))
        a = np.where(rs.rand(N, R) < 0.04, rs.randint(15, 18, size=(N, R)), a).astype(np.uint8)
        before = [e.snapshot() for e in ve.envs]
        want_r, want_sh, want_d = np.zeros(N), np.zeros(N, np.float32), np.zeros(N, bool)
        for n, e in enumerate(ve.envs):
            r_py, want_d[n] = e.step(a[n])  # python float/int like the reference returns (numpy>=2 weak promotion)
            want_r[n] = r_py
            want_sh[n] = orc.shaped_reward(e, r_py, before[n], e.snapshot(), a[n], consts)
            if want_d[n]:
                e.reset()
        r, sh, d, o, v, g, gv = b.step_featurize(torch.from_numpy(a), auto_reset=True)
        assert np.array_equal(r.cpu().numpy(), want_r), s
        assert np.array_equal(d.cpu().numpy().astype(bool), want_d), s
        assert np.array_equal(_bits(sh.cpu().numpy()), _bits(want_sh)), (s, sh.cpu().numpy(), want_sh)
        o, v, g, gv = o.cpu().numpy(), v.cpu().numpy(), g.cpu().numpy(), gv.cpu().numpy()
        for n, e in enumerate(ve.envs):
            snap = e.snapshot()
            st = b.get_state(n) if s % 7 == 0 else None
            if st is not None:
                assert (st["robot_row"] * e.W + st["robot_col"]).tolist() == e.rcell and st["robot_carry"].tolist() == e.carry
                assert st["pkg_status"].tolist() == e.p_status and st["t"] == e.t and st["total_reward"] == e.total_reward
            wo = np.stack([orc.observation(e, snap, i) for i in range(R)])
            wv = np.stack([orc.vector_features(e, snap, i, K_o, K_p) for i in range(R)])
            wg, wgv = orc.global_state(e, snap, MR, MP)
            assert np.array_equal(o[n], wo), (s, n)
            assert np.array_equal(_bits(v[n]), _bits(wv)), (s, n)
            assert np.array_equal(g[n], wg), (s, n)
            assert np.array_equal(_bits(gv[n]), _bits(wgv)), (s, n)
    assert int(b.episode_stats().cpu()[7]) == 0


def test_heterogeneous_maps_padded_planes():
    """C4-style batch mixing map1 (10x10) with 20x20 maps: per-env map id, planes padded to 20x20 with the
    map in the top-left corner, per-env n_rows/n_cols normalisation.  Parity is checked
    on the [:H, :W] sub-block against the oracle; the padding must be zero."""
    names = ["map1.txt", "map2.txt", "map3.txt", "map4.txt", "map5.txt"]
    grids = [mdb.maps.stock_map(n) for n in names]
    N, R, P, T = 10, 10, 40, 50
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file=names, n_robots=R, n_packages=P, max_time_steps=T, seed=77,
                           k_others=9, k_packages=5)
    envs = [orc.OracleEnv(grids[i % 5], T, R, P, seed=77 + i) for i in range(N)]
    ve.reset()
    for e in envs:
        e.reset()
    rs = np.random.RandomState(3)
    for s in range(2 * T + 3):
        a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
        want = []
        for n, e in enumerate(envs):
            before = e.snapshot()
            r, d = e.step(a[n])
            sh = orc.shaped_reward(e, r, before, e.snapshot(), a[n])
            if d:
                e.reset()
            want.append((r, d, sh))
        r, sh, d, o, v, g, gv = ve.rollout_step(torch.from_numpy(a))
        o, v, g, gv = o.cpu().numpy(), v.cpu().numpy(), g.cpu().numpy(), gv.cpu().numpy()
        for n, e in enumerate(envs):
            assert float(r[n]) == want[n][0] and bool(d[n]) == want[n][1] and float(sh[n]) == float(want[n][2]), (s, n)
            snap = e.snapshot()
            H, W = e.H, e.W
            wo = np.stack([orc.observation(e, snap, i) for i in range(R)])
            wg, wgv = orc.global_state(e, snap)
            wv = np.stack([orc.vector_features(e, snap, i, 9, 5) for i in range(R)])
            assert np.array_equal(o[n][:, :, :H, :W], wo) and o[n].sum() == wo.sum(), (s, n)
            assert np.array_equal(g[n][:, :H, :W], wg) and g[n].sum() == wg.sum(), (s, n)
            assert np.array_equal(_bits(v[n]), _bits(wv)), (s, n)
            assert np.array_equal(_bits(gv[n]), _bits(wgv)), (s, n)


def test_vectorized_env_reference_api_with_indices():
    """QMIX/env_vectorized.py:13-37: reset(indices) / step(actions, indices) return lists for the
    selected envs only; untouched envs keep their state; stepping a finished env is allowed (done is
    not latched)."""
    grid = mdb.maps.stock_map("map1.txt")
    ve = mdb.VectorizedEnv(mdb.Environment, 4, map_file="map1.txt", n_robots=3, n_packages=6, max_time_steps=5, seed=9)
    ov = orc.OracleVecEnv(4, grid, seed=9, max_time_steps=5, n_robots=3, n_packages=6)
    states = ve.reset()
    ov.reset()
    assert len(states) == 4 and [s == e.state_dict() for s, e in zip(states, ov.envs)] == [True] * 4
    rs = np.random.RandomState(4)
    moves, ops = "DLRSU", "012"
    for it in range(12):
        idx = [1, 3] if it % 2 else [0, 1, 2, 3]
        a = rs.randint(0, 15, size=(len(idx), 3))
        acts = [[(moves[x % 5], ops[x // 5]) for x in row] for row in a]
        st, r, d, info = ve.step(acts, idx if it % 2 else None)
        wr, wd = ov.step(a, idx)
        assert r == wr and d == [bool(x) for x in wd]
        for k, i in enumerate(idx):
            assert st[k] == ov.envs[i].state_dict()
            assert (info[k] == {}) == (not wd[k])
            if wd[k]:
                assert info[k] == {"total_reward": ov.envs[i].total_reward, "total_time_steps": ov.envs[i].t}
        if it == 7:
            got = ve.reset([2])
            ov.reset([2])
            assert got == [ov.envs[2].state_dict()]
    assert ve.envs[0].t == ov.envs[0].t > 5  # kept stepping past T
    assert [p.status for p in ve.envs[3].packages] == [mdb.batch.STATUS_NAMES[s] for s in ov.envs[3].p_status]


def test_helper_dropins_match_golden():
    """The per-call helper functions (same signatures as MAPPO/helper.py) on state/tracker dicts."""
    meta, z = load_traj("map7_home")
    e = orc.OracleEnv(z["grid"], meta["T"], meta["R"], meta["P"], seed=meta["seed"], **reward_kw(meta))
    e.reset()
    fi = 0
    for step in range(60):
        e.step(z["actions"][step])
        if z["done"][step]:
            e.reset()
        if step % 6 == 0:
            st = e.state_dict()
            tracker = {j + 1: {"id": j + 1, "start_pos": divmod(e.p_start[j], e.W), "target_pos": divmod(e.p_target[j], e.W),
                               "start_time": e.p_t0[j], "deadline": e.p_deadline[j],
                               "status": "waiting" if e.p_status[j] == 1 else "in_transit"}
                       for j in range(e.P) if e.p_status[j] in (1, 2)}
            for i in range(meta["R"]):
                assert np.array_equal(mdb.helper.convert_observation(st, tracker, i), z["obs"][step][i].astype(np.float32))
                got = mdb.helper.generate_vector_features(st, tracker, i, meta["T"], meta["K_o"], meta["K_p"])
                assert np.array_equal(_bits(got), _bits(z["vec"][step][i]))
            g, gv = mdb.helper.convert_global_state(st, tracker, meta["T"], meta["MR"], meta["MP"])
            assert np.array_equal(g, z["gstate"][step].astype(np.float32)) and np.array_equal(_bits(gv), _bits(z["gvec"][step]))
    assert mdb.helper.generate_vector_features(st, tracker, 99, 10).shape == (1007,)


@pytest.mark.parametrize("name", ["map1_mix", "map5_c4", "map1_home"])
def test_generic_featuriser_path_matches_golden(name, monkeypatch):
    """The kernels pick a plane-sliced fast path for H*W <= 512; MDX_GENERIC=1 forces the generic flat
    path (any H*W, e.g. maps larger than 22x22), which must produce the same bits."""
    monkeypatch.setenv("MDX_GENERIC", "1")
    meta, z = load_traj(name)
    b = _batch_for(meta, z, n_envs=2)
    monkeypatch.delenv("MDX_GENERIC")
    b.reset()
    fi = 0
    for step in range(min(meta["n_steps"], 260)):
        a = torch.from_numpy(np.repeat(z["actions"][step][None], 2, 0).copy())
        r, sh, d, o, v, g, gv = b.step_featurize(a, auto_reset=True)
        assert np.all(sh.cpu().numpy() == z["shaped_mappo"][step]), step
        if step % meta["feat_every"] == 0:
            assert np.array_equal(o.cpu().numpy()[1], z["obs"][fi].astype(np.float32)), step
            assert np.array_equal(_bits(v.cpu().numpy()[1]), _bits(z["vec"][fi])), step
            assert np.array_equal(g.cpu().numpy()[0], z["gstate"][fi].astype(np.float32)), step
            assert np.array_equal(_bits(gv.cpu().numpy()[1]), _bits(z["gvec"][fi])), step
            fi += 1


@pytest.mark.parametrize("lpe", [8, 16, 32])
@pytest.mark.parametrize("name", ["map1_mix", "map7_odd", "map7_home"])
def test_lanes_per_env_variants_match_golden(name, lpe, monkeypatch):
    """The kernels pack 4 / 2 / 1 environments per warp (8 / 16 / 32 lanes per env) depending on the
    robot count; MDX_LPE forces a larger group so that every variant is checked on the same trajectory.
    An odd env count leaves a lane group without an env (it must simply leave)."""
    monkeypatch.setenv("MDX_LPE", str(lpe))
    meta, z = load_traj(name)
    N = 7
    b = _batch_for(meta, z, n_envs=N)
    monkeypatch.delenv("MDX_LPE")
    b.reset()
    fi = 0
    for step in range(min(meta["n_steps"], 200)):
        a = torch.from_numpy(np.repeat(z["actions"][step][None], N, 0).copy())
        r, sh, d, o, v, g, gv = b.step_featurize(a, auto_reset=True)
        assert np.all(r.cpu().numpy() == z["reward"][step]) and np.all(d.cpu().numpy() == z["done"][step]), step
        assert np.all(sh.cpu().numpy() == z["shaped_mappo"][step]), step
        if step % meta["feat_every"] == 0:
            o, v, g, gv = o.cpu().numpy(), v.cpu().numpy(), g.cpu().numpy(), gv.cpu().numpy()
            for i in (0, 3, 6):
                assert np.array_equal(o[i], z["obs"][fi].astype(np.float32)), (step, i)
                assert np.array_equal(_bits(v[i]), _bits(z["vec"][fi])), (step, i)
                assert np.array_equal(g[i], z["gstate"][fi].astype(np.float32)), (step, i)
                assert np.array_equal(_bits(gv[i]), _bits(z["gvec"][fi])), (step, i)
            fi += 1


@pytest.mark.parametrize("name,R,P,T,K_o,K_p,MR,MP", [
    ("map1.txt", 1, 1, 1, 0, 1, 1, 1),          # smallest everything: one robot, one package, one step per episode
    ("map1.txt", 1, 0, 5, 0, 3, 2, 2),          # no packages at all: `all delivered` holds vacuously, done at every step
    ("map5.txt", 3, 80, 30, 2, 8, 5, 30),       # long list, > 5 slots: shared-memory key array path
    ("map2.txt", 2, 120, 40, 1, 3, 3, 10),      # long list, <= 5 slots: register top-k path, no key array
    ("map3.txt", 4, 100, 30, 100, 100, 100, 100),  # K = 100 defaults on a long list: rank-count path
    ("map5.txt", 32, 64, 20, 31, 5, 32, 64),    # full warp of robots, list that just fits the lane-held selection
    ("map4.txt", 8, 65, 25, 7, 6, 8, 65),       # first size past the 64-package register path
])
def test_edge_shapes_against_oracle(name, R, P, T, K_o, K_p, MR, MP):
    """Shapes at the edges of every selection / staging path of the featuriser, fused auto-reset launch vs the oracle,
    every step, every tensor, bit for bit."""
    grid = mdb.maps.stock_map(name)
    N = 5
    ve = orc.OracleVecEnv(N, grid, seed=77, max_time_steps=T, n_robots=R, n_packages=P)
    ve.reset()
    b = mdb.DeliveryBatch([grid], N, R, P, T, base_seed=77, k_others=K_o, k_packages=K_p, max_robots_in_state=MR,
                          max_packages_in_state=MP)
    b.reset()
    rs = np.random.RandomState(3)
    for s in range(2 * T + 3):
        a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
        if s % 3 == 1:
            a = (a % 5 + 5).astype(np.uint8)  # everybody tries to pick up
        wsh, wd, wo, wv, wg, wgv = orc.mappo_rollout_step(ve, a, K_o, K_p, MR, MP)
        r, sh, d, o, v, g, gv = b.step_featurize(torch.from_numpy(a), auto_reset=True)
        assert np.array_equal(_bits(sh.cpu().numpy()), _bits(wsh)) and np.array_equal(d.cpu().numpy().astype(bool), wd), s
        assert np.array_equal(o.cpu().numpy(), wo) and np.array_equal(g.cpu().numpy(), wg), s
        assert np.array_equal(_bits(v.cpu().numpy()), _bits(wv)) and np.array_equal(_bits(gv.cpu().numpy()), _bits(wgv)), s
    assert int(b.episode_stats().cpu()[7]) == 0


def test_caller_owned_outputs_are_validated():
    """A wrong dtype / size / layout of a caller-owned output tensor is refused before any kernel writes it."""
    b = mdb.DeliveryBatch([mdb.maps.stock_map("map1.txt")], 8, 5, 20, 50, base_seed=3, k_others=4, k_packages=5)
    b.reset()
    a = torch.zeros((8, 5), dtype=torch.uint8, device="cuda")
    good = (torch.empty(8, dtype=torch.float64, device="cuda"), torch.empty(8, dtype=torch.float32, device="cuda"),
            torch.empty(8, dtype=torch.uint8, device="cuda"))
    b.step(a, out=good)
    with pytest.raises(ValueError):
        b.step(a, out=(good[0], torch.empty(8, dtype=torch.float64, device="cuda"), good[2]))  # shaped must be float32
    with pytest.raises(ValueError):
        b.step(a, out=(good[0], good[1], torch.empty(7, dtype=torch.uint8, device="cuda")))  # too short
    with pytest.raises(ValueError):
        b.step(a, out=(good[0], good[1], torch.empty(8, dtype=torch.uint8)))  # host tensor
    feat = b.alloc_features(8)
    with pytest.raises(ValueError):
        b.featurize(out=(feat[0][:, :, :, :, :5], feat[1], feat[2], feat[3]))  # wrong size and not contiguous
    with pytest.raises(ValueError):
        b.featurize(out=(feat[0].transpose(3, 4), feat[1], feat[2], feat[3]))  # right size, not contiguous
    b.featurize(out=feat)
