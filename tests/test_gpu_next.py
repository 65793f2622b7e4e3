"""GPU parity tests of the callers either side of the path: the batched greedy
policy on BFS tables (f2), int64 policy actions / GAE / env-state checkpointing (f4) and the opt-in
BFS-distance shaping mode.  Everything goes through the C ABI (libmdx.so)."""
import json
import os

import numpy as np
import pytest
import torch

import marl_delivery_b200 as mdb
from golden_util import GOLDEN
from oracle import oracle_env as orc
from oracle.greedy_oracle import GreedyOracle

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


# ---------------------------------------------------------------------------------------------- f2 greedy
def test_bfs_table_matches_breadth_first_search():
    for name in ("map.txt", "map1.txt", "map5.txt", "map2.txt"):
        grid = mdb.maps.stock_map(name)
        b = mdb.DeliveryBatch([grid], 1, 2, 4, 10)
        assert np.array_equal(b.bfs_table(0), orc.bfs_all_pairs(grid)), name
    # heterogeneous batch: tables live in the padded output plane
    g1, g5 = mdb.maps.stock_map("map1.txt"), mdb.maps.stock_map("map5.txt")
    b = mdb.DeliveryBatch([g1, g5], 2, 2, 4, 10, env_map=[0, 1])
    t1 = b.bfs_table(0).reshape(20, 20, 20, 20)[:10, :10, :10, :10].reshape(100, 100)
    assert np.array_equal(t1, orc.bfs_all_pairs(g1))
    assert np.array_equal(b.bfs_table(1), orc.bfs_all_pairs(g5))


def test_greedy_golden_traces():
    """Per-step actions and rewards recorded from the unmodified GreedyAgents + Environment."""
    z = np.load(os.path.join(GOLDEN, "kats_greedy.npz"))
    for k, m in enumerate(json.loads(str(z["greedy_meta"]))):
        b = mdb.DeliveryBatch([z[f"greedy_grid_{k}"]], 1, m["R"], m["P"], m["T"], seeds=[m["seed"]])
        b.reset()
        want_a, want_r = z[f"greedy_actions_{k}"], z[f"greedy_rewards_{k}"]
        for s in range(m["steps"]):
            a = b.greedy_actions()
            assert np.array_equal(a.cpu().numpy()[0], want_a[s]), (m, s)
            r, _, d = b.step(a, want_shaped=False)
            assert float(r.cpu()[0]) == want_r[s], (m, s)
        assert bool(d.cpu()[0]) and b.get_state(0)["total_reward"] == m["total"]


def test_readme_greedy_row_on_the_gpu():
    """README.md:117-121 Greedy row (evaluation.py: seed 10, map1, 5 robots, 100 packages, T = 1000, 100
    episodes): the 100 episodes are the 100 envs of one batch (env i = seed 10 + i, second reset block)."""
    z = np.load(os.path.join(GOLDEN, "kats_greedy.npz"))
    ve = mdb.VectorizedEnv(mdb.Environment, 100, map_file="map1.txt", max_time_steps=1000, n_robots=5, n_packages=100, seed=10)
    b = ve.batch
    b.reset()
    total = torch.zeros(100, dtype=torch.float64, device="cuda")
    finished = torch.zeros(100, dtype=torch.bool, device="cuda")
    rec = []
    for s in range(1000):
        a = b.greedy_actions()
        if s < 1000:
            rec.append(a[:3].cpu().numpy().copy())
        r, _, d = b.step(a, want_shaped=False)
        total += torch.where(finished, torch.zeros_like(r), r)
        finished |= d.bool()
    assert bool(finished.all())
    rec = np.stack(rec, 1)
    for ep in range(3):
        n = len(z[f"readme_greedy_actions_ep{ep}"])
        assert np.array_equal(rec[ep][:n], z[f"readme_greedy_actions_ep{ep}"]), ep
    delivered = np.array([int((b.get_state(i)["pkg_status"] == 3).sum()) for i in range(100)])
    rewards = np.array([b.get_state(i)["total_reward"] for i in range(100)])
    assert np.array_equal(delivered, z["readme_greedy_delivered"])
    assert np.array_equal(rewards, z["readme_greedy_rewards"])
    assert f"{rewards.mean():.2f} {rewards.std():.2f}" == "34.04 14.83"  # README.md:119
    assert f"{delivered.mean():.2f} {delivered.std():.2f}" == "3.79 1.55"


@pytest.mark.parametrize("maps,R,P,T", [(["map1.txt"], 5, 50, 60), (["map5.txt"], 8, 70, 90), (["map.txt"], 3, 10, 40),
                                        (["map2.txt", "map4.txt", "map5.txt"], 12, 100, 80), (["map5.txt"], 32, 300, 50)])
def test_greedy_vs_oracle_with_auto_reset(maps, R, P, T):
    """Fresh seeds, several episodes per env (on-device auto-reset -> the agent restarts at t = 0)."""
    grids = [mdb.maps.stock_map(m) for m in maps]
    N = 24
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file=maps if len(maps) > 1 else maps[0], n_robots=R, n_packages=P,
                           max_time_steps=T, seed=77)
    b = ve.batch
    b.reset()
    envs = [orc.OracleEnv(grids[i % len(grids)], T, R, P, seed=77 + i) for i in range(N)]
    agents = []
    for e in envs:
        e.reset()
        agents.append(GreedyOracle(e))
    for s in range(2 * T + 7):
        a = b.greedy_actions()
        an = a.cpu().numpy()
        r, _, d = b.step(a, auto_reset=True, want_shaped=False)
        rn, dn = r.cpu().numpy(), d.cpu().numpy()
        for n, e in enumerate(envs):
            want = agents[n].get_actions()
            assert list(an[n]) == want, (s, n, list(an[n]), want)
            rr, dd = e.step(want)
            assert rn[n] == rr and bool(dn[n]) == dd, (s, n)
            if dd:
                e.reset()
                agents[n] = GreedyOracle(e)


def test_greedy_agents_drop_in():
    """greedyagent.GreedyAgents API on a single Environment, like evaluation.py:9-66."""
    z = np.load(os.path.join(GOLDEN, "kats_greedy.npz"))
    env = mdb.Environment("map1.txt", max_time_steps=1000, n_robots=5, n_packages=100, seed=10)
    state = env.reset()
    ag = mdb.greedyagent.GreedyAgents()
    ag.init_agents(state)
    want = z["readme_greedy_actions_ep0"]
    for s in range(40):
        acts = ag.get_actions(state)
        assert [mdb.encode_action(m, o) for m, o in acts] == list(want[s]), s
        state, r, done, info = env.step(acts)
    with pytest.raises(TypeError):
        ag.get_actions(dict(state))


def test_greedy_subset_of_envs():
    b = mdb.DeliveryBatch([mdb.maps.stock_map("map5.txt")], 6, 4, 30, 50, base_seed=3)
    b2 = mdb.DeliveryBatch([mdb.maps.stock_map("map5.txt")], 6, 4, 30, 50, base_seed=3)
    for s in range(30):
        full = b.greedy_actions()
        part = b2.greedy_actions(ids=[4, 1])
        assert torch.equal(part, full[[4, 1]]), s
        b.step(full, want_shaped=False)
        other = b2.greedy_actions(ids=[0, 2, 3, 5])
        assert torch.equal(other, full[[0, 2, 3, 5]]), s
        b2.step(full, want_shaped=False)


# ---------------------------------------------------------------------------------------------- BFS shaping (opt-in)
def test_bfs_shaping_equals_manhattan_without_inner_walls():
    kw = dict(map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=80, seed=9, k_others=4, k_packages=5)
    a_env = mdb.VectorizedEnv(mdb.Environment, 64, **kw)
    b_env = mdb.VectorizedEnv(mdb.Environment, 64, bfs_shaping=True, **kw)
    gen = torch.Generator().manual_seed(1)
    for ve in (a_env, b_env):
        ve.batch.reset()
    for s in range(120):
        a = torch.randint(0, 15, (64, 5), dtype=torch.uint8, generator=gen)
        x = a_env.rollout_step(a)
        y = b_env.rollout_step(a)
        for u, v in zip(x, y):
            assert torch.equal(u, v), s


@pytest.mark.parametrize("mapname,R,P", [("map5.txt", 6, 60), ("map2.txt", 5, 40), ("map5.txt", 20, 120)])
def test_bfs_shaping_vs_oracle_variant(mapname, R, P):
    grid = mdb.maps.stock_map(mapname)
    table = orc.bfs_all_pairs(grid)
    N, T = 16, 70
    for consts, name in ((orc.MAPPO_SHAPING, mdb.MAPPO_SHAPING), (orc.QMIX_SHAPING, mdb.QMIX_SHAPING)):
        ve = mdb.VectorizedEnv(mdb.Environment, N, map_file=mapname, n_robots=R, n_packages=P, max_time_steps=T, seed=21,
                               shaping=name, bfs_shaping=True)
        ve.batch.reset()
        ov = orc.OracleVecEnv(N, grid, seed=21, max_time_steps=T, n_robots=R, n_packages=P)
        ov.reset()
        rs = np.random.RandomState(4)
        differs = 0
        for s in range(T):
            a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
            r, sh, d = ve.batch.step(torch.from_numpy(a))
            sh = sh.cpu().numpy()
            for n, e in enumerate(ov.envs):
                before = e.snapshot()
                rr, dd = e.step(a[n])
                after = e.snapshot()
                want = orc.shaped_reward(e, rr, before, after, a[n], consts, bfs=table)
                assert sh[n] == want, (s, n, sh[n], want)
                differs += want != orc.shaped_reward(e, rr, before, after, a[n], consts)
        assert differs > 0  # the mode is not vacuous on a map with walls
    # the standalone shaped-reward entry point honours the mode too (helper drop-in path)


# ---------------------------------------------------------------------------------------------- f4 rollout glue
def test_int64_policy_actions_equal_action_codes():
    kw = dict(map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=60, seed=5, k_others=4, k_packages=5)
    a_env = mdb.VectorizedEnv(mdb.Environment, 128, **kw)
    b_env = mdb.VectorizedEnv(mdb.Environment, 128, **kw)
    for ve in (a_env, b_env):
        ve.batch.reset()
    gen = torch.Generator(device="cuda").manual_seed(2)
    for s in range(90):
        a64 = torch.randint(0, 15, (128, 5), dtype=torch.int64, device="cuda", generator=gen)
        if s % 7 == 3:
            a64[::5] += 15  # package-op index >= 3 is clamped to '0' (MAPPO/trainer.py:199-201)
        codes = torch.where(a64 < 15, a64, a64 % 5).to(torch.uint8)
        x = a_env.rollout_step(a64)
        y = b_env.rollout_step(codes)
        for u, v in zip(x, y):
            assert torch.equal(u, v), s
        u = a_env.batch.step(a64, auto_reset=True)
        v = b_env.batch.step(codes, auto_reset=True)
        for p, q in zip(u, v):
            assert torch.equal(p, q), s


@pytest.mark.parametrize("no_tma", [False, True])
@pytest.mark.parametrize("T,N", [(1, 1), (7, 3), (500, 5), (64, 4097), (500, 65536), (1, 16), (3, 128), (10, 144), (41, 1040), (39, 4096),
                                 (127, 2064)])
def test_gae_matches_torch_reference(T, N, no_tma, monkeypatch):
    """MAPPO/trainer.py:288-301.  Tolerance 1e-6 relative as the stated bar; equality is what holds.  N % 16 == 0 takes the
    bulk-async-copy (TMA) kernel -- ragged last CTAs (144, 1040, 2064 columns), T shorter than / not a multiple of the stage
    depth -- everything else, and MDX_GAE_NO_TMA=1, the register kernel."""
    if no_tma:
        monkeypatch.setenv("MDX_GAE_NO_TMA", "1")
    g = torch.Generator().manual_seed(T * 131 + N)
    rew = torch.randn(T, N, generator=g) * 3
    val = torch.randn(T, N, generator=g)
    done = torch.rand(T, N, generator=g) < 0.02
    last = torch.randn(N, generator=g)
    adv, ret = mdb.compute_gae(rew.cuda(), val.cuda(), done.cuda(), last.cuda(), 0.99, 0.95)
    wadv, wret = orc.gae_reference(rew, val, done, last, 0.99, 0.95)
    assert torch.allclose(adv.cpu(), wadv, rtol=1e-6, atol=1e-6) and torch.allclose(ret.cpu(), wret, rtol=1e-6, atol=1e-6)
    assert torch.equal(adv.cpu(), wadv) and torch.equal(ret.cpu(), wret)


def test_checkpoint_resume_is_bit_identical():
    kw = dict(map_file=["map2.txt", "map5.txt"], n_robots=6, n_packages=40, max_time_steps=30, seed=13, k_others=5, k_packages=5)
    ve = mdb.VectorizedEnv(mdb.Environment, 96, **kw)
    b = ve.batch
    b.reset()
    gen = torch.Generator().manual_seed(8)
    acts = torch.randint(0, 15, (110, 96, 6), dtype=torch.uint8, generator=gen)
    for s in range(25):
        b.step_featurize(acts[s])
    for s in range(3):  # greedy bookkeeping is part of the state
        b.step(b.greedy_actions(), auto_reset=True)
    blob = b.state_dict()
    assert blob.dtype == torch.uint8 and blob.numel() > 96 * 2500

    def run(batch):
        out = []
        for s in range(25, 110):
            if s % 2:
                res = batch.step_featurize(acts[s])
            else:
                res = batch.step(batch.greedy_actions(), auto_reset=True) + tuple(batch.featurize())
            out.append([t.clone() for t in res])
        return out, batch.episode_stats().clone()

    want, wstats = run(b)
    b.load_state_dict(blob)  # same instance, rewound
    got, gstats = run(b)
    ve2 = mdb.VectorizedEnv(mdb.Environment, 96, **kw)  # fresh instance
    ve2.batch.load_state_dict(blob)
    got2, gstats2 = run(ve2.batch)
    for g_ in (got, got2):
        for x, y in zip(want, g_):
            for u, v in zip(x, y):
                assert torch.equal(u, v)
    assert torch.equal(wstats, gstats) and torch.equal(wstats, gstats2)
    assert int(wstats[0]) >= 96 * 2 and int(wstats[7]) == 0  # several auto-resets happened, nobody starved
    with pytest.raises(ValueError):
        mdb.DeliveryBatch([mdb.maps.stock_map("map1.txt")], 4, 5, 20, 50).load_state_dict(blob)


def test_collect_rollouts_matches_trainer_loop():
    """MAPPO/trainer.py:154-318 with a deterministic stand-in policy: every rollout buffer, the advantages and the
    returns equal what the trainer's loop computes from the reference path (oracle) on the same actions."""
    N, R, P, T_env, K = 6, 5, 50, 25, 60  # 60 rollout steps over 25-step episodes: two auto-resets per env
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file="map1.txt", n_robots=R, n_packages=P, max_time_steps=T_env, seed=91,
                           k_others=R - 1, k_packages=5)
    ve.batch.reset()
    gen = torch.Generator(device="cuda").manual_seed(17)

    def policy(obs, vec, gstate, gvec):
        a = torch.randint(0, 15, (N, R), dtype=torch.int64, device="cuda", generator=gen)
        return a, vec[:, :, 0] - 1.0, gvec[:, :30].sum(1) * 0.1  # any deterministic functions of the features

    def value(gstate, gvec):
        return gvec[:, :30].sum(1) * 0.1

    out = mdb.rollout.collect_rollouts(ve, policy, value, K, gamma=0.99, gae_lambda=0.95)
    b_obs, b_vec, b_gs, b_gv, b_act, b_lp, b_adv, b_ret, mb_rewards, buf = out
    assert b_obs.shape == (K * N * R, 6, 10, 10) and b_vec.shape == (K * N * R, 52)  # MAPPO/trainer.py:304-305
    assert b_gs.shape == (K * N, 4, 10, 10) and b_gv.shape == (K * N, 1301)
    assert b_act.shape == (K * N * R,) and b_adv.shape == (K * N * R,) and b_ret.shape == (K * N,)
    # the reference loop on the CPU
    grid = mdb.maps.stock_map("map1.txt")
    ov = orc.OracleVecEnv(N, grid, seed=91, max_time_steps=T_env, n_robots=R, n_packages=P)
    ov.reset()
    acts = buf.actions.cpu().numpy()
    snap = [e.snapshot() for e in ov.envs]
    o0 = np.stack([np.stack([orc.observation(e, s, i) for i in range(R)]) for e, s in zip(ov.envs, snap)])
    assert np.array_equal(buf.obs[0].cpu().numpy(), o0)
    for t in range(K):
        shaped, dones, o, v, g, gv = orc.mappo_rollout_step(ov, acts[t], R - 1, 5)
        assert np.array_equal(buf.rewards[t].cpu().numpy(), shaped), t
        assert np.array_equal(buf.dones[t].cpu().numpy().astype(bool), dones), t
        assert np.array_equal(buf.obs[t + 1].cpu().numpy(), o) and np.array_equal(_bits(buf.vector_obs[t + 1].cpu().numpy()), _bits(v)), t
        assert np.array_equal(buf.global_states[t + 1].cpu().numpy(), g) and np.array_equal(_bits(buf.global_vector[t + 1].cpu().numpy()), _bits(gv)), t
    assert int(buf.dones.sum()) == 2 * N
    wadv, wret = orc.gae_reference(buf.rewards.cpu(), buf.values.cpu(), buf.dones.cpu(), value(buf.global_states[K], buf.global_vector[K]).cpu(), 0.99, 0.95)
    assert torch.equal(b_adv.cpu(), wadv.reshape(K * N, 1).repeat(1, R).reshape(-1)) and torch.equal(b_ret.cpu(), wret.reshape(-1))
    assert torch.equal(b_lp, (buf.vector_obs[:K, :, :, 0] - 1.0).reshape(-1))


def test_collect_rollouts_over_two_env_ranges_equals_one_range():
    """parts = 2: each env range has its own stream for its policy evaluation and its fused env launch.  With a deterministic
    row-wise policy every buffer, the advantages and the returns equal the single-range rollout."""
    N, R, K = 64, 5, 45
    kw = dict(map_file="map1.txt", n_robots=R, n_packages=50, max_time_steps=20, seed=7, k_others=R - 1, k_packages=5)

    def policy(obs, vec, gstate, gvec):  # row-wise and deterministic: a hash of each env's own features
        h = (vec * 997.0).sum(2) + obs.flatten(2).sum(2) * 3.0
        a = (h.abs() * 31.0).long() % 15
        return a, vec[:, :, 1] * 0.5, gvec[:, :40].sum(1) * 0.01

    def value(gstate, gvec):
        return gvec[:, :40].sum(1) * 0.01

    outs = []
    for parts in (1, 2, 4):
        ve = mdb.VectorizedEnv(mdb.Environment, N, **kw)
        ve.batch.reset()
        res = mdb.rollout.collect_rollouts(ve, policy, value, K, parts=parts)
        torch.cuda.synchronize()
        outs.append([t.clone() for t in res[:9]] + [res[9].dones.clone(), res[9].env_rewards.clone(), ve.batch.episode_stats().clone()])
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert torch.equal(x, y)
    assert int(outs[0][9].sum()) == 2 * N  # two auto-resets per env inside the rollout


def test_host_pipeline_equals_synchronous_rollout_steps():
    """VectorizedEnv.host_pipeline(): host actions in, (shaped, done) out, copies overlapped with the kernel -- every
    step's results and the final features / episode statistics equal the step-by-step rollout_step() on the same actions,
    and both equal the oracle."""
    N, R, P, T_env, K = 64, 5, 50, 20, 50
    kw = dict(map_file="map1.txt", n_robots=R, n_packages=P, max_time_steps=T_env, seed=123, k_others=R - 1, k_packages=5)
    va, vb = mdb.VectorizedEnv(mdb.Environment, N, **kw), mdb.VectorizedEnv(mdb.Environment, N, **kw)
    for v in (va, vb):
        v.batch.reset()
    ov = orc.OracleVecEnv(N, mdb.maps.stock_map("map1.txt"), seed=123, max_time_steps=T_env, n_robots=R, n_packages=P)
    ov.reset()
    acts = np.random.RandomState(5).randint(0, 15, size=(K, N, R)).astype(np.uint8)
    h_acts = torch.from_numpy(acts).pin_memory()
    pipe = va.host_pipeline()
    got, prev = [], None
    for t in range(K):
        slot = pipe.submit(h_acts[t])
        if prev is not None:
            sh, dn = pipe.result(prev)
            got.append((sh.copy(), dn.copy()))
        prev = slot
    sh, dn = pipe.result(prev)
    got.append((sh.copy(), dn.copy()))
    pipe.drain()
    torch.cuda.synchronize()
    for t in range(K):
        _, wsh, wdn, *wfeat = vb.rollout_step(torch.from_numpy(acts[t]))
        osh, odn, *ofeat = orc.mappo_rollout_step(ov, acts[t], R - 1, 5)
        assert np.array_equal(_bits(got[t][0]), _bits(wsh.cpu().numpy())) and np.array_equal(got[t][1], wdn.cpu().numpy()), t
        assert np.array_equal(_bits(got[t][0]), _bits(osh)) and np.array_equal(got[t][1].astype(bool), odn), t
    for x, w, o in zip(pipe.feat, wfeat, ofeat):
        assert torch.equal(x, w) and np.array_equal(_bits(x.cpu().numpy()), _bits(o))
    assert torch.equal(va.batch.episode_stats(), vb.batch.episode_stats())
