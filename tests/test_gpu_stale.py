"""`stale_tracker_compat`: the CUDA path with MDX_CFG_STALE_TRACKER against (a) fixtures
recorded from the UNMODIFIED reference with the tracker driven exactly like MAPPO/trainer.py:95-130, 229-234 (never cleared
on auto-reset) -- tests/golden/traj_*_stale.npz, kats_stale.npz (KAT-B second sum 1243.89) -- and (b) the tracker oracle on
fresh seeded rollouts.  Everything is bit-exact: shaped reward (f32), all four feature tensors, and the dict itself
(ids in insertion order, statuses)."""
import os

import numpy as np
import pytest
import torch

import marl_delivery_b200 as mdb
from golden_util import GOLDEN, load_traj
from oracle import oracle_env as orc
from oracle import tracker_oracle as trk

pytestmark = pytest.mark.gpu

STALE_NAMES = ["map1_stale", "map7_stale", "map5_stale", "rect_stale"]  # rect_stale: a non-square 10x15 map


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


def _batch(meta, z, n=1, K=None, shaping=mdb.MAPPO_SHAPING, stale=True):
    K_o, K_p, MR, MP = K if K else (meta["K_o"], meta["K_p"], meta["MR"], meta["MP"])
    return mdb.DeliveryBatch([z["grid"]], n, meta["R"], meta["P"], meta["T"], seeds=[meta["seed"]] * n, k_others=K_o, k_packages=K_p,
                             max_robots_in_state=MR, max_packages_in_state=MP, shaping=shaping, stale_tracker=stale)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", STALE_NAMES)
def test_stale_golden_trajectory(name, fused):
    """fused: one launch per trainer iteration (auto-reset inside the kernel).  unfused: step / explicit reset / featurize,
    the order MAPPO/trainer.py:208-280 itself uses."""
    meta, z = load_traj(name)
    assert meta["stale"]
    N = 3 if fused else 1
    b = _batch(meta, z, N)
    bq = _batch(meta, z, N, shaping=mdb.QMIX_SHAPING)
    bk = _batch(meta, z, N, K=(100, 100, 10, 20)) if meta["big_vec_every"] else None
    for x in (b, bq, bk):
        if x is not None:
            x.reset()
    o, v, g, gv = [t.cpu().numpy() for t in b.featurize()]
    assert np.array_equal(o[0], z["obs0"].astype(np.float32)) and np.array_equal(_bits(v[0]), _bits(z["vec0"]))
    fi = bi = 0
    for s in range(meta["n_steps"]):
        a = torch.from_numpy(np.repeat(z["actions"][s][None], N, 0).copy())
        if fused:
            r, sh, d, o, v, g, gv = b.step_featurize(a, auto_reset=True)
            _, shq, _ = bq.step(a, auto_reset=True)
            if bk is not None:
                _, _, _, _, vb, _, gq = bk.step_featurize(a, auto_reset=True)
        else:
            r, sh, d = b.step(a)
            _, shq, _ = bq.step(a)
            if int(d.cpu()[0]):
                b.reset()
                bq.reset()
            o, v, g, gv = b.featurize()
            if bk is not None:
                _, _, dk = bk.step(a)
                if int(dk.cpu()[0]):
                    bk.reset()
                _, vb, _, gq = bk.featurize()
        assert np.all(r.cpu().numpy() == z["reward"][s]) and np.all(d.cpu().numpy() == z["done"][s]), s
        assert np.all(_bits(sh.cpu().numpy()) == _bits(z["shaped_mappo"][s:s + 1])), (s, sh, z["shaped_mappo"][s])
        assert np.all(_bits(shq.cpu().numpy()) == _bits(z["shaped_qmix"][s:s + 1])), s
        if s % 5 == 0 or z["done"][s]:
            for i in {0, N - 1}:
                tk = b.get_tracker(i)
                assert [e[0] for e in tk] == z["trk_ids"][s][:len(tk)].tolist() and not z["trk_ids"][s][len(tk):].any(), s
                assert [e[5] for e in tk] == z["trk_status"][s][:len(tk)].tolist(), s
        if s % meta["feat_every"] == 0:
            o, v, g, gv = o.cpu().numpy(), v.cpu().numpy(), g.cpu().numpy(), gv.cpu().numpy()
            for i in {0, N - 1}:
                assert np.array_equal(o[i], z["obs"][fi].astype(np.float32)), s
                assert np.array_equal(_bits(v[i]), _bits(z["vec"][fi])), s
                assert np.array_equal(g[i], z["gstate"][fi].astype(np.float32)), s
                assert np.array_equal(_bits(gv[i]), _bits(z["gvec"][fi])), s
            fi += 1
        if bk is not None and s % meta["big_vec_every"] == 0:
            assert np.array_equal(_bits(vb.cpu().numpy()[0]), _bits(z["vec_big"][bi])), s
            assert np.array_equal(_bits(gq.cpu().numpy()[0]), _bits(z["gvec_qmix"][bi])), s
            bi += 1


def test_kat_b_with_the_trainers_own_tracker_sums_to_1243_89():
    k = np.load(os.path.join(GOLDEN, "kats_stale.npz"))
    ve = mdb.VectorizedEnv(mdb.Environment, 5, map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=500, seed=42,
                           k_others=4, k_packages=5, stale_tracker_compat=True)
    ve.batch.reset()
    rs = np.random.RandomState(0)
    total = 0.0
    for s in range(1200):
        a = rs.randint(0, 15, size=(5, 5)).astype(np.uint8)
        r, sh, d, o, v, g, gv = ve.rollout_step(torch.from_numpy(a))
        sh = sh.cpu().numpy()
        assert np.array_equal(_bits(sh), _bits(k["kat_b_stale_shaped"][s])), s
        assert np.array_equal(d.cpu().numpy(), k["kat_b_stale_done"][s]), s
        fs = v.cpu().numpy()[:, 0].astype(np.float64).sum(axis=1) + gv.cpu().numpy().astype(np.float64).sum(axis=1)
        assert np.array_equal(fs, k["kat_b_stale_featsum"][s]), s
        total += float(sh.astype(np.float64).sum())
    assert abs(total - 1243.89) < 5e-3  # SURVEY.md KAT-B, second sum (cleared tracker: 1627.37)


@pytest.mark.parametrize("trial", range(8))
def test_stale_random_configs_against_tracker_oracle(trial):
    rs = np.random.RandomState(4000 + trial)
    name = ["map.txt", "map1.txt", "map3.txt", "map5.txt"][trial % 4]
    grid = mdb.maps.stock_map(name)
    nfree = int((grid == 0).sum())
    R = int(rs.randint(1, min(12, nfree) + 1)) if trial % 3 else int(rs.randint(12, min(32, nfree) + 1))
    P = int(rs.randint(2, 90))
    T = int(rs.randint(4, 30))
    K_o, K_p, MR, MP = int(rs.randint(0, R + 2)), int(rs.randint(0, 9)) if trial % 2 else 100, int(rs.randint(1, 40)), int(rs.randint(0, 40))
    N = 5
    seed = int(rs.randint(0, 2 ** 30))
    consts = orc.QMIX_SHAPING if trial % 2 else orc.MAPPO_SHAPING
    ov = orc.OracleVecEnv(N, grid, seed=seed, max_time_steps=T, n_robots=R, n_packages=P)
    ov.reset()
    tks = [trk.Tracker() for _ in range(N)]
    for t_, e in zip(tks, ov.envs):
        t_.update(e)
    b = mdb.DeliveryBatch([grid], N, R, P, T, base_seed=seed, k_others=K_o, k_packages=K_p, max_robots_in_state=MR,
                          max_packages_in_state=MP, shaping=tuple(float(x) for x in consts), stale_tracker=True)
    b.reset()
    for s in range(5 * T + 3):
        a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
        if s % 3 == 1:
            a = (a % 5 + 5).astype(np.uint8)  # everybody tries to pick up: many in_transit entries survive the resets' filter
        wsh, wd, wo, wv, wg, wgv = trk.mappo_rollout_step(ov, tks, a, K_o, K_p, MR, MP, consts)
        r, sh, d, o, v, g, gv = b.step_featurize(torch.from_numpy(a), auto_reset=True)
        assert np.array_equal(_bits(sh.cpu().numpy()), _bits(wsh)) and np.array_equal(d.cpu().numpy().astype(bool), wd), s
        assert np.array_equal(o.cpu().numpy(), wo) and np.array_equal(g.cpu().numpy(), wg), s
        assert np.array_equal(_bits(v.cpu().numpy()), _bits(wv)) and np.array_equal(_bits(gv.cpu().numpy()), _bits(wgv)), s
        if s % 7 == 0:
            for n in (0, N - 1):
                assert b.get_tracker(n) == tks[n].snapshot(), (s, n)
    assert any(len(t_.d) and list(t_.d) != sorted(t_.d) for t_ in tks) or T > 20  # the dict order really left id order


def test_clearing_the_tracker_before_each_reset_gives_the_env_truth():
    """QMIX/trainer.py:523-526 clears its tracker when an env resets: tracker_clear + reset == the default (truth) mode."""
    grid = mdb.maps.stock_map("map1.txt")
    kw = dict(base_seed=12, k_others=4, k_packages=5)
    a_ = mdb.DeliveryBatch([grid], 7, 5, 20, 15, **kw)
    b_ = mdb.DeliveryBatch([grid], 7, 5, 20, 15, stale_tracker=True, **kw)
    a_.reset()
    b_.reset()
    rs = np.random.RandomState(3)
    for s in range(70):
        a = torch.from_numpy(rs.randint(0, 15, size=(7, 5)).astype(np.uint8))
        ra, sa, da = a_.step(a)
        rb, sb, db = b_.step(a)
        assert torch.equal(sa, sb) and torch.equal(da, db), s
        done = np.flatnonzero(da.cpu().numpy()).tolist()
        if done:
            a_.reset(done)
            b_.tracker_clear(done)
            b_.reset(done)
        for x, y in zip(a_.featurize(), b_.featurize()):
            assert torch.equal(x, y), s
    with pytest.raises(ValueError):
        a_.tracker_clear()  # not a stale-tracker instance


def test_helper_drop_ins_accept_the_trainers_unordered_dict():
    """The per-call converters / shaping receive whatever dict the trainer holds; for the MAPPO trainer that dict is not in id
    order after the first auto-reset.  Results equal the tracker oracle (pinned to the reference above)."""
    grid = mdb.maps.stock_map("map1.txt")
    R, P, T = 5, 30, 12
    e = orc.OracleEnv(grid, T, R, P, seed=61)
    e.reset()
    tk = trk.Tracker()
    tk.update(e)
    rs = np.random.RandomState(8)
    unordered = 0
    for s in range(60):
        a = rs.randint(0, 15, size=R)
        if s % 2:
            a = a % 5 + 5
        before, pp_prev = e.snapshot(), tk.snapshot()
        prev_state = e.state_dict()
        r, d = e.step(a)
        after = e.snapshot()
        cur_state = e.state_dict()
        as_dict = lambda snap: {pid: dict(id=pid, start_pos=divmod(st, e.W), target_pos=divmod(tg, e.W), start_time=t0, deadline=dl,  # noqa: E731
                                          status="waiting" if stt == orc.WAITING else "in_transit") for (pid, st, tg, t0, dl, stt) in snap}
        acts = [("DLRSU"[x % 5], str(x // 5)) for x in a]
        got = mdb.helper.compute_shaped_rewards(r, prev_state, cur_state, acts, as_dict(pp_prev), R)
        assert _bits(np.float32(got)) == _bits(np.float32(trk.shaped_reward(e, r, before, after, a, pp_prev))), s
        if d:
            e.reset()
        tk.update(e)
        ts, snap, st = tk.snapshot(), e.snapshot(), e.state_dict()
        ids = [x[0] for x in ts]
        if ids != sorted(ids):
            unordered += 1
            pp = as_dict(ts)
            assert np.array_equal(mdb.helper.convert_observation(st, pp, 2), trk.observation(e, snap, ts, 2)), s
            assert np.array_equal(_bits(mdb.helper.generate_vector_features(st, pp, 1, T, R - 1, 5)), _bits(trk.vector_features(e, snap, ts, 1, R - 1, 5))), s
            gm, gvv = mdb.helper.convert_global_state(st, pp, T)
            wm, wv = trk.global_state(e, snap, ts)
            assert np.array_equal(gm, wm) and np.array_equal(_bits(gvv), _bits(wv)), s
    assert unordered > 5
