"""GPU parity at BASELINE.json sizes: the CUDA path against the C oracle on thousands of environments
over whole episodes, plus size-independent properties on the full 65 536-env C3 batch."""
import numpy as np
import pytest
import torch

import marl_delivery_b200 as mdb
from oracle import oracle_env as orc
from oracle_c import COracleEnv, COracleVec

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


def _compare_features(feat, cv, tag):
    o, v, g, gv = [x.cpu().numpy() for x in feat]
    assert np.array_equal(o, cv.obs), tag
    assert np.array_equal(_bits(v), _bits(cv.vec)), tag
    assert np.array_equal(g, cv.gstate), tag
    assert np.array_equal(_bits(gv), _bits(cv.gvec)), tag


def test_c2_c3_4096_envs_whole_episode_vs_c_oracle():
    """BASELINE configs[1]/[2] at 4096 envs: map1, 5 robots, 50 packages, T=500, 520 fused steps with
    on-device auto-reset (every env resets at step 500).  Reward (f64 ==), shaped (f32 ==) and done for all
    envs at every step; all four feature tensors of all envs every 40 steps and around the reset."""
    N, R, P, T = 4096, 5, 50, 500
    grid = mdb.maps.stock_map("map1.txt")
    cv = COracleVec(N, grid, T, R, P, 42, 4, 5, 100, 100, orc.MAPPO_SHAPING)
    cv.reset()
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file="map1.txt", n_robots=R, n_packages=P, max_time_steps=T, seed=42,
                           k_others=4, k_packages=5)
    ve.batch.reset()
    gen = torch.Generator().manual_seed(7)
    n_done = 0
    for s in range(520):
        a = torch.randint(0, 15, (N, R), dtype=torch.uint8, generator=gen)
        check = s % 40 == 0 or s in (498, 499, 500, 519)
        r, sh, d = cv.rollout_step(a.numpy(), features=check)
        gr, gsh, gd, *feat = ve.rollout_step(a)
        assert np.array_equal(gr.cpu().numpy(), r), s
        assert np.array_equal(_bits(gsh.cpu().numpy()), _bits(sh)), s
        assert np.array_equal(gd.cpu().numpy(), d), s
        n_done += int(d.sum())
        if check:
            _compare_features(feat, cv, s)
    assert n_done == N
    st = ve.batch.episode_stats().cpu().numpy()
    assert st[0] == N and st[5] == N * 520 and st[6] == N * 500 and st[7] == 0


def test_c5_stress_32_robots_500_packages_vs_c_oracle():
    """BASELINE configs[4] shape (map5, 32 robots = a full warp, 500 packages, T=1000, K_o=31) on 48 envs over a
    whole episode + reset."""
    N, R, P, T = 48, 32, 500, 1000
    grid = mdb.maps.stock_map("map5.txt")
    cv = COracleVec(N, grid, T, R, P, 123, 31, 5, 100, 100, orc.MAPPO_SHAPING)
    cv.reset()
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file="map5.txt", n_robots=R, n_packages=P, max_time_steps=T, seed=123,
                           k_others=31, k_packages=5)
    assert ve.batch.vec_dim == 187
    ve.batch.reset()
    gen = torch.Generator().manual_seed(11)
    for s in range(1030):
        a = torch.randint(0, 15, (N, R), dtype=torch.uint8, generator=gen)
        check = s % 64 == 0 or s in (999, 1000, 1029)
        r, sh, d = cv.rollout_step(a.numpy(), features=check)
        gr, gsh, gd, *feat = ve.rollout_step(a)
        assert np.array_equal(gr.cpu().numpy(), r) and np.array_equal(gd.cpu().numpy(), d), s
        assert np.array_equal(_bits(gsh.cpu().numpy()), _bits(sh)), s
        if check:
            _compare_features(feat, cv, s)


def test_c4_heterogeneous_2048_envs_vs_c_oracle():
    """BASELINE configs[3] shape: map1..map5 mixed by env index, 10 robots, 100 packages, planes padded to
    20x20; compared on the [:H, :W] sub-block, padding must be zero."""
    names = ["map1.txt", "map2.txt", "map3.txt", "map4.txt", "map5.txt"]
    grids = [mdb.maps.stock_map(n) for n in names]
    N, R, P, T = 2048, 10, 100, 120
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file=names, n_robots=R, n_packages=P, max_time_steps=T, seed=5,
                           k_others=9, k_packages=5)
    envs = [COracleEnv(grids[i % 5], T, R, P, 5 + i) for i in range(N)]
    for e in envs:
        e.reset()
    ve.batch.reset()
    consts = np.array(orc.MAPPO_SHAPING, np.float32)
    gen = torch.Generator().manual_seed(3)
    for s in range(130):
        a = torch.randint(0, 15, (N, R), dtype=torch.uint8, generator=gen)
        an = a.numpy()
        gr, gsh, gd, o, v, g, gv = ve.rollout_step(a)
        gr, gsh, gd = gr.cpu().numpy(), gsh.cpu().numpy(), gd.cpu().numpy()
        check = s % 43 == 0 or s in (119, 120)
        if check:
            o, v, g, gv = o.cpu().numpy(), v.cpu().numpy(), g.cpu().numpy(), gv.cpu().numpy()
        for n, e in enumerate(envs):
            r, d = e.step(an[n])
            assert gr[n] == r and bool(gd[n]) == d and gsh[n] == e.shaped(an[n], consts, r), (s, n)
            if d:
                e.reset()
            if check and n % 8 == s % 8:
                wo, wv, wg, wgv = e.features(9, 5, 100, 100)
                H, W = e.H, e.W
                assert np.array_equal(o[n][:, :, :H, :W], wo) and o[n].sum() == wo.sum(), (s, n)
                assert np.array_equal(g[n][:, :H, :W], wg) and g[n].sum() == wg.sum(), (s, n)
                assert np.array_equal(_bits(v[n]), _bits(wv)) and np.array_equal(_bits(gv[n]), _bits(wgv)), (s, n)


def test_c3_full_size_properties_and_shard_independence():
    """65 536 envs (BASELINE configs[2] per-GPU size).  Properties that hold for any correct featuriser,
    checked on every env; determinism; independence of the result from how envs are sharded over ranks
    (seeds are base + GLOBAL env index); spot check of 128 random envs against the C oracle."""
    N, R, P, T = 65536, 5, 50, 500
    mk = lambda n, first: mdb.VectorizedEnv(mdb.Environment, n, map_file="map1.txt", n_robots=R, n_packages=P,  # noqa: E731
                                            max_time_steps=T, seed=42, k_others=4, k_packages=5, first_env_index=first)
    full = mk(N, 0)
    lo, hi = mk(N // 2, 0), mk(N // 2, N // 2)
    for ve in (full, lo, hi):
        ve.batch.reset()
    grid = mdb.maps.stock_map("map1.txt")
    pick = np.sort(np.random.RandomState(0).choice(N, 128, replace=False))
    spot = [COracleEnv(grid, T, R, P, 42 + int(i)) for i in pick]
    for e in spot:
        e.reset()
    consts = np.array(orc.MAPPO_SHAPING, np.float32)
    gen = torch.Generator(device="cuda").manual_seed(5)
    grid_t = torch.from_numpy(grid.astype(np.float32)).cuda()
    for s in range(60):
        a = torch.randint(0, 15, (N, R), dtype=torch.uint8, device="cuda", generator=gen)
        r, sh, d, o, v, g, gv = full.rollout_step(a)
        r1, sh1, d1, o1, v1, g1, gv1 = lo.rollout_step(a[: N // 2])
        r2, sh2, d2, o2, v2, g2, gv2 = hi.rollout_step(a[N // 2:])
        # shard independence == determinism across instances
        for x, x1, x2 in ((r, r1, r2), (sh, sh1, sh2), (d, d1, d2), (o, o1, o2), (v, v1, v2), (g, g1, g2), (gv, gv1, gv2)):
            assert torch.equal(x[: N // 2], x1) and torch.equal(x[N // 2:], x2), s
        if s % 10 == 0 or s == 59:
            # size-independent properties
            assert bool(((o == 0) | (o == 1)).all()) and bool(((g == 0) | (g == 1)).all())
            assert torch.equal(o[:, :, 0], grid_t.expand(N, R, 10, 10)) and torch.equal(g[:, 0], grid_t.expand(N, 10, 10))
            assert bool((o[:, :, 1].sum((2, 3)) == 1).all()) and bool((o[:, :, 2].sum((2, 3)) == R - 1).all())
            assert torch.equal(o[:, :, 1] + o[:, :, 2], g[:, 1:2].expand(N, R, 10, 10))  # self + others == all robots
            assert torch.equal(o[:, :, 3], g[:, 2:3].expand(N, R, 10, 10)) and torch.equal(o[:, :, 4], g[:, 3:4].expand(N, R, 10, 10))
            assert bool((o[:, :, 5].sum((2, 3)) <= 1).all()) and bool((o[:, :, 5] <= o[:, :, 4]).all())
            tt = torch.full((N,), np.float32((s + 1) / T), device="cuda")
            assert torch.equal(v[:, :, -1], tt[:, None].expand(N, R)) and torch.equal(gv[:, -1], tt)
            assert torch.equal(v[:, :, 0], gv[:, 0:30].view(N, 5, 6)[:, :, 0]) and torch.equal(v[:, :, 2], gv[:, 0:30].view(N, 5, 6)[:, :, 2])
            assert bool((gv[:, 30:600] == 0).all())  # robot rows beyond R are zero padding
        an = a[torch.from_numpy(pick).cuda()].cpu().numpy()
        rs, shs = r[pick].cpu().numpy(), sh[pick].cpu().numpy()
        chk = s % 20 == 0 or s == 59
        if chk:
            os_, vs_, gs_, gvs_ = o[pick].cpu().numpy(), v[pick].cpu().numpy(), g[pick].cpu().numpy(), gv[pick].cpu().numpy()
        for k, e in enumerate(spot):
            rr, dd = e.step(an[k])
            assert rs[k] == rr and shs[k] == e.shaped(an[k], consts, rr), (s, k)
            if chk:
                wo, wv, wg, wgv = e.features(4, 5, 100, 100)
                assert np.array_equal(os_[k], wo) and np.array_equal(_bits(vs_[k]), _bits(wv)), (s, k)
                assert np.array_equal(gs_[k], wg) and np.array_equal(_bits(gvs_[k]), _bits(wgv)), (s, k)
    tot = full.batch.episode_stats() - lo.batch.episode_stats() - hi.batch.episode_stats()
    assert bool((tot == 0).all())  # the int64 counters all-reduce exactly


def test_step_only_large_batch_packs_four_envs_per_warp(monkeypatch):
    """Step-only launches of >= 32 768 environments with <= 8 robots run four environments per warp (8-lane groups); the
    results must not depend on that choice: a twin batch forced to one warp per environment (MDX_LPE=32) gives the same
    rewards, shaped rewards, dones, statistics and final state, bit for bit, over whole episodes with auto-reset."""
    N, R, P, T = 32768, 5, 50, 25
    kw = dict(map_file="map1.txt", n_robots=R, n_packages=P, max_time_steps=T, seed=4242, k_others=R - 1, k_packages=5)
    packed = mdb.VectorizedEnv(mdb.Environment, N, **kw).batch
    monkeypatch.setenv("MDX_LPE", "32")
    whole = mdb.VectorizedEnv(mdb.Environment, N, **kw).batch
    monkeypatch.delenv("MDX_LPE")
    for b in (packed, whole):
        b.reset()
    gen = torch.Generator(device="cuda").manual_seed(99)
    for s in range(2 * T + 5):
        a = torch.randint(0, 15, (N, R), dtype=torch.uint8, device="cuda", generator=gen)
        if s % 4 == 0:
            a = a % 5 + 5
        r1, s1, d1 = packed.step(a, auto_reset=True)
        r2, s2, d2 = whole.step(a, auto_reset=True)
        assert torch.equal(r1, r2) and torch.equal(s1.view(torch.int32), s2.view(torch.int32)) and torch.equal(d1, d2), s
    assert torch.equal(packed.episode_stats(), whole.episode_stats())
    assert int(packed.episode_stats()[0]) == 2 * N and int(packed.episode_stats()[7]) == 0
    assert torch.equal(packed.state_dict(), whole.state_dict())
