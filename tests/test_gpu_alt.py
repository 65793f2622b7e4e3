"""GPU parity tests of the f3 row: IDQ / legacy-qmix featurisers, IDQ reward_shaping
and the QMIX-sized MAPPO featurisers, against values recorded from the unmodified reference
(tests/golden/kats_alt.npz) and, on fresh seeds, against the pinned oracle.  All through the C ABI."""
import json
import os

import numpy as np
import pytest
import torch

import marl_delivery_b200 as mdb
from golden_util import GOLDEN
from oracle import alt_oracle as alt
from oracle import oracle_env as orc

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def _cases():
    z = np.load(os.path.join(GOLDEN, "kats_alt.npz"))
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4, 5])  # 4, 5: non-square maps the reference does not ship (10x14 sliced, 9x7 generic)
def test_alt_golden_replay(idx):
    z, metas = _cases()
    m = metas[idx]
    n, R = m["name"], m["R"]
    mk = lambda: mdb.DeliveryBatch([z[f"{n}_grid"]], 1, R, m["P"], m["T"], seeds=[m["seed"]], k_others=R - 1, k_packages=5,  # noqa: E731
                                   max_robots_in_state=10, max_packages_in_state=20, shaping=mdb.QMIX_SHAPING)
    b_int, b_str = mk(), mk()
    for b in (b_int, b_str):
        b.reset()

    def check(k):
        o6, g7 = [t.cpu().numpy() for t in b_int.featurize_alt()]
        assert np.array_equal(_bits(o6[0]), _bits(z[f"{n}_obs6"][k])), (n, k)
        assert np.array_equal(_bits(g7[0]), _bits(z[f"{n}_g7"][k])), (n, k)
        _, v, _, gv = [t.cpu().numpy() for t in b_int.featurize()]
        assert np.array_equal(_bits(v[0]), _bits(z[f"{n}_qvec"][k])), (n, k)
        assert np.array_equal(_bits(gv[0]), _bits(z[f"{n}_qgvec"][k])), (n, k)

    check(0)
    for s in range(m["n_steps"]):
        a = torch.from_numpy(z[f"{n}_actions"][s][None].copy())
        _, ag_int, d = b_int.step_idq(a, ops_as_int=True)
        _, ag_str, d2 = b_str.step_idq(a, ops_as_int=False)
        assert np.array_equal(ag_int.cpu().numpy()[0], z[f"{n}_shape_int"][s]), (n, s)
        assert np.array_equal(ag_str.cpu().numpy()[0], z[f"{n}_shape_str"][s]), (n, s)
        assert int(d.cpu()[0]) == z[f"{n}_done"][s] == int(d2.cpu()[0])
        if int(d.cpu()[0]):
            b_int.reset()
            b_str.reset()
        check(s + 1)


@pytest.mark.parametrize("maps,R,P,T", [(["map1.txt"], 5, 50, 40), (["map.txt"], 6, 20, 30), (["map5.txt"], 32, 200, 45),
                                        (["map1.txt", "map3.txt", "map5.txt"], 9, 80, 35)])
def test_alt_random_sweep_vs_oracle(maps, R, P, T):
    """Fresh seeds, random + pickup-heavy actions, on-device auto-reset, heterogeneous (padded) maps."""
    grids = [mdb.maps.stock_map(x) for x in maps]
    N = 12
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file=maps if len(maps) > 1 else maps[0], n_robots=R, n_packages=P,
                           max_time_steps=T, seed=501)
    b = ve.batch
    b.reset()
    envs = [orc.OracleEnv(grids[i % len(grids)], T, R, P, seed=501 + i) for i in range(N)]
    for e in envs:
        e.reset()
    rs = np.random.RandomState(6)
    for s in range(2 * T + 5):
        a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
        if s % 3 == 0:
            a = (a % 5 + 5).astype(np.uint8)  # everybody tries to pick up
        _, ag, d = b.step_idq(torch.from_numpy(a), auto_reset=True, ops_as_int=True)
        ag, d = ag.cpu().numpy(), d.cpu().numpy()
        o6, g7 = [t.cpu().numpy() for t in b.featurize_alt()]
        for n, e in enumerate(envs):
            before = e.snapshot()
            _, dd = e.step(a[n])
            assert list(ag[n]) == alt.idq_reward_shaping(e, before, e.snapshot(), a[n], True), (s, n)
            assert bool(d[n]) == dd
            if dd:
                e.reset()
            if (s + n) % 4 == 0:
                snap = e.snapshot()
                H, W = e.H, e.W
                want6 = np.stack([alt.idq_observation(e, snap, i) for i in range(R)])
                assert np.array_equal(_bits(o6[n][:, :, :H, :W]), _bits(want6)), (s, n)
                assert np.array_equal(_bits(g7[n][:, :H, :W]), _bits(alt.legacy_global_tensor(e, snap))), (s, n)
                assert np.count_nonzero(o6[n]) == np.count_nonzero(want6) and np.count_nonzero(g7[n]) == np.count_nonzero(alt.legacy_global_tensor(e, snap))  # padding stays zero


@pytest.mark.parametrize("maps,R,P,T", [(["map1.txt"], 5, 50, 40), (["map.txt"], 6, 20, 30), (["map5.txt"], 32, 200, 45),
                                        (["map1.txt", "map3.txt", "map5.txt"], 9, 80, 35), (["map5.txt"], 2, 3, 25)])
@pytest.mark.parametrize("unfused", [False, True])
def test_idq_fused_iteration_equals_two_launches(maps, R, P, T, unfused, monkeypatch):
    """mdx_step_idq_featurize (ONE launch: step + per-agent shaping + auto-reset + convert_state x R + 7-channel tensor) against
    mdx_step_idq followed by mdx_featurize_alt on a twin batch; map.txt (7x7: planes are not whole float4s) and
    MDX_IDQ_UNFUSED=1 exercise the internal two-launch fallback.  The unfused pair is the one checked against the oracle above."""
    if unfused:
        monkeypatch.setenv("MDX_IDQ_UNFUSED", "1")
    N = 40
    kw = dict(map_file=maps if len(maps) > 1 else maps[0], n_robots=R, n_packages=P, max_time_steps=T, seed=77)
    a_, b_ = mdb.VectorizedEnv(mdb.Environment, N, **kw).batch, mdb.VectorizedEnv(mdb.Environment, N, **kw).batch
    for x in (a_, b_):
        x.reset()
    rs = np.random.RandomState(9)
    ids = torch.tensor([3, 0, 17, 39, 8], dtype=torch.int32)
    for s in range(2 * T + 7):
        sub = s % 11 == 5  # now and then only a subset of the envs steps (QMIX-style indices)
        n = len(ids) if sub else N
        a = rs.randint(0, 15, size=(n, R)).astype(np.uint8)
        if s % 3 == 0:
            a = (a % 5 + 5).astype(np.uint8)
        kws = dict(ids=ids if sub else None, auto_reset=True, ops_as_int=bool(s & 1))
        r1, g1, d1, o1, t1 = a_.step_idq_featurize(torch.from_numpy(a), **kws)
        r2, g2, d2 = b_.step_idq(torch.from_numpy(a), **kws)
        o2, t2 = b_.featurize_alt(ids=ids if sub else None)
        assert torch.equal(r1, r2) and torch.equal(g1, g2) and torch.equal(d1, d2), s
        assert torch.equal(o1, o2) and torch.equal(t1, t2), s
    assert torch.equal(a_.episode_stats(), b_.episode_stats())
    assert torch.equal(a_.state_dict(), b_.state_dict())


@pytest.mark.parametrize("maps,R,P,T,spec", [(["map1.txt"], 5, 50, 40, False), (["map1.txt"], 7, 30, 25, True), (["map5.txt"], 32, 200, 45, True),
                                             (["map1.txt", "map3.txt", "map5.txt"], 9, 80, 35, True), (["map.txt"], 6, 20, 30, False)])
def test_idq_fused_iteration_by_env_ranges_and_run_time_specialised(maps, R, P, T, spec):
    """mdx_step_idq_featurize_range: three contiguous env ranges of one instance, each on its own stream and writing its rows of the
    full-batch tensors, against the whole-batch launch of a twin; with specialize=True the fused IDQ kernel is the one compiled at
    run time for the shape (mdx_jit.cpp: mode 5).  map.txt has no fused kernel: the range entry point must refuse it."""
    N = 44
    kw = dict(map_file=maps if len(maps) > 1 else maps[0], n_robots=R, n_packages=P, max_time_steps=T, seed=31)
    a_ = mdb.VectorizedEnv(mdb.Environment, N, specialize=spec, **kw).batch
    b_ = mdb.VectorizedEnv(mdb.Environment, N, specialize=False, **kw).batch
    if spec:
        assert a_.jit_status()[0] == 1
    for x in (a_, b_):
        x.reset()
    dev, H, W = a_.device, a_.rows_out, a_.cols_out
    out = (torch.empty(N, dtype=torch.float64, device=dev), torch.empty((N, R), dtype=torch.float64, device=dev), torch.empty(N, dtype=torch.uint8, device=dev))
    feat = (torch.empty((N, R, 6, H, W), dtype=torch.float32, device=dev), torch.empty((N, 7, H, W), dtype=torch.float32, device=dev))
    ranges = [(0, 12), (12, 4), (16, 28)]
    streams = [torch.cuda.Stream(dev) for _ in ranges]
    rs = np.random.RandomState(5)
    if maps == ["map.txt"]:
        act = torch.zeros((N, R), dtype=torch.uint8, device=dev)
        with pytest.raises(ValueError):
            a_.step_idq_featurize(act[:12], env_range=(0, 12), out=tuple(t[:12] for t in out), feat_out=tuple(t[:12] for t in feat))
        return
    for s in range(2 * T + 5):
        act = torch.from_numpy(rs.randint(0, 15, size=(N, R)).astype(np.uint8)).to(dev)
        torch.cuda.synchronize()
        for (lo, n), st in zip(ranges, streams):
            a_.step_idq_featurize(act[lo:lo + n], auto_reset=True, ops_as_int=bool(s & 1), out=tuple(t[lo:lo + n] for t in out),
                                  feat_out=tuple(t[lo:lo + n] for t in feat), env_range=(lo, n), stream=st)
        torch.cuda.synchronize()
        r2, g2, d2, o2, t2 = b_.step_idq_featurize(act, auto_reset=True, ops_as_int=bool(s & 1))
        assert torch.equal(out[0], r2) and torch.equal(out[1], g2) and torch.equal(out[2], d2), s
        assert torch.equal(feat[0], o2) and torch.equal(feat[1], t2), s
    assert torch.equal(a_.episode_stats(), b_.episode_stats())
    with pytest.raises(ValueError):
        a_.step_idq_featurize(act[:4], env_range=(N - 2, 4), out=tuple(t[:4] for t in out), feat_out=tuple(t[:4] for t in feat))


@pytest.mark.parametrize("wide", [False, True])
def test_split_idq_equals_the_whole_batch_call(wide):
    """VectorizedEnv.split_idq: fork / step / join over 3 env ranges (uint8 codes or the policy's int64 samples) against
    step_idq_featurize on a twin; then other full-batch tensors passed per call (step_part(out=, feat_out=))."""
    N, R, T = 36, 5, 30
    kw = dict(map_file="map1.txt", n_robots=R, n_packages=40, max_time_steps=T, seed=8)
    va, vb = mdb.VectorizedEnv(mdb.Environment, N, **kw), mdb.VectorizedEnv(mdb.Environment, N, **kw)
    va.batch.reset(), vb.batch.reset()
    dev = va.batch.device
    si = va.split_idq(3, ops_as_int=True)
    assert [hi - lo for lo, hi in si.ranges] == [12, 12, 12]
    rs = np.random.RandomState(4)
    for s in range(2 * T + 3):
        codes = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
        act = torch.from_numpy(codes).to(dev)
        a_in = act.to(torch.int64) if wide else act
        si.fork()
        si.step(a_in)
        si.join()
        r2, g2, d2, o2, t2 = vb.batch.step_idq_featurize(act, auto_reset=True, ops_as_int=True)
        torch.cuda.synchronize()
        assert torch.equal(si.out[0], r2) and torch.equal(si.out[1], g2) and torch.equal(si.out[2], d2), s
        assert torch.equal(si.feat[0], o2) and torch.equal(si.feat[1], t2), s
    H, W = va.batch.rows_out, va.batch.cols_out
    out = (torch.empty(N, dtype=torch.float64, device=dev), torch.empty((N, R), dtype=torch.float64, device=dev), torch.empty(N, dtype=torch.uint8, device=dev))
    feat = (torch.empty((N, R, 6, H, W), dtype=torch.float32, device=dev), torch.empty((N, 7, H, W), dtype=torch.float32, device=dev))
    act = torch.from_numpy(rs.randint(0, 15, size=(N, R)).astype(np.uint8)).to(dev)
    si.fork()
    for k in range(3):
        si.step_part(k, act, out=out, feat_out=feat)
    si.join()
    r2, g2, d2, o2, t2 = vb.batch.step_idq_featurize(act, auto_reset=True, ops_as_int=True)
    torch.cuda.synchronize()
    assert torch.equal(out[0], r2) and torch.equal(out[1], g2) and torch.equal(out[2], d2)
    assert torch.equal(feat[0], o2) and torch.equal(feat[1], t2)
    assert torch.equal(va.batch.episode_stats(), vb.batch.episode_stats())


def _tracker_of(env):
    return {p.package_id: {"id": p.package_id, "start_pos": p.start, "target_pos": p.target, "start_time": p.start_time,
                           "deadline": p.deadline, "status": p.status}
            for p in env.packages if p.status in ("waiting", "in_transit")}


def test_idq_and_qmix_drop_in_functions():
    """Per-call drop-ins with the reference's signatures, driven from state dicts like the trainers do."""
    z, metas = _cases()
    m = metas[0]
    n, R = m["name"], m["R"]
    env = mdb.Environment("map1.txt", max_time_steps=m["T"], n_robots=R, n_packages=m["P"], seed=m["seed"])
    state = env.reset()
    for s in range(60):
        pp = _tracker_of(env)
        if s % 6 == 0:
            for i in range(R):
                assert np.array_equal(_bits(mdb.idq.convert_state(state, pp, i)), _bits(z[f"{n}_obs6"][s, i])), (s, i)
                got = mdb.qmix_helper.generate_vector_features(state, pp, i, m["T"], R - 1, 5)
                assert np.array_equal(_bits(got), _bits(z[f"{n}_qvec"][s, i])), (s, i)
            assert np.array_equal(_bits(mdb.idq.convert_global_state_to_tensor(state, pp, (7, 10, 10))), _bits(z[f"{n}_g7"][s]))
            assert np.array_equal(_bits(mdb.qmix_helper.convert_global_state(state, pp, m["T"], 10, 20)[1]), _bits(z[f"{n}_qgvec"][s]))
        a = z[f"{n}_actions"][s]
        acts = [("DLRSU"[x % 5], str(x // 5)) for x in a]
        nxt, r, done, _ = env.step(acts)
        assert mdb.idq.reward_shaping(state, nxt, acts, pp, R) == list(z[f"{n}_shape_str"][s]), s
        assert mdb.idq.reward_shaping(state, nxt, [(mv, int(o)) for mv, o in acts], pp, R) == list(z[f"{n}_shape_int"][s]), s
        if done:
            nxt = env.reset()
        state = nxt
    with pytest.raises(TypeError):
        mdb.qmix_helper.generate_vector_features(state, pp, 0, 100)  # QMIX/helper.py:70 has no defaults
    with pytest.raises(ValueError):
        mdb.idq.convert_global_state_to_tensor(state, pp, (7, 12, 12))
