"""Round-2 GPU tests: packed state rows and the host-facing step (what the dict drop-ins ride on), the multi-step /
CUDA-graph entry points, guard regions around every caller-owned output, the process-wide shared-memory attribute, and
the checkpoint / state-injection validation."""
import os

import numpy as np
import pytest
import torch

import marl_delivery_b200 as mdb
from oracle import oracle_env as orc

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


# ---------------------------------------------------------------------------------------------- packed rows / host step
def test_packed_rows_round_trip_and_match_single_env_views():
    grid = mdb.maps.stock_map("map3.txt")
    N, R, P, T = 23, 7, 37, 60
    b = mdb.DeliveryBatch([grid], N, R, P, T, base_seed=11, k_others=4, k_packages=5)
    b.reset()
    rs = np.random.RandomState(1)
    for s in range(17):
        b.step(torch.from_numpy(rs.randint(0, 15, size=(N, R)).astype(np.uint8)), want_shaped=False)
    rows = b.get_rows()
    assert rows.shape == (N,) and rows.dtype.itemsize == b._row_bytes
    states = b.get_states()
    for i in (0, 5, 22):
        one = b.get_state(i)
        for k, v in one.items():
            assert np.array_equal(v, states[i][k]), (i, k)
        assert one["t"] == 17 and one["episode"] == 1
    sub = b.get_rows([22, 3, 3, 0])
    for f in rows.dtype.names:  # (numpy does not copy the padding between fields on fancy indexing: compare field by field)
        assert np.array_equal(sub[f], rows[f][[22, 3, 3, 0]]), f
    # write the rows of env 3 into env 9 (same map): both continue identically
    L, C = b._L, __import__("ctypes")
    ids = np.array([9], np.int32)
    src = np.ascontiguousarray(rows[3:4])
    mdb._lib.check(L.mdx_set_states(b._h, ids.ctypes.data_as(C.c_void_p), 1, src.ctypes.data_as(C.c_void_p), b._stream()))
    b._version += 1
    a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
    a[9] = a[3]
    r, sh, d = b.step(torch.from_numpy(a))
    r, sh = r.cpu().numpy(), sh.cpu().numpy()
    assert r[9] == r[3] and _bits(sh)[9] == _bits(sh)[3]
    after = b.get_rows([3, 9])
    for f in ("t", "total_reward", "robots", "sched", "status"):
        assert np.array_equal(after[f][0], after[f][1]), f


def test_step_host_is_the_step_plus_a_snapshot():
    """mdx_step_host (one C call) == step() + get_rows(), for all envs and for index subsets, against the oracle."""
    grid = mdb.maps.stock_map("map1.txt")
    N, R, P, T = 9, 5, 30, 25
    ov = orc.OracleVecEnv(N, grid, seed=70, max_time_steps=T, n_robots=R, n_packages=P)
    ov.reset()
    b = mdb.DeliveryBatch([grid], N, R, P, T, base_seed=70, k_others=4, k_packages=5)
    b.reset()
    rs = np.random.RandomState(2)
    for s in range(2 * T + 4):
        ids = None if s % 3 else sorted(set(int(x) for x in rs.randint(0, N, size=4)))
        idx = list(range(N)) if ids is None else ids
        a = rs.randint(0, 15, size=(len(idx), R)).astype(np.uint8)
        reward, shaped, done, rows = b.step_host(a, ids, want_shaped=True)
        for k, i in enumerate(idx):
            e = ov.envs[i]
            before = e.snapshot()
            rr, dd = e.step(a[k])
            assert reward[k] == rr and bool(done[k]) == dd, (s, i)
            assert _bits(shaped)[k] == _bits(np.float32(orc.shaped_reward(e, rr, before, e.snapshot(), a[k]))), (s, i)
            assert int(rows["t"][k]) == e.t and rows["total_reward"][k] == e.total_reward
            assert (rows["robots"][k] & 0xffff).tolist() == e.rcell and (rows["robots"][k] >> 16).tolist() == e.carry
            assert rows["status"][k][:P].tolist() == e.p_status
            if dd:
                e.reset()
                b.reset([i])
    b.close()


def test_dict_drop_in_matches_reference_shape_through_one_round_trip_calls():
    """VectorizedEnv.reset / .step (lists of dicts) and Environment.step ride on the packed-row calls."""
    ve = mdb.VectorizedEnv(mdb.Environment, 4, map_file="map1.txt", n_robots=5, n_packages=20, max_time_steps=30, seed=5)
    grid = mdb.maps.stock_map("map1.txt")
    ov = orc.OracleVecEnv(4, grid, seed=5, max_time_steps=30, n_robots=5, n_packages=20)
    states = ve.reset()
    ov.reset()
    rs = np.random.RandomState(4)
    for s in range(40):
        for i, st in enumerate(states):
            want = ov.envs[i].state_dict()
            assert st["time_step"] == want["time_step"] and st["robots"] == want["robots"] and st["packages"] == want["packages"]
        a = rs.randint(0, 15, size=(4, 5))
        acts = [[("DLRSU"[x % 5], str(x // 5)) for x in row] for row in a]
        states, rewards, dones, infos = ve.step(acts)
        for i in range(4):
            rr, dd = ov.envs[i].step(a[i])
            assert rewards[i] == rr and dones[i] == dd
            if dd:
                assert infos[i]["total_reward"] == ov.envs[i].total_reward and infos[i]["total_time_steps"] == ov.envs[i].t
                states[i] = ve.envs[i].reset()
                ov.envs[i].reset()
    e = mdb.Environment("map1.txt", 30, 5, 20, seed=5)
    st = e.reset()
    assert st["robots"] == ov.envs[0].__class__(grid, 30, 5, 20, seed=5).reset().state_dict()["robots"]


# ---------------------------------------------------------------------------------------------- multi-step entry points
def test_step_n_and_rollout_graph_equal_step_by_step():
    kw = dict(map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=20, seed=33, k_others=4, k_packages=5)
    N, S = 300, 48  # 20-step episodes: every env auto-resets twice inside the 48 steps
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = torch.randint(0, 15, (S, N, 5), dtype=torch.uint8, device="cuda", generator=gen)
    ref = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    ref.batch.reset()
    want = [tuple(t.clone() for t in ref.step_tensor(acts[s], auto_reset=True)) for s in range(S)]
    a = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    a.batch.reset()
    r, sh, d = a.batch.step_n(acts)
    g = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    g.batch.reset()
    graph = g.rollout_graph(S)
    gr, gsh, gd = graph.run(acts)
    torch.cuda.synchronize()
    for s in range(S):
        for got in ((r[s], sh[s], d[s]), (gr[s], gsh[s], gd[s])):
            assert torch.equal(got[0], want[s][0]) and torch.equal(got[1], want[s][1]) and torch.equal(got[2], want[s][2]), s
    # a second replay continues the rollout (the graph advances the same env state)
    want2 = [tuple(t.clone() for t in ref.step_tensor(acts[s], auto_reset=True)) for s in range(S)]
    gr, gsh, gd = graph.run()
    torch.cuda.synchronize()
    for s in range(S):
        assert torch.equal(gr[s], want2[s][0]) and torch.equal(gd[s], want2[s][2]), s
    assert torch.equal(ref.batch.episode_stats(), g.batch.episode_stats())
    with pytest.raises(ValueError):
        a.batch.step_n(acts[:, :7])
    # the pre-bound stepper: same results as step(), fused and step-only
    c1 = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    c2 = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    c1.batch.reset()
    c2.batch.reset()
    out = (torch.empty(N, dtype=torch.float64, device="cuda"), torch.empty(N, dtype=torch.float32, device="cuda"), torch.empty(N, dtype=torch.uint8, device="cuda"))
    feat = c1.batch.alloc_features(N)
    st1, st2 = c1.batch.stepper(out, feat_out=feat), c2.batch.stepper(out)
    for s in range(S):
        st1(acts[s])
        assert all(torch.equal(x, y) for x, y in zip(out, want[s]))
        st2(acts[s])
        assert all(torch.equal(x, y) for x, y in zip(out, want[s]))
    with pytest.raises(ValueError):
        st1(acts[0].to(torch.int64))


@pytest.mark.parametrize("extra", [{}, {"stale_tracker_compat": True}, {"specialize": True, "n_robots": 6},
                                   {"stale_tracker_compat": True, "map_file": ["map1.txt", "map5.txt"], "n_robots": 9, "n_packages": 70}])
def test_split_rollout_on_two_streams_equals_whole_batch_steps(extra):
    """Also with the trainer's never-cleared tracker (its per-env table is read and written by the range's own lane groups only) and
    with the kernel compiled at run time for the shape."""
    kw = dict(map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=25, seed=77, k_others=4, k_packages=5)
    kw.update(extra)
    N, S, R = 600, 60, kw["n_robots"]
    gen = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.randint(0, 15, (S, N, R), dtype=torch.uint8, device="cuda", generator=gen)
    for parts in (2, 3):
        ref = mdb.VectorizedEnv(mdb.Environment, N, **{**kw, "specialize": False})
        ref.batch.reset()
        ve = mdb.VectorizedEnv(mdb.Environment, N, **kw)
        ve.batch.reset()
        sp = ve.split(parts)
        assert [hi - lo for lo, hi in sp.ranges] == [N // parts] * parts
        for s in range(S):
            want = [t.clone() for t in ref.rollout_step(acts[s])]
            sp.fork()
            if s % 2:
                sp.step(acts[s])
            else:  # ranges may also be stepped one at a time, in any order
                for k in reversed(range(parts)):
                    sp.step_part(k, acts[s])
            sp.join()
            got = list(sp.out) + list(sp.feat)
            for x, y in zip(want, got):
                assert torch.equal(x, y), (parts, s)
        assert torch.equal(ref.batch.episode_stats(), ve.batch.episode_stats())
    # many steps in flight without joining: ranges never wait for each other, the results are the same
    ve = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    ve.batch.reset()
    ref = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    ref.batch.reset()
    sp = ve.split(2, featurize=False)
    sp.fork()
    for s in range(S):
        sp.step(acts[s])
        ref.step_tensor(acts[s], auto_reset=True)
    sp.join()
    torch.cuda.synchronize()
    assert ve.batch.get_rows().tobytes() == ref.batch.get_rows().tobytes()
    with pytest.raises(ValueError):
        ve.split(2).step_part(0, acts[0][:10])


def test_zero_copy_pinned_actions_and_outputs():
    """Pinned host tensors may stand in for the action tensor and the small outputs: the kernel reads / writes them over PCIe itself
    (no copy engine); results equal the device-tensor path."""
    kw = dict(map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=15, seed=23, k_others=4, k_packages=5)
    N, S = 200, 35
    h_act = torch.randint(0, 15, (S, N, 5), dtype=torch.uint8, generator=torch.Generator().manual_seed(4)).pin_memory()
    ref = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    ref.batch.reset()
    ve = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    ve.batch.reset()
    h_out = (torch.empty(N, dtype=torch.float64).pin_memory(), torch.empty(N, dtype=torch.float32).pin_memory(),
             torch.empty(N, dtype=torch.uint8).pin_memory())
    feat = ve.batch.alloc_features(N)
    for s in range(S):
        want = [t.clone() for t in ref.rollout_step(h_act[s].cuda())]
        ve.rollout_step(h_act[s], out=h_out, feat_out=feat)
        torch.cuda.current_stream().synchronize()
        for x, y in zip(want[:3], h_out):
            assert torch.equal(x.cpu(), y), s
        for x, y in zip(want[3:], feat):
            assert torch.equal(x, y), s
    with pytest.raises(ValueError):
        ve.rollout_step(h_act[0], out=(torch.empty(N, dtype=torch.float64), h_out[1], h_out[2]), feat_out=feat)  # pageable host memory


def test_split_host_stepper_equals_host_pipeline_results():
    kw = dict(map_file="map1.txt", n_robots=5, n_packages=50, max_time_steps=15, seed=19, k_others=4, k_packages=5)
    N, S = 256, 40
    h_act = torch.randint(0, 15, (S, N, 5), dtype=torch.uint8, generator=torch.Generator().manual_seed(2)).pin_memory()
    ref = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    ref.batch.reset()
    ve = mdb.VectorizedEnv(mdb.Environment, N, **kw)
    ve.batch.reset()
    hs = ve.split_host(2)
    for s in range(S):
        r, sh, d, o, v, g, gv = ref.rollout_step(h_act[s].cuda())
        for k in range(hs.parts):
            hs.submit(k, h_act[s])
        got_sh = np.concatenate([hs.result(k)[0] for k in range(hs.parts)])
        got_dn = np.concatenate([hs.result(k)[1] for k in range(hs.parts)])
        assert np.array_equal(_bits(got_sh), _bits(sh.cpu().numpy())) and np.array_equal(got_dn, d.cpu().numpy()), s
        hs.join()
        torch.cuda.synchronize()
        for x, y in zip((o, v, g, gv), hs.feat):
            assert torch.equal(x, y), s


# ---------------------------------------------------------------------------------------------- run-time shape specialisation
@pytest.mark.parametrize("name,R,P,T,K_o,K_p,MR,MP", [
    ("map1.txt", 5, 50, 30, 100, 100, 100, 100),   # the K = 100 defaults of generate_vector_features
    ("map5.txt", 10, 100, 40, 9, 5, 100, 100),     # C4-like: four-slice planes, long package list
    ("map.txt", 3, 8, 12, 2, 3, 4, 6),             # 7x7: generic (H*W % 4 != 0) path
    ("map5.txt", 32, 70, 20, 31, 5, 32, 64),       # full warp of robots
    ("map3.txt", 4, 20, 25, 3, 4, 10, 20),         # QMIX sizes: qualifies for the interleaved selection (R * slots <= 32, P <= 64)
    ("map1.txt", 5, 50, 30, 4, 5, 100, 100),       # the C3 shape itself (normally served by the baked variant: forced here), > 32 waiting
    ("map1.txt", 6, 64, 30, 5, 5, 6, 64),          # 6 x 5 = 30 lanes of (robot, slot) pairs, a full 64-entry list
])
def test_runtime_specialised_kernel_is_bit_identical(name, R, P, T, K_o, K_p, MR, MP, tmp_path, monkeypatch):
    """MDX_CFG_JIT_SHAPE: the fused kernel compiled by NVRTC for the instance's shape must reproduce the precompiled run-time-shape
    kernel bit for bit, through auto-resets, and the second instance of the same shape must come from the disk cache."""
    monkeypatch.setenv("MDX_JIT_CACHE", str(tmp_path))
    monkeypatch.setenv("MDX_JIT_FORCE", "1")
    grid = mdb.maps.stock_map(name)
    N = 9
    kw = dict(base_seed=31, k_others=K_o, k_packages=K_p, max_robots_in_state=MR, max_packages_in_state=MP)
    a_ = mdb.DeliveryBatch([grid], N, R, P, T, specialize=False, **kw)
    b_ = mdb.DeliveryBatch([grid], N, R, P, T, specialize=True, **kw)
    st, msg = b_.jit_status()
    assert st == 1, msg
    assert a_.jit_status()[0] == 0 or os.environ.get("MDX_JIT")  # (MDX_JIT=1 forces the run-time-compiled kernel onto every instance)
    assert len(list(tmp_path.glob("*.cubin"))) == 1
    c_ = mdb.DeliveryBatch([grid], N, R, P, T, specialize=True, **kw)   # served from the cache
    assert c_.jit_status()[0] == 1 and len(list(tmp_path.glob("*.cubin"))) == 1
    for x in (a_, b_, c_):
        x.reset()
    rs = np.random.RandomState(6)
    for s in range(2 * T + 3):
        a = torch.from_numpy(rs.randint(0, 15, size=(N, R)).astype(np.uint8))
        want = [t.clone() for t in a_.step_featurize(a, auto_reset=True)]
        for x in (b_, c_) if s % 5 == 0 else (b_,):
            got = x.step_featurize(a, auto_reset=True)
            for u, v in zip(want, got):
                assert torch.equal(u, v), s
        if s % 5:
            c_.step_featurize(a, auto_reset=True)
        if s % 7 == 0:  # the featurise-only kernel of the shape is compiled along with the fused one
            for u, v in zip(a_.featurize(), b_.featurize()):
                assert torch.equal(u, v), s
            sub = [N - 1, 0, 3]
            for u, v in zip(a_.featurize(ids=sub), b_.featurize(ids=sub)):
                assert torch.equal(u.clone(), v), s
    assert torch.equal(a_.episode_stats(), b_.episode_stats())


@pytest.mark.parametrize("chunk", ["8", "24"])
@pytest.mark.parametrize("name,R,P,T,K_o,K_p,MR,MP", [
    ("map1.txt", 5, 60, 40, 4, 5, 100, 100),     # more active packages than a chunk, zero tail behind the last chunk
    ("map5.txt", 20, 90, 30, 19, 5, 20, 37),     # MP not a multiple of the chunk; bitonic other-robot ranking (> 16 robots)
    ("map.txt", 3, 30, 25, 2, 3, 4, 30),         # 7x7 generic path
])
def test_gvec_package_rows_flushed_in_chunks(chunk, name, R, P, T, K_o, K_p, MR, MP, monkeypatch):
    """A smaller gvec stage (MDX_GVEC_CHUNK rows per flush; chosen automatically where it buys a resident CTA, e.g. C5) must not
    change a bit of gvec -- or of anything else."""
    grid = mdb.maps.stock_map(name)
    N = 7
    kw = dict(base_seed=5, k_others=K_o, k_packages=K_p, max_robots_in_state=MR, max_packages_in_state=MP, specialize=False)
    a_ = mdb.DeliveryBatch([grid], N, R, P, T, **kw)
    monkeypatch.setenv("MDX_GVEC_CHUNK", chunk)
    b_ = mdb.DeliveryBatch([grid], N, R, P, T, **kw)
    c_ = mdb.DeliveryBatch([grid], N, R, P, T, **{**kw, "specialize": True})
    monkeypatch.delenv("MDX_GVEC_CHUNK")
    assert c_.jit_status()[0] == 1
    for x in (a_, b_, c_):
        x.reset()
    rs = np.random.RandomState(7)
    for s in range(2 * T + 2):
        a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
        if s % 2:
            a = (a % 5 + 5).astype(np.uint8)  # many pickups: in-transit rows with carrier indices
        a = torch.from_numpy(a)
        want = [t.clone() for t in a_.step_featurize(a, auto_reset=True)]
        for x in (b_, c_):
            for u, v in zip(want, x.step_featurize(a, auto_reset=True)):
                assert torch.equal(u, v), s


# ---------------------------------------------------------------------------------------------- memory safety
CANARY = 1024  # elements on each side


def _guarded(shape, dtype):
    n = int(np.prod(shape))
    # 16-byte alignment of the payload is part of the contract: keep the guard a multiple of 16 bytes
    raw = torch.full((n + 2 * CANARY,), 0x5A if dtype == torch.uint8 else 0, dtype=dtype, device="cuda")
    if dtype != torch.uint8:
        raw.view(torch.uint8).fill_(0x5A)
    return raw, raw[CANARY:CANARY + n].view(shape)


def _intact(raw, n):
    b = raw.view(torch.uint8)
    es = raw.element_size()
    return bool((b[:CANARY * es] == 0x5A).all()) and bool((b[(CANARY + n) * es:] == 0x5A).all())


SHAPES = [
    ("map1.txt", 5, 50, 30, 4, 5, 100, 100),     # the headline shape (specialised variant)
    ("map.txt", 3, 8, 12, 2, 3, 4, 6),           # 7x7: generic (H*W % 4 != 0) path
    ("map5.txt", 32, 64, 20, 31, 5, 32, 64),     # full warp
    ("map2.txt", 2, 120, 40, 1, 3, 3, 10),       # long list, register top-k path
    ("map3.txt", 4, 100, 30, 100, 100, 100, 100),  # K = 100 defaults
    ("map1.txt", 1, 0, 5, 0, 3, 2, 2),           # no packages
]


@pytest.mark.parametrize("lpe", ["", "8", "16"])
@pytest.mark.parametrize("name,R,P,T,K_o,K_p,MR,MP", SHAPES)
def test_guard_regions_around_caller_owned_outputs(lpe, name, R, P, T, K_o, K_p, MR, MP):
    """Every caller-owned output sits between two 0x5A canaries; after fused, step-only, featurise-only, IDQ and greedy
    launches (odd env counts, index subsets) the canaries must be untouched.  compute-sanitizer cannot run on this pool;
    this is the evidence that the kernels write exactly the bytes the contract names."""
    if lpe and R > int(lpe):
        pytest.skip("lane group smaller than the robot count")
    old = os.environ.get("MDX_LPE")
    if lpe:
        os.environ["MDX_LPE"] = lpe
    try:
        grid = mdb.maps.stock_map(name)
        N = 13
        b = mdb.DeliveryBatch([grid], N, R, P, T, base_seed=3, k_others=K_o, k_packages=K_p, max_robots_in_state=MR,
                              max_packages_in_state=MP)
    finally:
        if old is None:
            os.environ.pop("MDX_LPE", None)
        else:
            os.environ["MDX_LPE"] = old
    b.reset()
    H, W, D, Dg = b.rows_out, b.cols_out, b.vec_dim, b.gvec_dim
    rs = np.random.RandomState(0)
    for n, ids in ((N, None), (5, [12, 0, 7, 3, 4])):
        bufs = {"reward": _guarded((n,), torch.float64), "shaped": _guarded((n,), torch.float32), "done": _guarded((n,), torch.uint8),
                "obs": _guarded((n, R, 6, H, W), torch.float32), "vec": _guarded((n, R, D), torch.float32),
                "gstate": _guarded((n, 4, H, W), torch.float32), "gvec": _guarded((n, Dg), torch.float32),
                "agent": _guarded((n, R), torch.float64), "obs6": _guarded((n, R, 6, H, W), torch.float32),
                "g7": _guarded((n, 7, H, W), torch.float32), "acts": _guarded((n, R), torch.uint8)}
        out = (bufs["reward"][1], bufs["shaped"][1], bufs["done"][1])
        feat = (bufs["obs"][1], bufs["vec"][1], bufs["gstate"][1], bufs["gvec"][1])
        for s in range(T + 3):
            a = torch.from_numpy(rs.randint(0, 15, size=(n, R)).astype(np.uint8)).cuda()
            k = s % 5
            if k == 0:
                b.step_featurize(a, ids=ids, auto_reset=True, out=out, feat_out=feat)
            elif k == 1:
                b.step(a, ids=ids, auto_reset=True, out=out)
            elif k == 2:
                b.featurize(ids=ids, out=feat)
            elif k == 3:
                b.step_idq_featurize(a, ids=ids, auto_reset=True, ops_as_int=True, out=(out[0], bufs["agent"][1], out[2]),
                                     feat_out=(bufs["obs6"][1], bufs["g7"][1]))
            else:
                b.step(b.greedy_actions(ids=ids, out=bufs["acts"][1]), ids=ids, auto_reset=True, out=out)
        torch.cuda.synchronize()
        for nm, (raw, view) in bufs.items():
            assert _intact(raw, view.numel()), (nm, n)
    b.close()


# ---------------------------------------------------------------------------------------------- advisor findings
def test_two_batches_sharing_a_kernel_variant_with_different_shared_memory():
    """cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel function, not to an instance: a small batch
    created after a big one (same template variant) must not lower the limit under the big one's feet."""
    g5 = mdb.maps.stock_map("map5.txt")
    big = mdb.DeliveryBatch([g5], 64, 6, 500, 50, base_seed=1, k_others=100, k_packages=100)   # ~58 KB per CTA
    small = mdb.DeliveryBatch([g5], 64, 6, 100, 50, base_seed=1, k_others=5, k_packages=5)
    big.reset()
    small.reset()
    a = torch.zeros((64, 6), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        big.step_featurize(a)
        small.step_featurize(a)
        big.step_featurize(a)
    torch.cuda.synchronize()
    assert int(big.episode_stats().cpu()[5]) == 64 * 6 and int(small.episode_stats().cpu()[5]) == 64 * 3


def test_set_state_refuses_what_the_reference_constructor_refuses():
    b = mdb.DeliveryBatch([mdb.maps.stock_map("map1.txt")], 2, 2, 2, 50, base_seed=0, blank=True)
    grid = mdb.maps.stock_map("map1.txt")
    free = np.argwhere(grid == 0)
    wall = np.argwhere(grid == 1)
    (r0, c0), (r1, c1) = free[0], free[1]
    pk = [(r0, c0, r1, c1, 0, 30, 1), (r1, c1, r0, c0, 5, 40, 0)]
    b.set_state(0, 3, 0.0, [(r0, c0, 0), (r1, c1, 0)], pk)
    with pytest.raises(ValueError):
        b.set_state(0, -1, 0.0, [(r0, c0, 0), (r1, c1, 0)], pk)          # negative time
    with pytest.raises(ValueError):
        b.set_state(0, 70000, 0.0, [(r0, c0, 0), (r1, c1, 0)], pk)       # beyond the u16 deadline arithmetic
    with pytest.raises(ValueError):
        b.set_state(0, 3, 0.0, [(r0, c0, 0), (r0, c0, 0)], pk)           # two robots on one cell (env.py:73-79)
    if len(wall):
        with pytest.raises(ValueError):
            b.set_state(0, 3, 0.0, [(int(wall[0][0]), int(wall[0][1]), 0), (r1, c1, 0)], pk)  # robot on an obstacle
    with pytest.raises(ValueError):
        b.set_state(0, 3, 0.0, [(r0, c0, 3), (r1, c1, 0)], pk)           # unknown carried id
    assert b.get_state(0)["t"] == 3  # refused calls changed nothing


def test_env_ids_are_validated():
    b = mdb.DeliveryBatch([mdb.maps.stock_map("map1.txt")], 8, 5, 20, 50, base_seed=3, k_others=4, k_packages=5)
    a = torch.zeros((2, 5), dtype=torch.uint8, device="cuda")
    with pytest.raises((IndexError, ValueError)):
        b.step(a, ids=[1, 8])
    with pytest.raises((IndexError, ValueError)):
        b.step(a, ids=torch.tensor([1, -1], device="cuda"))
    with pytest.raises((IndexError, ValueError)):
        b.step(a, ids=torch.tensor([2, 2], device="cuda"))  # the same env twice in one launch would race
    with pytest.raises((IndexError, ValueError)):
        b.step(a, ids=[2, 2])
    b.step(a, ids=torch.tensor([7, 0], device="cuda"))


def test_checkpoint_refuses_foreign_or_corrupt_blobs_and_stays_untouched():
    g1, g3 = mdb.maps.stock_map("map1.txt"), mdb.maps.stock_map("map3.txt")
    mk = lambda grid, **kw: mdb.DeliveryBatch([grid], 16, 5, 20, 40, base_seed=9, k_others=4, k_packages=5, **kw)  # noqa: E731
    b = mk(g1)
    b.reset()
    a = torch.ones((16, 5), dtype=torch.uint8, device="cuda")
    for _ in range(7):
        b.step(a)
    blob = b.state_dict()
    before = b.get_rows().tobytes()
    other_map = mdb.DeliveryBatch([np.zeros((10, 10), np.uint8)], 16, 5, 20, 40, base_seed=9, k_others=4, k_packages=5)
    with pytest.raises(ValueError):
        other_map.load_state_dict(blob)  # same shape, other map: cell ids would land anywhere
    with pytest.raises(ValueError):
        mk(g1, shaping=mdb.QMIX_SHAPING).load_state_dict(blob)
    with pytest.raises(ValueError):
        mdb.DeliveryBatch([g3], 16, 5, 20, 40, base_seed=9, k_others=4, k_packages=5).load_state_dict(blob)  # other plane size
    bad = blob.clone()
    # corrupt one generator position (third device segment: t, ep, mtpos) -> refused before anything is written
    n, R, Pp8, Pp16 = 16, 5, 24, 32
    segs = 4 * n + 4 * n + 4 * n + 8 * n + 4 * n * R + n * Pp16 + 2 * n * 4 * Pp8 + 32 * n + 2496 * n  # t, ep, mtpos, total, robots, status, cur, counters, mt
    hdr = bad.numel() - segs
    assert 64 <= hdr <= 96
    off = hdr + 2 * 4 * n + 4 * 5  # mtpos of env 5
    bad[off:off + 4] = torch.tensor([0xff, 0xff, 0xff, 0x7f], dtype=torch.uint8)
    for _ in range(3):
        b.step(a)
    mid = b.get_rows().tobytes()
    with pytest.raises(ValueError):
        b.load_state_dict(bad)
    assert b.get_rows().tobytes() == mid  # the refused load left the instance as it was
    with pytest.raises(ValueError):
        b.load_state_dict(blob[:-5])
    b.load_state_dict(blob)
    assert b.get_rows().tobytes() == before
