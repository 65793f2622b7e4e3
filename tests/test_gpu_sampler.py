"""The DEVICE reset sampler (csrc/mdx_sampler.cuh: per-env MT19937 state in HBM, numpy's masked-rejection randint, the
reference's draw order, stable sort by start_time) against (a) the reset blocks the unmodified reference produced
(tests/golden/reset_blocks.npz), (b) the host restatement of the same stream (mdx_sampler_*, itself pinned to numpy
and to the same fixtures on the CPU) over many resets -- every env crosses several generator twists -- and (c) the
oracle through auto-resets drawn INSIDE the step kernels, for whole-warp and packed lane groups."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import marl_delivery_b200 as mdb
from golden_util import load_reset_blocks
from marl_delivery_b200 import _lib
from oracle import oracle_env as orc

pytestmark = pytest.mark.gpu


def _host_blocks(grid, R, P, T, seed, n_blocks):
    L = _lib.lib()
    g = np.ascontiguousarray(grid, dtype=np.uint8)
    h = C.c_void_p()
    _lib.check(L.mdx_sampler_create(seed, g.ctypes.data_as(C.POINTER(C.c_uint8)), g.shape[0], g.shape[1], R, P, T, C.byref(h)))
    out = []
    for _ in range(n_blocks):
        rc = np.zeros(max(R, 1), np.int32)
        pk = np.zeros((4, max(P, 1)), np.int32)
        _lib.check(L.mdx_sampler_next(h, rc.ctypes.data_as(C.POINTER(C.c_int32)), pk.ctypes.data_as(C.POINTER(C.c_int32))))
        out.append((rc[:R].copy(), pk.reshape(-1)[:4 * P].reshape(4, P).copy()))
    L.mdx_sampler_destroy(h)
    return out


def _device_block(b, i, W):
    s = b.get_state(i)
    rc = s["robot_row"] * W + s["robot_col"]
    pk = np.stack([s["pkg_start_row"] * W + s["pkg_start_col"], s["pkg_target_row"] * W + s["pkg_target_col"],
                   s["pkg_start_time"], s["pkg_deadline"]], 0)
    return s, rc, pk


def test_device_sampler_matches_reference_reset_blocks():
    """Block 0 is the constructor's reset (env.py:41), blocks 1 and 2 the next two reset() calls."""
    meta, z = load_reset_blocks()
    for k, m in enumerate(meta):
        grid = z[f"grid_{k}"]
        W = grid.shape[1]
        b = mdb.DeliveryBatch([grid], 3, m["R"], m["P"], m["T"], seeds=[m["seed"]] * 3, k_others=4, k_packages=5)
        for blk in range(3):
            if blk:
                b.reset()
            for i in (0, 2):
                s, rc, pk = _device_block(b, i, W)
                assert s["episode"] == blk and s["t"] == 0
                assert np.array_equal(rc, z[f"robots_{k}"][blk]), (m, blk)
                assert np.array_equal(pk.T, z[f"pkgs_{k}"][blk]), (m, blk)
                assert np.array_equal(s["pkg_status"], (pk[2] == 0).astype(np.int32)), (m, blk)
                assert np.all(s["robot_carry"] == 0)
        b.close()


@pytest.mark.parametrize("name,R,P,T,n_resets", [
    ("map1.txt", 5, 50, 500, 40),    # ~205 draws per block: a twist every third reset
    ("map5.txt", 32, 500, 1000, 6),  # ~2000 draws per block: several twists inside ONE reset
    ("map.txt", 25, 3, 50, 30),      # as many robots as free cells: the shrinking list runs down to one entry (no draw)
    ("map3.txt", 1, 0, 10, 5),       # no packages
    ("map1.txt", 5, 5, 2, 20),       # T = 2: randint(1, 2) is a single value, no draw
    ("map2.txt", 7, 129, 60000, 8),  # widest start_time range
])
def test_device_sampler_matches_host_stream_over_many_resets(name, R, P, T, n_resets):
    grid = mdb.maps.stock_map(name)
    W = grid.shape[1]
    N = 6
    seeds = [7, 8, 2 ** 32 - 1, 0, 123456789, 7]
    b = mdb.DeliveryBatch([grid], N, R, P, T, seeds=seeds, k_others=4, k_packages=5)
    want = [_host_blocks(grid, R, P, T, sd, 2 * n_resets + 2) for sd in seeds]
    count = [0] * N
    rs = np.random.RandomState(5)
    for it in range(n_resets):
        if it % 3 == 2:  # subsets with a repeated id: that env is reset twice (the reference would draw two blocks)
            ids = [int(x) for x in rs.randint(0, N, size=3)] + [1, 1]
        else:
            ids = None
        b.reset(ids)
        for i in (range(N) if ids is None else ids):
            count[i] += 1
        for i in range(N):
            s, rc, pk = _device_block(b, i, W)
            hr, hp = want[i][count[i]]
            assert s["episode"] == count[i]
            assert np.array_equal(rc, hr), (it, i)
            assert np.array_equal(pk, hp), (it, i)
    b.close()


@pytest.mark.parametrize("lpe", ["", "8", "16"])
def test_auto_reset_draws_inside_the_step_kernels(lpe):
    """Short episodes, step-only and fused launches, whole-warp and packed lane groups: every auto-reset draws the env's
    next block inside the kernel and the rollout stays bit-identical to the oracle (numpy RandomState) throughout."""
    old = os.environ.get("MDX_LPE")
    if lpe:
        os.environ["MDX_LPE"] = lpe
    try:
        grid = mdb.maps.stock_map("map1.txt")
        N, R, P, T = 37, 5, 12, 9
        ov = orc.OracleVecEnv(N, grid, seed=500, max_time_steps=T, n_robots=R, n_packages=P)
        ov.reset()
        b = mdb.DeliveryBatch([grid], N, R, P, T, base_seed=500, k_others=4, k_packages=5)
        b.reset()
        rs = np.random.RandomState(9)
        for s in range(12 * T + 1):  # 12 episodes per env: ~600 draws, at least one twist each
            a = rs.randint(0, 15, size=(N, R)).astype(np.uint8)
            if s % 2:
                wsh, wd, wo, wv, wg, wgv = orc.mappo_rollout_step(ov, a, 4, 5)
                r, sh, d, o, v, g, gv = b.step_featurize(torch.from_numpy(a), auto_reset=True)
                assert np.array_equal(v.cpu().numpy().view(np.uint32), wv.view(np.uint32)), s
                assert np.array_equal(o.cpu().numpy(), wo) and np.array_equal(gv.cpu().numpy().view(np.uint32), wgv.view(np.uint32)), s
            else:
                wsh = np.zeros(N, np.float32)
                wd = np.zeros(N, bool)
                for n, e in enumerate(ov.envs):
                    before = e.snapshot()
                    rr, dd = e.step(a[n])
                    wsh[n] = orc.shaped_reward(e, rr, before, e.snapshot(), a[n])
                    wd[n] = dd
                    if dd:
                        e.reset()
                r, sh, d = b.step(torch.from_numpy(a), auto_reset=True)
            assert np.array_equal(sh.cpu().numpy().view(np.uint32), wsh.view(np.uint32)), s
            assert np.array_equal(d.cpu().numpy().astype(bool), wd), s
        for n in (0, 17, 36):
            st = b.get_state(n)
            e = ov.envs[n]
            assert (st["robot_row"] * e.W + st["robot_col"]).tolist() == e.rcell and st["pkg_status"].tolist() == e.p_status
            assert st["pkg_start_time"].tolist() == e.p_t0 and st["pkg_deadline"].tolist() == e.p_deadline
        stats = b.episode_stats().cpu().tolist()
        assert stats[0] == 12 * N and stats[7] == 0
        b.close()
    finally:
        if old is None:
            os.environ.pop("MDX_LPE", None)
        else:
            os.environ["MDX_LPE"] = old


def test_create_refuses_what_the_reference_reset_cannot_draw():
    grid = mdb.maps.stock_map("map.txt")  # 25 free cells
    with pytest.raises(ValueError):
        mdb.DeliveryBatch([grid], 2, 26, 0, 10)  # more robots than free cells
    with pytest.raises(ValueError):
        mdb.DeliveryBatch([grid], 2, 2, 10, 1)  # randint(1, 1) for the 4th package on
    mdb.DeliveryBatch([grid], 2, 2, 3, 1).close()  # packages i <= min(R, 20) never draw a start time
    one = np.ones((3, 3), np.uint8)
    one[1, 1] = 0
    with pytest.raises(ValueError):
        mdb.DeliveryBatch([one], 1, 1, 1, 10)  # a single free cell: start != target can never be drawn
