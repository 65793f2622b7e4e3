"""CPU-side checks of libmdx: it loads without a GPU, exports every symbol include/mdx.h declares,
and its host-only pieces (MT19937 reset sampler, lookup tables, numpy-sum restatement) agree with
numpy / plain Python arithmetic and with the reference's golden reset blocks.  No compute kernels
are launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import marl_delivery_b200 as mdb
from golden_util import load_reset_blocks
from marl_delivery_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mdx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mdx_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.mdx_version() == 2


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MdxError):
        mdb.Environment("map1.txt")


@pytest.mark.parametrize("lo,hi", [(0, 64), (0, 63), (0, 25), (5, 30), (10, 60), (1, 500), (1, 1000), (1, 2), (0, 197),
                                   (3, 21), (0, 1), (0, 2 ** 31)])
def test_randint_stream_matches_numpy(lo, hi):
    L = _lib.lib()
    for seed in (0, 1, 42, 2025, 2 ** 32 - 1):
        out = (C.c_int64 * 3000)()
        _lib.check(L.mdx_rng_randint(seed, lo, hi, 3000, out))
        ref = np.random.RandomState(seed).randint(lo, hi, size=None)  # scalar path first
        assert out[0] == ref
        rs = np.random.RandomState(seed)
        want = np.array([rs.randint(lo, hi) for _ in range(3000)])
        assert np.array_equal(np.frombuffer(out, dtype=np.int64), want)


def _sample_blocks(grid, R, P, T, seed, n_blocks):
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


def test_sampler_matches_golden_reset_blocks():
    meta, z = load_reset_blocks()
    for k, m in enumerate(meta):
        blocks = _sample_blocks(z[f"grid_{k}"], m["R"], m["P"], m["T"], m["seed"], 3)
        for b, (rc, pk) in enumerate(blocks):
            assert np.array_equal(rc, z[f"robots_{k}"][b]), (m, b)
            assert np.array_equal(pk.T, z[f"pkgs_{k}"][b]), (m, b)


def test_sampler_matches_oracle_random_configs():
    from oracle import oracle_env as orc

    rs = np.random.RandomState(123)
    for trial in range(25):
        name = ["map.txt", "map1.txt", "map3.txt", "map5.txt"][trial % 4]
        grid = mdb.maps.stock_map(name)
        nfree = int((grid == 0).sum())
        R = int(rs.randint(1, min(32, nfree) + 1))
        P = int(rs.randint(0, 120))
        T = int(rs.randint(2, 400))
        seed = int(rs.randint(0, 2 ** 31))
        e = orc.OracleEnv(grid, T, R, P, seed=seed)
        blocks = _sample_blocks(grid, R, P, T, seed, 3)
        for b, (rc, pk) in enumerate(blocks):
            if b:
                e.reset()
            assert rc.tolist() == e.rcell
            assert pk[0].tolist() == e.p_start and pk[1].tolist() == e.p_target
            assert pk[2].tolist() == e.p_t0 and pk[3].tolist() == e.p_deadline


def test_sampler_error_cases():
    L = _lib.lib()
    grid = mdb.maps.stock_map("map.txt")
    with pytest.raises(ValueError):  # 26 robots on 25 free cells: the reference's randint(0, 0) raises
        _sample_blocks(grid, 26, 0, 10, 1, 1)
    with pytest.raises(ValueError):  # T = 1: randint(1, 1) raises in the reference
        _sample_blocks(grid, 2, 10, 1, 1, 1)
    assert _sample_blocks(grid, 2, 3, 1, 1, 1)  # i <= min(R, 20) packages never draw a start time
    assert L.mdx_last_error() is not None


@pytest.mark.parametrize("n,n_out", [(7, 7), (10, 10), (20, 20), (10, 20), (13, 17)])
def test_quotient_lut_is_python_division(n, n_out):
    out = (C.c_float * (2 * n_out - 1))()
    _lib.check(_lib.lib().mdx_lut_quotient(n, n_out, out))
    want = np.array([d / n for d in range(-(n_out - 1), n_out)], dtype=np.float32)
    assert np.array_equal(np.frombuffer(out, dtype=np.float32).view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("rows,cols,ro,co", [(7, 7, 7, 7), (10, 10, 10, 10), (20, 20, 20, 20), (10, 10, 20, 20), (6, 9, 8, 11)])
def test_rank_lut_orders_like_python_float_keys(rows, cols, ro, co):
    """the order must be that of the float64 key x0**2 + x1**2 with x = d / n, which
    differs from the integer order of dr^2 + dc^2."""
    nr, nc = 2 * ro - 1, 2 * co - 1
    out = (C.c_uint16 * (nr * nc))()
    nranks = C.c_int32()
    _lib.check(_lib.lib().mdx_lut_rank(rows, cols, ro, co, out, C.byref(nranks)))
    rank = np.frombuffer(out, dtype=np.uint16).reshape(nr, nc)
    keys = np.array([[(dr / rows) ** 2 + (dc / cols) ** 2 for dc in range(-(co - 1), co)] for dr in range(-(ro - 1), ro)])
    uniq = np.unique(keys)
    assert nranks.value == len(uniq)
    assert np.array_equal(rank, np.searchsorted(uniq, keys))
    if rows == cols == ro == co == 10:  # the survey's counter-example: 0.1**2 + 0.7**2 < 0.5**2 + 0.5**2
        assert rank[9 + 1, 9 + 7] < rank[9 + 5, 9 + 5]


def test_np_sum_restatement_matches_numpy():
    L = _lib.lib()
    rs = np.random.RandomState(5)
    vals = np.array([5, 200, 20, 0.02, -0.05, -0.01, 0.0, 0.5, -0.1, -0.02, 2.0, 0.2], np.float32)
    for n in list(range(0, 33)) + [100, 128, 129, 300]:
        for _ in range(60):
            a = (rs.choice(vals, size=n) + rs.choice(vals, size=n)).astype(np.float32)
            got = L.mdx_np_sum_f32(a.ctypes.data_as(C.POINTER(C.c_float)), n)
            assert np.float32(got) == a.sum(), (n, a)


def test_action_encoding_mirrors_reference_comparisons():
    enc = mdb.encode_action
    assert [enc(m, "0") for m in "DLRSU"] == [0, 1, 2, 3, 4]
    assert enc("U", "1") == 9 and enc("S", "2") == 13
    assert enc("X", "1") == 16 and enc("stay", "0") == 15  # unknown move strings: stay, but not 'S'
    assert enc("L", 1) == 1 and enc("L", 2) == 1  # ints are not package ops (env.py:264,277 compare strings)
    assert enc("L", "3") == 1


def test_stock_maps():
    assert mdb.maps.stock_map("map1.txt").shape == (10, 10) and int((mdb.maps.stock_map("map1.txt") == 0).sum()) == 64
    assert int((mdb.maps.stock_map("map5.txt") == 0).sum()) == 197
    assert int((mdb.maps.stock_map("map.txt") == 0).sum()) == 25
    txt = "1 1 1\n1 0 1\n1 1 1"  # no trailing newline, like map4/map5
    assert mdb.maps.parse_map_text(txt).tolist() == [[1, 1, 1], [1, 0, 1], [1, 1, 1]]


def test_output_tensor_validation_refuses_host_tensors():
    """DeliveryBatch._check_out guards every kernel that writes caller-owned tensors; without a GPU its refusals can
    still be exercised on a bare instance."""
    import torch

    import marl_delivery_b200 as mdb

    b = mdb.DeliveryBatch.__new__(mdb.DeliveryBatch)
    b.device = torch.device("cuda", 0)
    b._ok_out = {}
    specs = (("reward", 4, torch.float64), ("shaped", 4, torch.float32), ("done", 4, torch.uint8))
    with pytest.raises(ValueError, match="CUDA tensor"):
        b._check_out((torch.zeros(4, dtype=torch.float64), None, None), specs)
    with pytest.raises(ValueError, match="expected 3 output tensors"):
        b._check_out((None, None), specs)
    b._check_out((None, None, None), specs)  # outputs that are not wanted are fine
    b.__dict__.clear()  # nothing to close
