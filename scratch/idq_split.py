"""IDQ fused iteration (f3): one launch over all envs vs two env ranges on two streams, precompiled vs run-time specialised."""
import importlib, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mdb = importlib.import_module("marl-delivery_b200")
dev = torch.device("cuda:0")
N, R, K = 65536, 5, 100
res = {}
for spec in (False, True):
    if spec:
        os.environ["MDX_JIT_FORCE"] = "1"
    ve = mdb.VectorizedEnv(mdb.Environment, N, map_file="map1.txt", n_robots=R, n_packages=50, max_time_steps=500, seed=3, specialize=spec)
    b = ve.batch
    b.reset()
    H, W = b.rows_out, b.cols_out
    out = (torch.empty(N, dtype=torch.float64, device=dev), torch.empty((N, R), dtype=torch.float64, device=dev), torch.empty(N, dtype=torch.uint8, device=dev))
    feat = (torch.empty((N, R, 6, H, W), dtype=torch.float32, device=dev), torch.empty((N, 7, H, W), dtype=torch.float32, device=dev))
    acts = [torch.randint(0, 15, (N, R), dtype=torch.uint8, device=dev) for _ in range(16)]
    Bytes = 4 * (6 * R + 7) * H * W + 8 * R + 8 + 1 + R
    for parts in (1, 2, 3):
        bounds = [N * i // parts for i in range(parts + 1)]
        rng = [(lo, hi - lo) for lo, hi in zip(bounds[:-1], bounds[1:])]
        sts = [torch.cuda.Stream(dev) for _ in rng]
        args = [dict(out=tuple(t[lo:lo + n] for t in out), feat_out=tuple(t[lo:lo + n] for t in feat), env_range=(lo, n), stream=st)
                for (lo, n), st in zip(rng, sts)]
        sl = [[a[lo:lo + n] for (lo, n) in rng] for a in acts]
        def run(k):
            for s in range(k):
                for j, kw in enumerate(args):
                    b.step_idq_featurize(sl[s % 16][j], auto_reset=True, ops_as_int=True, **kw)
        run(20)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        fork = torch.cuda.Event()
        fork.record()
        for st in sts:
            st.wait_event(fork)
        ev[0].record(sts[0])
        t0 = time.perf_counter()
        run(K)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / K
        res[f"jit={b.jit_status()} parts={parts}"] = dict(ms_wall=round(wall, 4), M_env_steps=round(N / wall / 1e3, 1), GBs=round(Bytes * N / wall / 1e6, 1))
    del ve, b
print(json.dumps(res, indent=1))
