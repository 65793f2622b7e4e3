import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marl_delivery_b200 as mdb
N, R = 4096, 5
ve = mdb.VectorizedEnv(mdb.Environment, N, map_file="map1.txt", n_robots=R, n_packages=50, max_time_steps=500, seed=42, k_others=4, k_packages=5)
b = ve.batch; b.reset()
acts = torch.randint(0, 15, (64, N, R), dtype=torch.uint8, device="cuda")
out = (torch.empty(N, dtype=torch.float64, device="cuda"), torch.empty(N, dtype=torch.float32, device="cuda"), torch.empty(N, dtype=torch.uint8, device="cuda"))
outn = b.step_n(acts)
def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {"pdl": os.environ.get("MDX_PDL", "default")}
res["step_n_us"] = timeit(lambda: b.step_n(acts, out=outn), 10) / 64 * 1e3
stp = b.stepper(out)
seq = [acts[i] for i in range(64)]
def loop():
    for a in seq: stp(a)
res["stepper_us"] = timeit(loop, 10) / 64 * 1e3
def loop2():
    for a in seq: b.step(a, auto_reset=True, out=out)
res["step_us"] = timeit(loop2, 10) / 64 * 1e3
g = ve.rollout_graph(64); g.run(acts)
res["graph_us"] = timeit(lambda: g.run(), 10) / 64 * 1e3
print(json.dumps(res))
