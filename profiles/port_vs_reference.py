"""Measures oracle port vs the unmodified reference per core on the MAPPO rollout iteration (build container only)."""
import sys, time, types, os
import numpy as np
sys.path.insert(0, "/root/repo")
sys.modules.setdefault("pygame", types.ModuleType("pygame"))
sys.path.insert(0, "/root/reference")
import env as ref_env
from MAPPO import helper as H
from MAPPO.env_vectorized import VectorizedEnv
from oracle import oracle_env as orc
sys.path.insert(0, "/root/repo/tests/golden")

def decode(a):
    return ("DLRSU"[a % 5], str(a // 5))

N, R, P, T, steps = 5, 5, 50, 500, 400
# ---- reference: the trainer's loop (MAPPO/trainer.py:208-280) minus torch
ve = VectorizedEnv(ref_env.Environment, N, map_file="/root/reference/map1.txt", n_robots=R, n_packages=P, max_time_steps=T, seed=42)
states = ve.reset()
pp = [{} for _ in range(N)]
def upd(i, st):
    d = pp[i]
    for t in st['packages']:
        if t[0] not in d:
            d[t[0]] = dict(id=t[0], start_pos=(t[1]-1, t[2]-1), target_pos=(t[3]-1, t[4]-1), start_time=t[5], deadline=t[6], status='waiting')
    carried = {r[2] for r in st['robots'] if r[2] != 0}
    for k, v in list(d.items()):
        if k in carried: v['status'] = 'in_transit'
        elif v['status'] == 'in_transit': del d[k]
for i in range(N): upd(i, states[i])
rs = np.random.RandomState(0)
acts_all = rs.randint(0, 15, size=(steps, N, R))
t0 = time.perf_counter()
for s in range(steps):
    acts = [[decode(int(x)) for x in row] for row in acts_all[s]]
    nxt, r, d, _ = ve.step(acts)
    for i in range(N):
        H.compute_shaped_rewards(r[i], states[i], nxt[i], acts[i], pp[i], R)
        if d[i]:
            nxt[i] = ve.envs[i].reset(); pp[i] = {}
        upd(i, nxt[i])
        H.convert_global_state(nxt[i], pp[i], T)[0]; H.convert_global_state(nxt[i], pp[i], T)[1]
        for a in range(R):
            H.convert_observation(nxt[i], pp[i], a)
            H.generate_vector_features(nxt[i], pp[i], a, T, R - 1, 5)
    states = nxt
t_ref = time.perf_counter() - t0
grid = orc.parse_map_text(open("/root/reference/map1.txt").read())
ov = orc.OracleVecEnv(N, grid, seed=42, max_time_steps=T, n_robots=R, n_packages=P)
ov.reset()
t0 = time.perf_counter()
for s in range(steps):
    orc.mappo_rollout_step(ov, acts_all[s], R - 1, 5)
t_port = time.perf_counter() - t0
print("reference env-steps/s/core", N * steps / t_ref, "port", N * steps / t_port, "port/ref", t_ref / t_port)
