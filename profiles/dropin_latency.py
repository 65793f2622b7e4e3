#!/usr/bin/env python
"""Latency of the reference-facing drop-in API at the reference's own scale (5 envs, map1, 5 robots, 50 packages):
microseconds per Environment.step, VectorizedEnv.step (5 envs, lists of dicts) and per helper call, next to the same calls of
the oracle port on one host core of the same box (the unmodified reference measured 26 / 34 / 54 / 203 us per env.step /
convert_observation / generate_vector_features / convert_global_state in the build container).
Prints one JSON line.  Run on the GPU box:  python profiles/dropin_latency.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import marl_delivery_b200 as mdb  # noqa: E402
from marl_delivery_b200 import helper  # noqa: E402
from oracle import oracle_env as orc  # noqa: E402


def us(fn, n):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    R, P, T = 5, 50, 500
    rs = np.random.RandomState(0)
    acts_code = rs.randint(0, 15, size=(4000, 5, R))
    to_act = lambda row: [("DLRSU"[int(x) % 5], str(int(x) // 5)) for x in row]  # noqa: E731
    # the callers hand over ready (move, op) tuples (their agents produce them); build them outside the timed calls
    one_env = [to_act(acts_code[k, 0]) for k in range(4000)]
    five_env = [[to_act(r) for r in acts_code[k]] for k in range(4000)]
    out = {}
    # ---- single Environment
    e = mdb.Environment("map1.txt", T, R, P, seed=42)
    e.reset()
    it = iter(range(10 ** 9))
    out["Environment.step_us"] = us(lambda: e.step(one_env[next(it) % 4000]), 1000)
    # ---- VectorizedEnv, 5 envs, list-of-dicts API
    ve = mdb.VectorizedEnv(mdb.Environment, 5, map_file="map1.txt", n_robots=R, n_packages=P, max_time_steps=T, seed=42)
    ve.reset()
    out["VectorizedEnv.step_5envs_us"] = us(lambda: ve.step(five_env[next(it) % 4000]), 1000)
    out["VectorizedEnv.step_per_env_us"] = out["VectorizedEnv.step_5envs_us"] / 5
    out["VectorizedEnv.reset_5envs_us"] = us(lambda: ve.reset(), 50)
    # ---- tensor fast path at the same size (what a trainer on this framework calls)
    import torch
    a_dev = torch.randint(0, 15, (5, R), dtype=torch.uint8, device="cuda")
    ve.batch.prefetch(0)

    def fast():
        ve.rollout_step(a_dev)
        torch.cuda.synchronize()
    out["rollout_step_5envs_sync_us"] = us(fast, 400)
    # ---- helper per-call functions on a mid-episode state
    states = ve.reset()
    for s in range(60):
        states, _, _, _ = ve.step([to_act(r) for r in acts_code[s]])
    st = states[0]
    env0 = ve.envs[0]
    tracker = {p.package_id: dict(id=p.package_id, start_pos=p.start, target_pos=p.target, start_time=p.start_time,
                                  deadline=p.deadline, status=p.status) for p in env0.packages if p.status in ("waiting", "in_transit")}
    out["convert_observation_us"] = us(lambda: helper.convert_observation(st, tracker, 1), 300)
    out["generate_vector_features_us"] = us(lambda: helper.generate_vector_features(st, tracker, 1, T, R - 1, 5), 300)
    out["convert_global_state_us"] = us(lambda: helper.convert_global_state(st, tracker, T), 300)
    nxt = ve.step([to_act(r) for r in acts_code[61]])[0][0]
    out["compute_shaped_rewards_us"] = us(lambda: helper.compute_shaped_rewards(0.0, st, nxt, to_act(acts_code[61, 0]), tracker, R), 300)
    # ---- the oracle port on one host core of this box
    grid = mdb.maps.stock_map("map1.txt")
    oe = orc.OracleEnv(grid, T, R, P, seed=42)
    oe.reset()
    out["port.Environment.step_us"] = us(lambda: oe.step(acts_code[next(it) % 4000, 0]), 2000)  # (the port takes action codes; the reference's own step parses the same tuples)
    snap = oe.snapshot()
    out["port.convert_observation_us"] = us(lambda: orc.observation(oe, snap, 1), 1000)
    out["port.generate_vector_features_us"] = us(lambda: orc.vector_features(oe, snap, 1, R - 1, 5), 1000)
    out["port.convert_global_state_us"] = us(lambda: orc.global_state(oe, snap), 300)
    out["reference_us_build_container"] = {"Environment.step": 26, "convert_observation": 34, "generate_vector_features": 54,
                                           "convert_global_state": 203, "source": "BASELINE.md section 2"}
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
