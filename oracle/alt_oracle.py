"""CPU ORACLE (test infrastructure, NOT product code): restatement of the other featurisers / shaping
that sit on the same boundary, on the flat-state OracleEnv:

  idq_observation        IDQ/networks.py:112-217 convert_state (same code as qmix/networks.py:243-348)
  legacy_global_tensor   qmix/networks.py:350-468 convert_global_state_to_tensor with shape (7, H, W)
  idq_reward_shaping     IDQ/networks.py:228-349 reward_shaping

The package dict is the env truth restricted to waiting / in-transit packages (== the trainers' tracker).
Pinned by tests/test_alt_oracle.py against tests/golden/kats_alt.npz (recorded from the unmodified
reference by tests/golden/make_golden.py alt).
"""
from __future__ import annotations

import numpy as np

from .oracle_env import IN_TRANSIT, WAITING, split_action

# IDQ/networks.py:220-225
IDQ_PICKUP, IDQ_DELIVERY, IDQ_LATE, IDQ_WASTED_PICKUP, IDQ_WASTED_DROP, IDQ_STAY = 2, 10, -5, -0.1, 0, -0.1


def _urgency(t, st, dl):
    """IDQ/networks.py:172-179 / qmix/networks.py:442-446 (the caller has checked t >= st)."""
    if dl > st:
        return min(1.0, max(0.0, (t - st) / (dl - st)))
    if dl == st:
        return 1.0
    return 0.0


def idq_observation(env, snap, i):
    H, W = env.H, env.W
    t, rcell, carry, status = snap["t"], snap["rcell"], snap["carry"], snap["status"]
    o = np.zeros((6, H * W), np.float32)
    o[0] = env.grid.reshape(-1)
    if carry[i] == 0:  # :161-186
        for j in range(env.P):
            if status[j] == WAITING and t >= env.p_t0[j]:
                c = env.p_start[j]
                o[1, c] = max(o[1, c], _urgency(t, env.p_t0[j], env.p_deadline[j]))
                o[2, c] = 1.0
    for k in range(env.R):  # :189-195
        if k != i:
            o[3, rcell[k]] = 1.0
    o[4, rcell[i]] = 1.0  # :199-202
    if carry[i] != 0 and status[carry[i] - 1] in (WAITING, IN_TRANSIT):  # :205-213 "in persistent_packages"
        o[5, env.p_target[carry[i] - 1]] = 1.0
    return o.reshape(6, H, W)


def legacy_global_tensor(env, snap):
    H, W = env.H, env.W
    t, rcell, carry, status = snap["t"], snap["rcell"], snap["carry"], snap["status"]
    g = np.zeros((7, H * W), np.float32)
    g[0] = env.grid.reshape(-1)  # shape == map shape: the centring / cropping of :385-407 is the identity
    for k in range(env.R):  # :413-424
        g[1, rcell[k]] = 1.0
        if carry[k] != 0:
            g[2, rcell[k]] = 1.0
    for j in range(env.P):  # :428-466
        if status[j] not in (WAITING, IN_TRANSIT) or t < env.p_t0[j]:
            continue
        if status[j] == WAITING:
            c = env.p_start[j]
            g[3, c] = 1.0
            g[6, c] = max(g[6, c], _urgency(t, env.p_t0[j], env.p_deadline[j]))
            g[4, env.p_target[j]] = 1.0
        else:
            g[5, env.p_target[j]] = 1.0
    return g.reshape(7, H, W)


def idq_reward_shaping(env, prev, cur, actions, ops_as_int):
    """-> list of per-agent Python floats.  `ops_as_int` False is what the IDQ trainer does (it passes
    (move_str, op_str) tuples, IDQ/trainer.py:276-291, so `package_op == 1` never holds and only the
    stay penalty fires); True evaluates every branch."""
    pstat, pt = prev["status"], prev["t"]
    out = []
    for i in range(env.R):
        acc = 0
        _, op = split_action(actions[i])
        p0, p1 = prev["rcell"][i], cur["rcell"][i]
        k0, k1 = prev["carry"][i], cur["carry"][i]
        if p0 == p1:  # :283-284
            acc += IDQ_STAY
        if ops_as_int:
            if op == 1:  # :300-318
                if k0 == 0 and k1 != 0:
                    acc += IDQ_PICKUP
                elif k0 != 0:
                    acc += IDQ_WASTED_PICKUP
                elif k0 == 0 and k1 == 0:
                    if not any(pstat[j] == WAITING and env.p_start[j] == p0 and env.p_t0[j] <= pt for j in range(env.P)):
                        acc += IDQ_WASTED_PICKUP
            elif op == 2:  # :321-347
                if k0 != 0 and k1 == 0:
                    if pstat[k0 - 1] in (WAITING, IN_TRANSIT):
                        if p1 == env.p_target[k0 - 1]:
                            acc += IDQ_DELIVERY
                            if cur["t"] > env.p_deadline[k0 - 1]:
                                acc += IDQ_LATE
                elif k0 == 0:
                    acc += IDQ_WASTED_DROP
        out.append(acc)
    return out
