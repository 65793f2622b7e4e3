"""CPU ORACLE (test infrastructure, NOT product code) -- the MAPPO trainer's package tracker as the trainer really
drives it, and the featurisers / shaped reward evaluated on THAT tracker instead of the env truth.

MAPPO/trainer.py keeps one dict per env (`persistent_packages_list`, :91) that `_update_persistent_packages_for_env`
(:95-130) extends with the packages a state lists as newly spawned, flips to 'in_transit' while a robot carries the id and
deletes once an 'in_transit' id is no longer carried.  The trainer NEVER clears the dict when an env auto-resets
(:229-234; QMIX :523-526, IDQ and the legacy qmix trainer do): 'waiting' entries of the previous episode survive, shadow
the new episode's packages with the same ids (`if pkg_id not in current_persistent_packages`), and -- because Python
dicts iterate in insertion order -- change the order in which the featurisers meet the packages.  Everything the helper
functions read from the dict (MAPPO/helper.py:6-369) is restated here on an insertion-ordered tracker; with a tracker that
is cleared on reset the results equal oracle_env's truth-based functions (asserted in tests/test_tracker_oracle.py).

Pinned by tests/test_tracker_oracle.py against tests/golden/traj_*_stale.npz and kats.npz:kat_b_stale_* (generated from the
unmodified reference by tests/golden/make_golden.py with the un-cleared tracker, KAT-B second sum 1243.89).
"""
from __future__ import annotations

import numpy as np

from .oracle_env import IN_TRANSIT, MAPPO_SHAPING, MV_S, WAITING, OracleEnv, split_action


class Tracker:
    """`persistent_packages_list[env_idx]` of MAPPO/trainer.py:91-130: insertion-ordered id -> [start cell, target cell,
    start_time, deadline, status]."""

    def __init__(self):
        self.d = {}

    def clear(self):
        self.d = {}

    def update(self, env: OracleEnv):
        """_update_persistent_packages_for_env(env_idx, state) with `state` = the env's current get_state()."""
        d = self.d
        for j in range(env.P):  # state['packages']: the packages whose start_time equals the state's time step (env.py:134-137)
            if env.p_t0[j] == env.t and (j + 1) not in d:
                d[j + 1] = [env.p_start[j], env.p_target[j], env.p_t0[j], env.p_deadline[j], WAITING]
        carried = {c for c in env.carry if c != 0}
        for pid in list(d):
            if pid in carried:
                d[pid][4] = IN_TRANSIT
            elif d[pid][4] == IN_TRANSIT:
                del d[pid]

    def snapshot(self):
        """[(id, start, target, start_time, deadline, status)] in dict order."""
        return [(pid, *v) for pid, v in self.d.items()]


def _lookup(trk, pid):
    for e in trk:
        if e[0] == pid:
            return e
    return None


def observation(env: OracleEnv, snap, trk, i):
    """convert_observation, MAPPO/helper.py:6-66, on a tracker snapshot."""
    H, W = env.H, env.W
    t, rcell, carry = snap["t"], snap["rcell"], snap["carry"]
    o = np.zeros((6, H * W), np.float32)
    o[0] = env.grid.reshape(-1)
    o[1, rcell[i]] = 1.0
    for k in range(env.R):
        if k != i:
            o[2, rcell[k]] = 1.0
    for (_, s, g, t0, _, st) in trk:
        if st == WAITING and t0 <= t:
            o[3, s] = 1.0
        if (st == WAITING and t0 <= t) or st == IN_TRANSIT:
            o[4, g] = 1.0
    e = _lookup(trk, carry[i]) if carry[i] else None
    if e is not None and e[5] == IN_TRANSIT:
        o[5, e[2]] = 1.0
    return o.reshape(6, H, W)


def vector_features(env: OracleEnv, snap, trk, i, K_o=100, K_p=100):
    """generate_vector_features, MAPPO/helper.py:68-165, on a tracker snapshot (stable sorts keep DICT order among ties)."""
    H, W, T = env.H, env.W, env.T
    t, rcell, carry = snap["t"], snap["rcell"], snap["carry"]
    r0, c0 = divmod(rcell[i], W)
    f = [r0 / H, c0 / W, 1.0 if carry[i] else 0.0]
    e = _lookup(trk, carry[i]) if carry[i] else None
    if e is not None and e[5] == IN_TRANSIT:
        tr, tc = divmod(e[2], W)
        f += [(tr - r0) / H, (tc - c0) / W, (max(0, e[4] - t) / T) if T > 0 else 0.0]
    else:
        f += [0.0, 0.0, 0.0]
    rows = []
    for k in range(env.R):
        if k == i:
            continue
        r, c = divmod(rcell[k], W)
        row = [(r - r0) / H, (c - c0) / W, 1.0 if carry[k] else 0.0]
        e = _lookup(trk, carry[k]) if carry[k] else None
        if e is not None and e[5] == IN_TRANSIT:
            tr, tc = divmod(e[2], W)
            row += [(tr - r) / H, (tc - c) / W]
        else:
            row += [0.0, 0.0]
        rows.append(row)
    rows.sort(key=lambda x: x[0] ** 2 + x[1] ** 2)
    for k in range(K_o):
        f += rows[k] if k < len(rows) else [0.0] * 5
    prow = []
    for (_, s, g, t0, dl, st) in trk:
        if st == WAITING and t0 <= t:
            sr, sc = divmod(s, W)
            tr, tc = divmod(g, W)
            prow.append([(sr - r0) / H, (sc - c0) / W, (tr - r0) / H, (tc - c0) / W, (max(0, dl - t) / T) if T > 0 else 0.0])
    prow.sort(key=lambda x: (x[4], x[0] ** 2 + x[1] ** 2))
    for k in range(K_p):
        f += prow[k] if k < len(prow) else [0.0] * 5
    f.append(t / T if T > 0 else 0.0)
    return np.array(f, dtype=np.float32)


def global_state(env: OracleEnv, snap, trk, MR=100, MP=100):
    """convert_global_state, MAPPO/helper.py:167-255, on a tracker snapshot (active packages sorted by id, :229)."""
    H, W, T = env.H, env.W, env.T
    t, rcell, carry = snap["t"], snap["rcell"], snap["carry"]
    g = np.zeros((4, H * W), np.float32)
    g[0] = env.grid.reshape(-1)
    for c in rcell:
        g[1, c] = 1.0
    for (_, s, tg, t0, _, st) in trk:
        if st == WAITING and t0 <= t:
            g[2, s] = 1.0
        if (st == WAITING and t0 <= t) or st == IN_TRANSIT:
            g[3, tg] = 1.0
    v = []
    for k in range(MR):
        if k < env.R:
            r, c = divmod(rcell[k], W)
            v += [r / H, c / W, 1.0 if carry[k] else 0.0]
            e = _lookup(trk, carry[k]) if carry[k] else None
            if e is not None and e[5] == IN_TRANSIT:
                tr, tc = divmod(e[2], W)
                v += [tr / H, tc / W, (max(0, e[4] - t) / T) if T > 0 else 0.0]
            else:
                v += [0.0, 0.0, 0.0]
        else:
            v += [0.0] * 6
    act = sorted((e for e in trk if (e[5] == WAITING and e[3] <= t) or e[5] == IN_TRANSIT), key=lambda e: e[0])
    for k in range(MP):
        if k < len(act):
            pid, s, tg, t0, dl, st = act[k]
            sr, sc = divmod(s, W)
            tr, tc = divmod(tg, W)
            wait = st == WAITING
            v += [sr / H, sc / W] if wait else [0.0, 0.0]
            v += [tr / H, tc / W, (max(0, dl - t) / T) if T > 0 else 0.0, 0.0 if wait else 1.0]
            cn = -1.0
            if not wait:
                for ridx in range(env.R):
                    if carry[ridx] == pid:
                        cn = ridx / (MR - 1) if MR > 1 else 0.0
                        break
            v.append(cn)
        else:
            v += [0.0] * 7
    v.append(t / T if T > 0 else 0.0)
    return g.reshape(4, H, W), np.array(v, dtype=np.float32)


def shaped_reward(env: OracleEnv, reward, prev, cur, actions, trk_prev, consts=MAPPO_SHAPING):
    """compute_shaped_rewards, MAPPO/helper.py:257-369, with `persistent_packages_at_prev_state` = trk_prev (a tracker
    snapshot taken BEFORE the step, as MAPPO/trainer.py:211-218 passes it)."""
    (c_pick, c_ontime, c_late, c_closer, c_wpick, c_wdrop, c_stuck, c_idle, c_away) = consts
    W = env.W
    pt = prev["t"]
    acc = np.zeros(env.R, dtype=np.float32)

    def md(a, b):
        return abs(a // W - b // W) + abs(a % W - b % W)

    waiting_prev = [e for e in trk_prev if e[5] == WAITING and e[3] <= pt]  # dict order
    for i in range(env.R):
        p0, p1 = prev["rcell"][i], cur["rcell"][i]
        k0, k1 = prev["carry"][i], cur["carry"][i]
        mv, op = split_action(actions[i])
        e0 = _lookup(trk_prev, k0) if k0 else None
        if k0 == 0 and k1 != 0:
            acc[i] += c_pick
        elif k0 != 0 and k1 == 0:
            if e0 is not None and p1 == e0[2]:
                acc[i] += c_ontime if cur["t"] <= e0[4] else c_late
        if op == 1:
            if k0 != 0:
                acc[i] += c_wpick
            elif k1 == 0:
                if not any(e[1] == p1 for e in waiting_prev):
                    acc[i] += c_wpick
        elif op == 2:
            if k0 == 0:
                acc[i] += c_wdrop
            elif k1 != 0:
                if e0 is not None and p1 != e0[2]:
                    acc[i] += c_wdrop
        moved = p0 != p1
        if mv != MV_S and not moved:
            acc[i] += c_stuck
        goal = None
        if e0 is not None:
            goal = e0[2]
        elif waiting_prev:
            goal = min((e[1] for e in waiting_prev), key=lambda pos: md(p0, pos))  # first minimum in dict order
        if goal is not None and moved:
            d0, d1 = md(p0, goal), md(p1, goal)
            if d1 < d0:
                acc[i] += c_closer
            elif d1 > d0:
                acc[i] += c_away
        if (not moved) and mv == MV_S and k0 == 0:
            if any(md(p0, e[1]) <= 3 for e in waiting_prev):
                acc[i] += c_idle
    return reward + acc.sum()


def mappo_rollout_step(venv, trackers, actions, K_o, K_p, MR=100, MP=100, consts=MAPPO_SHAPING, clear_on_reset=False):
    """One iteration of MAPPO/trainer.py:172-286 for every env with the trainer's own tracker handling: the shaped reward
    reads the tracker of the state before the step; a done env is reset and the tracker is updated with the RESET state only
    (never with the terminal one, and never cleared); otherwise it is updated with the next state.  clear_on_reset=True gives
    the QMIX / IDQ behaviour, which equals the env truth."""
    N = venv.num_envs
    R = venv.envs[0].R
    shaped = np.zeros(N, np.float32)
    dones = np.zeros(N, bool)
    obs, vec, gst, gvec = [], [], [], []
    for n, e in enumerate(venv.envs):
        trk = trackers[n]
        before = e.snapshot()
        pp_prev = trk.snapshot()
        r, d = e.step(actions[n])
        shaped[n] = shaped_reward(e, r, before, e.snapshot(), actions[n], pp_prev, consts)
        dones[n] = d
        if d:
            e.reset()
            if clear_on_reset:
                trk.clear()
        trk.update(e)
        s, ts = e.snapshot(), trk.snapshot()
        g, gv = global_state(e, s, ts, MR, MP)
        gst.append(g)
        gvec.append(gv)
        obs.append(np.stack([observation(e, s, ts, i) for i in range(R)]))
        vec.append(np.stack([vector_features(e, s, ts, i, K_o, K_p) for i in range(R)]))
    return shaped, dones, np.stack(obs), np.stack(vec), np.stack(gst), np.stack(gvec)
