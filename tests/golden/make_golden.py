#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it imports /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

Nothing from the reference is copied: this script imports `env.Environment`,
`MAPPO.env_vectorized.VectorizedEnv`, `MAPPO.helper.*`, `QMIX.helper.*`, `greedyagent` and
`randomagent` from /root/reference (with an empty `pygame` stub), drives
them, and stores INPUTS and OUTPUTS as compressed .npz files.  Tests only read the .npz files.

Conventions used in every fixture (all indices 0-based, cell = r * W + c):
  action code a:  0..14 -> move = 'DLRSU'[a % 5], package op = str(a // 5)   
                  15..17 -> move = 'X' (not a legal move string), package op = str(a - 15)
  pkg status u8:  0 'None', 1 'waiting', 2 'in_transit', 3 'delivered'
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

sys.modules.setdefault("pygame", types.ModuleType("pygame"))
sys.path.insert(0, REF)

import env as ref_env  # noqa: E402
import greedyagent as ref_greedy  # noqa: E402
import randomagent as ref_random  # noqa: E402
from MAPPO import helper as mappo_helper  # noqa: E402
from MAPPO.env_vectorized import VectorizedEnv as MappoVec  # noqa: E402

# QMIX/helper.py imports yaml (present) -- load it as a separate module object
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("qmix_helper_ref", os.path.join(REF, "QMIX", "helper.py"))
qmix_helper = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(qmix_helper)

MOVES = "DLRSU"
STATUS = {"None": 0, "waiting": 1, "in_transit": 2, "delivered": 3}


def decode(a):
    a = int(a)
    if a < 15:
        return (MOVES[a % 5], str(a // 5))
    return ("X", str(a - 15))


def encode(move, op):
    return MOVES.index(move) + 5 * int(op)


def grid_of(map_name):
    with open(os.path.join(REF, map_name)) as f:  # (os.path.join keeps an absolute map_name as it is: the synthetic maps of gen_shapes)
        return np.array([[int(x) for x in line.strip().split(" ")] for line in f], dtype=np.uint8)


def truth_tracker(e):
    """env-truth package dict (== the trainers' tracker in cleared mode)."""
    d = {}
    for p in e.packages:
        if p.status in ("waiting", "in_transit"):
            d[p.package_id] = {
                "id": p.package_id,
                "start_pos": p.start,
                "target_pos": p.target,
                "start_time": p.start_time,
                "deadline": p.deadline,
                "status": p.status,
            }
    return d


class Tracker:
    """The trainers' tracker, driven exactly like MAPPO/trainer.py:95-130 does (by calling the
    reference method on a shim object), so the fixture also proves tracker == env truth."""

    def __init__(self, n):
        # MAPPO/trainer.py does `from networks import ...`, so its own directory must be importable
        if os.path.join(REF, "MAPPO") not in sys.path:
            sys.path.append(os.path.join(REF, "MAPPO"))
        # matplotlib is absent here and only used by the trainer's end-of-run plots
        if "matplotlib" not in sys.modules:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = mpl.pyplot
        with contextlib.redirect_stdout(io.StringIO()):
            from MAPPO import trainer as _t  # heavy import (torch, sklearn) but unmodified

        self._fn = _t.MAPPOTrainer._update_persistent_packages_for_env
        self.persistent_packages_list = [{} for _ in range(n)]

    def update(self, i, state):
        self._fn(self, i, state)

    def clear(self, i):
        self.persistent_packages_list[i] = {}


def snapshot_blocks(map_name, R, P, T, seed, n_blocks):
    e = ref_env.Environment(os.path.join(REF, map_name), T, R, P, seed=seed)
    W = e.n_cols
    robots = np.zeros((n_blocks, R), np.int32)
    pk = np.zeros((n_blocks, P, 4), np.int32)  # start cell, target cell, start_time, deadline
    for b in range(n_blocks):
        if b > 0:
            e.reset()
        robots[b] = [r.position[0] * W + r.position[1] for r in e.robots]
        for j, p in enumerate(e.packages):
            assert p.package_id == j + 1
            pk[b, j] = (p.start[0] * W + p.start[1], p.target[0] * W + p.target[1], p.start_time, p.deadline)
    return robots, pk


def gen_reset_blocks():
    cases = [
        ("map1.txt", 5, 50, 500, 2025),
        ("map1.txt", 5, 50, 500, 42),
        ("map1.txt", 5, 100, 1000, 10),
        ("map.txt", 3, 8, 40, 7),
        ("map.txt", 20, 30, 60, 3),
        ("map2.txt", 10, 100, 300, 5),
        ("map3.txt", 7, 20, 100, 6),
        ("map4.txt", 10, 100, 500, 8),
        ("map5.txt", 10, 100, 200, 11),
        ("map5.txt", 32, 500, 1000, 123),
        ("map1.txt", 5, 5, 2, 9),  # T=2: randint(1, 2) has range 0 -> no draw
        ("map1.txt", 25, 3, 50, 4),  # more robots than packages; i <= min(R, 20) rule
    ]
    out = {}
    meta = []
    for k, (m, R, P, T, seed) in enumerate(cases):
        robots, pk = snapshot_blocks(m, R, P, T, seed, 3)
        out[f"robots_{k}"] = robots
        out[f"pkgs_{k}"] = pk
        out[f"grid_{k}"] = grid_of(m)
        meta.append(dict(map=m, R=R, P=P, T=T, seed=seed))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "reset_blocks.npz"), **out)
    print("reset_blocks.npz", len(cases), "cases")


def run_traj(name, map_name, R, P, T, seed, n_steps, act_seed, policy, K_o, K_p, MR=100, MP=100,
             feat_every=1, big_vec_every=0, invalid_frac=0.0, stale=False, rewards=None):
    """One env driven like one slot of MAPPO/trainer.py:collect_rollouts (auto-reset on done, the
    features recorded for a step are those of the state the NEXT action is taken from).

    stale=True keeps the tracker exactly as MAPPO/trainer.py:229-234 does -- it is NOT cleared when the env resets (SURVEY.md
    Q7), so 'waiting' entries of earlier episodes survive, shadow new packages with the same id and change the dict order;
    the tracker's ids (in dict order) and statuses are recorded per step as well."""
    path = os.path.join(REF, map_name)
    # rewards = (move_cost, delivery_reward, delay_reward) other than the constructor's defaults (env.py:20-22)
    e = ref_env.Environment(path, T, R, P, seed=seed) if rewards is None else ref_env.Environment(path, T, R, P, *rewards, seed=seed)
    H, W = e.n_rows, e.n_cols
    state = e.reset()
    trk = Tracker(1)
    trk.update(0, state)
    rs = np.random.RandomState(act_seed)
    greedy = None
    if policy not in ("random", "homing"):
        greedy = ref_greedy.GreedyAgents()
        greedy.init_agents(state)

    D = 6 + 5 * K_o + 5 * K_p + 1
    Dg = 6 * MR + 7 * MP + 1
    rec = dict(
        actions=np.zeros((n_steps, R), np.uint8),
        robots=np.zeros((n_steps, R, 3), np.int32),  # r, c, carry AFTER the step (pre auto-reset)
        reward=np.zeros(n_steps, np.float64),
        done=np.zeros(n_steps, np.uint8),
        t_after=np.zeros(n_steps, np.int32),
        status=np.zeros((n_steps, P), np.uint8),  # after the step incl. spawn (pre auto-reset)
        total_reward=np.zeros(n_steps, np.float64),
        shaped_mappo=np.zeros(n_steps, np.float32),
        shaped_qmix=np.zeros(n_steps, np.float32),
    )
    feat_steps = [s for s in range(n_steps) if s % feat_every == 0]
    obs = np.zeros((len(feat_steps), R, 6, H, W), np.uint8)
    gst = np.zeros((len(feat_steps), 4, H, W), np.uint8)
    vec = np.zeros((len(feat_steps), R, D), np.float32)
    gvec = np.zeros((len(feat_steps), Dg), np.float32)
    big_steps = [s for s in range(n_steps) if big_vec_every and s % big_vec_every == 0]
    vec_big = np.zeros((len(big_steps), R, 1007), np.float32)
    gvec_q = np.zeros((len(big_steps), 6 * 10 + 7 * 20 + 1), np.float32)
    # features of the very first state (before any step)
    feat0 = None
    robots0 = np.array([[r.position[0], r.position[1], r.carrying] for r in e.robots], np.int32)
    pk0 = np.array([[p.start[0] * W + p.start[1], p.target[0] * W + p.target[1], p.start_time, p.deadline]
                    for p in e.packages], np.int32)
    reset_pkgs = [pk0]
    reset_robots = [robots0[:, 0] * W + robots0[:, 1]]

    def feats(st, pp):
        o = np.stack([mappo_helper.convert_observation(st, pp, i) for i in range(R)])
        v = np.stack([mappo_helper.generate_vector_features(st, pp, i, T, K_o, K_p) for i in range(R)])
        g, gv = mappo_helper.convert_global_state(st, pp, T, MR, MP)
        assert set(np.unique(o)) <= {0.0, 1.0} and set(np.unique(g)) <= {0.0, 1.0}
        return o.astype(np.uint8), v, g.astype(np.uint8), gv

    feat0 = feats(state, trk.persistent_packages_list[0])
    fi = bi = 0
    n_tracker_mismatch = 0
    trk_ids = np.zeros((n_steps, P), np.int32)     # dict order after the step's tracker update, 0-padded
    trk_status = np.zeros((n_steps, P), np.uint8)
    n_shadowed = 0
    for s in range(n_steps):
        a = rs.randint(0, 15, size=R)
        if greedy is not None:
            with contextlib.redirect_stdout(io.StringIO()):
                ga = greedy.get_actions(state)
            use_g = rs.rand(R) < (1.0 if policy == "greedy" else 0.7)
            for i in range(R):
                if use_g[i]:
                    a[i] = encode(*ga[i])
        if policy == "homing":
            # test-harness policy (not reference code): walk straight at the goal, so that many
            # deliveries (on time and late) and all-delivered terminations are exercised
            jitter = rs.rand(R) < 0.15
            for i, rb in enumerate(e.robots):
                if jitter[i]:
                    continue
                if rb.carrying:
                    goal, op_at = e.packages[rb.carrying - 1].target, "2"
                else:
                    wait = [p for p in e.packages if p.status == "waiting"]
                    if not wait:
                        continue
                    goal = min(wait, key=lambda p: abs(p.start[0] - rb.position[0]) + abs(p.start[1] - rb.position[1])).start
                    op_at = "1"
                dr, dc = goal[0] - rb.position[0], goal[1] - rb.position[1]
                if dr == 0 and dc == 0:
                    a[i] = encode("S", op_at)
                else:
                    if dr != 0 and (dc == 0 or rs.rand() < 0.5):
                        mv, nr, nc = ("D" if dr > 0 else "U"), rb.position[0] + (1 if dr > 0 else -1), rb.position[1]
                    else:
                        mv, nr, nc = ("R" if dc > 0 else "L"), rb.position[0], rb.position[1] + (1 if dc > 0 else -1)
                    a[i] = encode(mv, op_at if (nr, nc) == goal else "0")
        if invalid_frac > 0:
            inv = rs.rand(R) < invalid_frac
            a = np.where(inv, rs.randint(15, 18, size=R), a)
        acts = [decode(x) for x in a]
        pp_prev = trk.persistent_packages_list[0]
        nxt, r, done, info = e.step(acts)
        rec["actions"][s] = a
        rec["robots"][s] = [[rb.position[0], rb.position[1], rb.carrying] for rb in e.robots]
        rec["reward"][s] = r
        rec["done"][s] = done
        rec["t_after"][s] = e.t
        rec["status"][s] = [STATUS[p.status] for p in e.packages]
        rec["total_reward"][s] = e.total_reward
        rec["shaped_mappo"][s] = mappo_helper.compute_shaped_rewards(r, state, nxt, acts, pp_prev, R)
        rec["shaped_qmix"][s] = qmix_helper.compute_shaped_rewards(r, state, nxt, acts, pp_prev, R)
        if done:
            nxt = e.reset()
            if not stale:
                trk.clear(0)  # env-truth semantics (QMIX/IDQ/legacy clear; default of the build)
            trk.update(0, nxt)
            reset_pkgs.append(np.array([[p.start[0] * W + p.start[1], p.target[0] * W + p.target[1],
                                         p.start_time, p.deadline] for p in e.packages], np.int32))
            reset_robots.append(np.array([rb.position[0] * W + rb.position[1] for rb in e.robots], np.int32))
            if greedy is not None:
                greedy = ref_greedy.GreedyAgents()
                greedy.init_agents(nxt)
        else:
            trk.update(0, nxt)
        pp = trk.persistent_packages_list[0]
        if pp != truth_tracker(e) or list(pp) != list(truth_tracker(e)):
            n_tracker_mismatch += 1
        for k, (pid, pv) in enumerate(pp.items()):
            trk_ids[s, k] = pid
            trk_status[s, k] = STATUS[pv["status"]]
            if (pv["start_pos"], pv["deadline"]) != (e.packages[pid - 1].start, e.packages[pid - 1].deadline):
                n_shadowed += 1
        state = nxt
        if s % feat_every == 0:
            obs[fi], vec[fi], gst[fi], gvec[fi] = feats(state, pp)
            fi += 1
        if big_vec_every and s % big_vec_every == 0:
            vec_big[bi] = np.stack([mappo_helper.generate_vector_features(state, pp, i, T) for i in range(R)])
            gvec_q[bi] = qmix_helper.convert_global_state(state, pp, T, 10, 20)[1]
            bi += 1
    assert stale or n_tracker_mismatch == 0, n_tracker_mismatch
    assert not stale or (n_tracker_mismatch > 0 and n_shadowed > 0), "the stale fixture must exercise the artefact"
    meta = dict(rewards=(list(rewards) if rewards else None), stale=bool(stale), n_tracker_mismatch=n_tracker_mismatch, n_shadowed=n_shadowed, name=name, map=map_name, R=R, P=P, T=T, seed=seed, n_steps=n_steps, act_seed=act_seed,
                policy=policy, K_o=K_o, K_p=K_p, MR=MR, MP=MP, feat_every=feat_every,
                big_vec_every=big_vec_every, H=H, W=W, n_resets=len(reset_pkgs),
                n_delivered_total=int((rec["reward"] > 0.5).sum()))
    np.savez_compressed(
        os.path.join(HERE, f"traj_{name}.npz"), meta=np.array(json.dumps(meta)), grid=grid_of(map_name),
        reset_pkgs=np.stack(reset_pkgs), reset_robots=np.stack(reset_robots),
        obs0=feat0[0], vec0=feat0[1], gstate0=feat0[2], gvec0=feat0[3],
        obs=obs, vec=vec, gstate=gst, gvec=gvec, vec_big=vec_big, gvec_qmix=gvec_q,
        feat_steps=np.array(feat_steps, np.int32), big_steps=np.array(big_steps, np.int32), trk_ids=trk_ids,
        trk_status=trk_status, **rec)
    sz = os.path.getsize(os.path.join(HERE, f"traj_{name}.npz"))
    print(f"traj_{name}.npz  {sz/1024:.0f} KiB  resets={len(reset_pkgs)-1} "
          f"sum_r={rec['reward'].sum():.4f} deliveries~{meta['n_delivered_total']} "
          f"shaped_sum={rec['shaped_mappo'].astype(np.float64).sum():.4f}")


def gen_kat_a():
    """KAT-A of SURVEY.md section 4: seed 2025, actions RandomState(7).randint(0,15,size=5)."""
    e = ref_env.Environment(os.path.join(REF, "map1.txt"), 500, 5, 50, seed=2025)
    st = e.reset()
    rs = np.random.RandomState(7)
    acts = np.zeros((500, 5), np.uint8)
    rew = np.zeros(500)
    for s in range(500):
        a = rs.randint(0, 15, size=5)
        acts[s] = a
        st, r, done, info = e.step([decode(x) for x in a])
        rew[s] = r
    assert done
    delivered = sum(p.status == "delivered" for p in e.packages)
    print("KAT-A total", repr(e.total_reward), "delivered", delivered, "robots0", st["robots"])
    return dict(kat_a_actions=acts, kat_a_rewards=rew, kat_a_total=np.float64(e.total_reward),
                kat_a_delivered=np.int32(delivered))


def gen_kat_b(stale=False):
    """KAT-B: 5 vectorised envs, 1200 steps with auto-reset; cleared tracker (sum 1627.37), or -- stale=True -- the tracker
    handling of MAPPO/trainer.py:229-234 itself, which never clears it (sum 1243.89)."""
    ve = MappoVec(ref_env.Environment, 5, map_file=os.path.join(REF, "map1.txt"), n_robots=5, n_packages=50,
                  max_time_steps=500, seed=42)
    states = ve.reset()
    trk = Tracker(5)
    for i in range(5):
        trk.update(i, states[i])
    rs = np.random.RandomState(0)
    shaped = np.zeros((1200, 5), np.float32)
    envr = np.zeros((1200, 5), np.float64)
    dones = np.zeros((1200, 5), np.uint8)
    gsum = np.zeros((1200, 5), np.float64)
    for s in range(1200):
        a = rs.randint(0, 15, size=(5, 5))
        acts = [[decode(x) for x in row] for row in a]
        nxt, r, d, _ = ve.step(acts)
        for i in range(5):
            shaped[s, i] = mappo_helper.compute_shaped_rewards(r[i], states[i], nxt[i], acts[i],
                                                               trk.persistent_packages_list[i], 5)
            envr[s, i] = r[i]
            dones[s, i] = d[i]
            if d[i]:
                nxt[i] = ve.envs[i].reset()
                if not stale:
                    trk.clear(i)
            trk.update(i, nxt[i])
            v = mappo_helper.generate_vector_features(nxt[i], trk.persistent_packages_list[i], 0, 500, 4, 5)
            g = mappo_helper.convert_global_state(nxt[i], trk.persistent_packages_list[i], 500)[1]
            gsum[s, i] = float(v.astype(np.float64).sum() + g.astype(np.float64).sum())
        states = nxt
    print("KAT-B shaped sum", shaped.astype(np.float64).sum(), "(stale tracker)" if stale else "(cleared tracker)")
    pre = "kat_b_stale_" if stale else "kat_b_"
    return {pre + "shaped": shaped, pre + "reward": envr, pre + "done": dones, pre + "featsum": gsum}


def gen_readme_random():
    """README.md:117-121 Random row: seed 10, map1, 5 robots, 100 packages, T=1000, 100 episodes.
    Same call pattern as evaluation.py:9-66 + randomagent.py:16-22 (global numpy RNG)."""
    np.random.seed(10)
    tmp = ref_env.Environment(os.path.join(REF, "map1.txt"), 1000, 5, 100, seed=10)
    tmp.reset()
    rewards, delivered = [], []
    all_actions = np.zeros((100, 1000, 5), np.uint8)
    for ep in range(100):
        e = ref_env.Environment(os.path.join(REF, "map1.txt"), 1000, 5, 100, seed=10 + ep)
        st = e.reset()
        ag = ref_random.RandomAgents()
        ag.init_agents(st)
        done = False
        s = 0
        while not done:
            acts = ag.get_actions(st)
            all_actions[ep, s] = [encode(m, o) for (m, o) in acts]
            st, r, done, info = e.step(acts)
            s += 1
        assert s == 1000
        rewards.append(info["total_reward"])
        delivered.append(sum(p.status == "delivered" for p in e.packages))
    print("README random: %.2f +- %.2f, delivered %.2f +- %.2f" % (
        np.mean(rewards), np.std(rewards), np.mean(delivered), np.std(delivered)))
    # actions are reproducible from np.random.seed(10) + np.random.choice calls; store a digest plus
    # the first episode to pin the generator, tests regenerate the rest
    import hashlib
    return dict(readme_rewards=np.array(rewards, np.float64), readme_delivered=np.array(delivered, np.int32),
                readme_actions_ep0=all_actions[0],
                readme_actions_sha=np.array(hashlib.sha256(all_actions.tobytes()).hexdigest()))


def gen_greedy():
    """README.md:117-121 Greedy row (seed 10, map1, 5 robots, 100 packages, T=1000, 100 episodes; a new
    GreedyAgents per episode like evaluation.py:9-66) plus two wall-heavy trajectories, with the agent's
    actions recorded per step.  greedyagent.py prints on every call; stdout is discarded."""
    out = {}
    rewards, delivered = [], []
    acts0 = None
    for ep in range(100):
        e = ref_env.Environment(os.path.join(REF, "map1.txt"), 1000, 5, 100, seed=10 + ep)
        st = e.reset()
        ag = ref_greedy.GreedyAgents()
        ag.init_agents(st)
        done = False
        rec = []
        while not done:
            with contextlib.redirect_stdout(io.StringIO()):
                acts = ag.get_actions(st)
            rec.append([encode(m, o) for (m, o) in acts])
            st, r, done, info = e.step(acts)
        rewards.append(info["total_reward"])
        delivered.append(sum(p.status == "delivered" for p in e.packages))
        if ep < 3:
            out[f"readme_greedy_actions_ep{ep}"] = np.array(rec, np.uint8)
    print("README greedy: %.2f +- %.2f, delivered %.2f +- %.2f" % (
        np.mean(rewards), np.std(rewards), np.mean(delivered), np.std(delivered)))
    out["readme_greedy_rewards"] = np.array(rewards, np.float64)
    out["readme_greedy_delivered"] = np.array(delivered, np.int32)
    import tempfile

    tmp = tempfile.mkdtemp()

    def synthetic(name, H, W, pct, seed):  # non-square maps the reference does not ship, written where it can read them
        g = (np.random.RandomState(seed).rand(H, W) < pct / 100.0).astype(np.uint8)
        path = os.path.join(tmp, f"{name}.txt")
        with open(path, "w") as f:
            f.write("\n".join(" ".join(str(int(x)) for x in row) for row in g) + "\n")
        return path

    cases = [("map5.txt", 4, 30, 300, 3), ("map2.txt", 10, 60, 200, 8), ("map.txt", 3, 8, 80, 21), ("map4.txt", 7, 40, 150, 5),
             (synthetic("greedy_rect_8x13", 8, 13, 15, 3001), 4, 20, 120, 9), (synthetic("greedy_rect_11x6", 11, 6, 12, 3002), 3, 10, 90, 10)]
    meta = []
    for k, (m, R, P, T, seed) in enumerate(cases):
        e = ref_env.Environment(os.path.join(REF, m), T, R, P, seed=seed)
        st = e.reset()
        ag = ref_greedy.GreedyAgents()
        ag.init_agents(st)
        rec, rew = [], []
        done = False
        while not done:
            with contextlib.redirect_stdout(io.StringIO()):
                acts = ag.get_actions(st)
            rec.append([encode(mm, o) for (mm, o) in acts])
            st, r, done, info = e.step(acts)
            rew.append(r)
        out[f"greedy_actions_{k}"] = np.array(rec, np.uint8)
        out[f"greedy_rewards_{k}"] = np.array(rew, np.float64)
        out[f"greedy_grid_{k}"] = grid_of(m)
        meta.append(dict(map=os.path.basename(m), R=R, P=P, T=T, seed=seed, steps=len(rec), total=float(e.total_reward),
                         delivered=int(sum(p.status == "delivered" for p in e.packages))))
        print("greedy case", meta[-1])
    out["greedy_meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "kats_greedy.npz"), **out)
    print("kats_greedy.npz", os.path.getsize(os.path.join(HERE, "kats_greedy.npz")) // 1024, "KiB")


def _load_ref_module(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def gen_alt():
    """the other featurisers / shaping of the same boundary, recorded from the
    unmodified reference along env trajectories -- IDQ/networks.py convert_state (:112-217; identical code in
    qmix/networks.py:243-348, asserted), qmix/networks.py convert_global_state_to_tensor (:350-468) with the
    trainer's (7, H, W) shape, IDQ/networks.py reward_shaping (:228-349) both with the (move_str, op_str)
    tuples the IDQ trainer passes (only the stay penalty can fire: '1' == 1 is False) and with integer package
    ops (every branch), QMIX/helper.py generate_vector_features / convert_global_state with the QMIX/config.yaml
    sizes (10 / 20)."""
    idq = _load_ref_module("idq_networks_ref", "IDQ/networks.py")
    legacy = _load_ref_module("legacy_qmix_networks_ref", "qmix/networks.py")
    import tempfile

    tmp = tempfile.mkdtemp()

    def synthetic(name, H, W, pct, seed):  # a map the reference does not ship (non-square), written where the reference can read it
        g = (np.random.RandomState(seed).rand(H, W) < pct / 100.0).astype(np.uint8)
        path = os.path.join(tmp, f"{name}.txt")
        with open(path, "w") as f:
            f.write("\n".join(" ".join(str(int(x)) for x in row) for row in g) + "\n")
        return path

    cases = [("alt_map1", "map1.txt", 5, 50, 120, 31, 200, "homing"), ("alt_map5", "map5.txt", 10, 100, 150, 32, 170, "random"),
             ("alt_map7", "map.txt", 4, 12, 40, 33, 100, "homing"), ("alt_map3", "map3.txt", 6, 30, 90, 34, 120, "homing"),
             ("alt_rect_10x14", synthetic("alt_rect_10x14", 10, 14, 15, 2001), 5, 30, 60, 35, 130, "homing"),
             ("alt_rect_9x7", synthetic("alt_rect_9x7", 9, 7, 10, 2002), 3, 15, 40, 36, 100, "homing")]
    out, meta = {}, []
    for name, map_name, R, P, T, seed, n_steps, policy in cases:
        e = ref_env.Environment(os.path.join(REF, map_name), T, R, P, seed=seed)
        H, W = e.n_rows, e.n_cols
        state = e.reset()
        trk = Tracker(1)
        trk.update(0, state)
        rs = np.random.RandomState(seed + 1000)
        acts_rec = np.zeros((n_steps, R), np.uint8)
        obs6 = np.zeros((n_steps + 1, R, 6, H, W), np.float32)
        g7 = np.zeros((n_steps + 1, 7, H, W), np.float32)
        sh_str = np.zeros((n_steps, R), np.float64)
        sh_int = np.zeros((n_steps, R), np.float64)
        qvec = np.zeros((n_steps + 1, R, 6 + 5 * (R - 1) + 5 * 5 + 1), np.float32)
        qg = np.zeros((n_steps + 1, 6 * 10 + 7 * 20 + 1), np.float32)
        done_rec = np.zeros(n_steps, np.uint8)

        def feats(k, st, pp):
            for i in range(R):
                a = idq.convert_state(st, pp, i)
                assert np.array_equal(a, legacy.convert_state(st, pp, i))
                obs6[k, i] = a
                qvec[k, i] = qmix_helper.generate_vector_features(st, pp, i, T, R - 1, 5)
            g7[k] = legacy.convert_global_state_to_tensor(st, pp, (7, H, W))
            qg[k] = qmix_helper.convert_global_state(st, pp, T, 10, 20)[1]

        feats(0, state, trk.persistent_packages_list[0])
        for s in range(n_steps):
            a = rs.randint(0, 15, size=R)
            if policy == "homing":
                for i, rb in enumerate(e.robots):
                    if rs.rand() < 0.2:
                        continue
                    if rb.carrying:
                        goal, op_at = e.packages[rb.carrying - 1].target, "2"
                    else:
                        wait = [p for p in e.packages if p.status == "waiting"]
                        if not wait:
                            continue
                        goal = min(wait, key=lambda p: abs(p.start[0] - rb.position[0]) + abs(p.start[1] - rb.position[1])).start
                        op_at = "1"
                    dr, dc = goal[0] - rb.position[0], goal[1] - rb.position[1]
                    if dr == 0 and dc == 0:
                        a[i] = encode("S", op_at)
                    elif dr != 0 and (dc == 0 or rs.rand() < 0.5):
                        a[i] = encode("D" if dr > 0 else "U", op_at if (abs(dr) == 1 and dc == 0) else str(rs.randint(0, 3)))
                    else:
                        a[i] = encode("R" if dc > 0 else "L", op_at if (abs(dc) == 1 and dr == 0) else str(rs.randint(0, 3)))
            acts = [decode(x) for x in a]
            pp_prev = {k: dict(v) for k, v in trk.persistent_packages_list[0].items()}
            nxt, r, done, info = e.step(acts)
            acts_rec[s] = a
            done_rec[s] = done
            sh_str[s] = idq.reward_shaping(state, nxt, acts, pp_prev, R)
            sh_int[s] = idq.reward_shaping(state, nxt, [(m, int(o)) for (m, o) in acts], pp_prev, R)
            if done:
                nxt = e.reset()
                trk.clear(0)
            trk.update(0, nxt)
            state = nxt
            feats(s + 1, state, trk.persistent_packages_list[0])
        out.update({f"{name}_actions": acts_rec, f"{name}_obs6": obs6, f"{name}_g7": g7, f"{name}_shape_str": sh_str,
                    f"{name}_shape_int": sh_int, f"{name}_qvec": qvec, f"{name}_qgvec": qg, f"{name}_done": done_rec,
                    f"{name}_grid": grid_of(map_name)})
        meta.append(dict(name=name, map=os.path.basename(map_name), R=R, P=P, T=T, seed=seed, n_steps=n_steps,
                         urgency_cells=int((obs6[:, :, 1] > 0).sum()), frac_cells=int(((g7[:, 6] > 0) & (g7[:, 6] < 1)).sum()),
                         int_branches=sorted(set(np.round(sh_int.reshape(-1), 6).tolist()))))
        print("alt case", meta[-1])
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "kats_alt.npz"), **out)
    print("kats_alt.npz", os.path.getsize(os.path.join(HERE, "kats_alt.npz")) // 1024, "KiB")


def main():
    os.makedirs(HERE, exist_ok=True)
    gen_reset_blocks()
    kats = {}
    kats.update(gen_kat_a())
    kats.update(gen_kat_b())
    kats.update(gen_readme_random())
    np.savez_compressed(os.path.join(HERE, "kats.npz"), **kats)
    print("kats.npz", os.path.getsize(os.path.join(HERE, "kats.npz")) // 1024, "KiB")

    #        name        map        R   P    T    seed steps aseed policy   K_o K_p
    run_traj("map1_c3", "map1.txt", 5, 50, 500, 2025, 700, 7, "random", 4, 5, feat_every=1, big_vec_every=10)
    run_traj("map1_mix", "map1.txt", 5, 50, 120, 42, 400, 1, "mixed", 4, 5, feat_every=1, big_vec_every=7,
             invalid_frac=0.03)
    run_traj("map7_odd", "map.txt", 3, 8, 40, 7, 200, 2, "mixed", 2, 5, feat_every=1, big_vec_every=5)
    run_traj("map7_dense", "map.txt", 20, 30, 60, 3, 150, 3, "random", 19, 5, feat_every=1)
    run_traj("map7_done", "map.txt", 5, 3, 300, 5, 300, 4, "greedy", 4, 5, feat_every=1, big_vec_every=3)
    run_traj("map5_c4", "map5.txt", 10, 100, 200, 11, 300, 5, "mixed", 9, 5, feat_every=2)
    run_traj("map5_c5", "map5.txt", 32, 500, 120, 123, 160, 6, "mixed", 31, 5, feat_every=4)
    run_traj("map1_home", "map1.txt", 5, 12, 150, 77, 500, 9, "homing", 4, 5, feat_every=1, big_vec_every=4)
    run_traj("map7_home", "map.txt", 4, 6, 400, 21, 400, 10, "homing", 3, 5, MR=10, MP=20, feat_every=1)
    run_traj("map3_small", "map3.txt", 7, 20, 100, 6, 250, 8, "mixed", 3, 3, MR=10, MP=20, feat_every=1)


def gen_shapes():
    """Maps the reference does not ship: non-square ones (the featurisers normalise rows by n_rows and columns by n_cols, deadlines
    use n_rows only: SURVEY.md Q11), plane sizes that are not whole float4s, and planes beyond the sliced kernels' 512 cells.  The
    map files are written to a temporary directory (the reference reads maps from files); the fixtures carry the grids."""
    import tempfile

    #        name          H   W  obstacle%  R   P   T  seed steps aseed policy  K_o K_p  kw
    cases = [("rect_12x17", 12, 17, 18, 6, 40, 40, 31, 200, 21, "homing", 5, 5, dict(feat_every=1, big_vec_every=9)),
             ("rect_9x7", 9, 7, 12, 4, 20, 30, 32, 160, 22, "mixed", 3, 5, dict(feat_every=1, big_vec_every=6, invalid_frac=0.02)),
             ("rect_5x24", 5, 24, 10, 5, 25, 35, 33, 150, 23, "homing", 4, 5, dict(MR=10, MP=20, feat_every=1)),
             ("big_40x28", 40, 28, 20, 12, 80, 60, 34, 150, 24, "mixed", 11, 5, dict(feat_every=3)),
             ("big_23x23", 23, 23, 15, 7, 45, 50, 35, 130, 25, "homing", 6, 4, dict(feat_every=2))]
    with tempfile.TemporaryDirectory() as tmp:
        for name, H, W, pct, R, P, T, seed, steps, aseed, policy, K_o, K_p, kw in cases:
            rs = np.random.RandomState(1000 + seed)
            g = (rs.rand(H, W) < pct / 100.0).astype(np.uint8)
            path = os.path.join(tmp, f"{name}.txt")
            with open(path, "w") as f:
                f.write("\n".join(" ".join(str(int(x)) for x in row) for row in g) + "\n")
            run_traj(name, path, R, P, T, seed, steps, aseed, policy, K_o, K_p, **kw)
    # reward constants other than the defaults: the env reward is a float64 sum in the reference's order, so values
    # that are not short decimals make every rounding step visible
    run_traj("map1_rewards", "map1.txt", 5, 14, 120, 36, 420, 26, "homing", 4, 5, feat_every=5, rewards=(-0.0371, 7.3, 0.45))
    run_traj("map7_rewards", "map.txt", 6, 10, 80, 37, 300, 27, "homing", 5, 5, feat_every=5, rewards=(-0.3, 1e-3, 123456.789))


def gen_stale():
    """SURVEY.md Q7 / VERDICT r1 row J1: what MAPPO/trainer.py really feeds the featurisers after the first auto-reset."""
    np.savez_compressed(os.path.join(HERE, "kats_stale.npz"), **gen_kat_b(stale=True))
    run_traj("map1_stale", "map1.txt", 5, 30, 40, 61, 420, 11, "homing", 4, 5, feat_every=1, big_vec_every=7, stale=True)
    run_traj("map7_stale", "map.txt", 4, 12, 25, 62, 300, 12, "mixed", 3, 5, MR=10, MP=20, feat_every=1, big_vec_every=5, stale=True)
    run_traj("map5_stale", "map5.txt", 10, 100, 60, 63, 250, 13, "homing", 9, 5, feat_every=2, invalid_frac=0.02, stale=True)
    # and on a non-square map the reference does not ship (written to a temporary file: the reference reads maps from files)
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        g = (np.random.RandomState(4001).rand(10, 15) < 0.14).astype(np.uint8)
        path = os.path.join(tmp, "rect_10x15.txt")
        with open(path, "w") as f:
            f.write("\n".join(" ".join(str(int(x)) for x in row) for row in g) + "\n")
        run_traj("rect_stale", path, 6, 24, 30, 64, 320, 14, "homing", 5, 5, feat_every=1, big_vec_every=6, stale=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "greedy":
        gen_greedy()
    elif len(sys.argv) > 1 and sys.argv[1] == "stale":
        gen_stale()
    elif len(sys.argv) > 1 and sys.argv[1] == "alt":
        gen_alt()
    elif len(sys.argv) > 1 and sys.argv[1] == "shapes":
        gen_shapes()
    else:
        main()
        gen_greedy()
        gen_alt()
        gen_stale()
        gen_shapes()
