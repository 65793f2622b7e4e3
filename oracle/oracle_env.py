"""CPU ORACLE (test infrastructure, NOT product code) -- pure Python/numpy restatement.

A from-scratch restatement of the reference's environment transition, reward, package tracker,
MAPPO featurisation and shaped reward, written against flat integer state instead of the
reference's Robot/Package objects.  Every function cites the reference lines it follows
(paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path
(marl-delivery_b200/) never does and has no CPU fallback.

Pinned by tests/test_oracle_golden.py against every fixture in tests/golden/ (generated from the
unmodified reference by tests/golden/make_golden.py): reset blocks, 11 trajectories (positions,
carried ids, statuses, f64 rewards, done flags, all four feature tensors, shaped rewards under the
MAPPO and the QMIX constants), KAT-A, KAT-B, the README Random row and the notebook dummy-shaping
known answers.

Action codes: a in 0..14 -> move 'DLRSU'[a % 5], package op a // 5;
a in 15..17 -> an illegal move string with package op a - 15.
Package status codes: 0 'None', 1 'waiting', 2 'in_transit', 3 'delivered'.
"""
from __future__ import annotations

import numpy as np

NONE, WAITING, IN_TRANSIT, DELIVERED = 0, 1, 2, 3
MV_D, MV_L, MV_R, MV_S, MV_U, MV_BAD = 0, 1, 2, 3, 4, 5
_DELTA = {MV_D: (1, 0), MV_L: (0, -1), MV_R: (0, 1), MV_S: (0, 0), MV_U: (-1, 0), MV_BAD: (0, 0)}

# shaping constants in the order: pickup, on-time, late, closer, wasted pick, wasted drop,
# collision/stuck, idle, away
MAPPO_SHAPING = (5, 200, 20, 0.02, 0, 0, -0.05, -0.05, -0.01)  # MAPPO/helper.py:271-279
QMIX_SHAPING = (0.5, 2.0, 0.2, 0.02, -0.1, -0.1, -0.05, -0.02, -0.02)  # QMIX/helper.py:270-278


def parse_map_text(text: str) -> np.ndarray:
    """env.py:45-57 -- one row per line, cells separated by single spaces."""
    rows = [[int(tok) for tok in line.strip().split(" ")] for line in text.splitlines() if line.strip() != ""]
    return np.array(rows, dtype=np.uint8)


def split_action(a: int):
    a = int(a)
    if a < 15:
        return a % 5, a // 5
    return MV_BAD, a - 15


class OracleEnv:
    """Flat-state restatement of env.py:18-345 (no rendering)."""

    def __init__(self, grid, max_time_steps=100, n_robots=5, n_packages=20, move_cost=-0.01,
                 delivery_reward=10.0, delay_reward=1.0, seed=2025):
        self.grid = np.asarray(grid, dtype=np.uint8)
        self.H, self.W = self.grid.shape
        self.T, self.R, self.P = int(max_time_steps), int(n_robots), int(n_packages)
        self.move_cost, self.delivery_reward, self.delay_reward = move_cost, delivery_reward, delay_reward
        self.rng = np.random.RandomState(seed)  # env.py:40
        self.free_cells = [int(c) for c in np.flatnonzero(self.grid.reshape(-1) == 0)]  # row-major, env.py:155
        self.reset()  # env.py:41 -- the constructor consumes one reset block

    # ------------------------------------------------------------------ reset, env.py:81-125
    def reset(self):
        self.t = 0
        self.total_reward = 0
        rng = self.rng
        # robots: draw an index into the SHRINKING row-major free list (env.py:96-100,161-170)
        avail = list(self.free_cells)
        self.rcell = []
        for _ in range(self.R):
            k = rng.randint(0, len(avail))
            self.rcell.append(avail.pop(k))
        self.carry = [0] * self.R
        nfree = len(self.free_cells)
        N = self.H
        drawn = []
        for i in range(self.P):
            s = self.free_cells[rng.randint(0, nfree)]  # env.py:106,150-158
            while True:  # env.py:107-110
                g = self.free_cells[rng.randint(0, nfree)]
                if g != s:
                    break
            slack = 10 + rng.randint(N / 2, 3 * N)  # env.py:112 (float low bound, numpy truncates)
            st = 0 if i <= min(self.R, 20) else rng.randint(1, self.T)  # env.py:113-116
            drawn.append((st, s, g, st + slack))
        drawn.sort(key=lambda rec: rec[0])  # stable, env.py:119
        self.p_start = [d[1] for d in drawn]
        self.p_target = [d[2] for d in drawn]
        self.p_t0 = [d[0] for d in drawn]
        self.p_deadline = [d[3] for d in drawn]
        self.p_status = [NONE] * self.P
        self._spawn()
        return self

    def _spawn(self):
        """The status side effect of get_state, env.py:134-137."""
        for j in range(self.P):
            if self.p_t0[j] == self.t:
                self.p_status[j] = WAITING

    # ------------------------------------------------------------------ step, env.py:173-306
    def step(self, actions):
        if len(actions) != self.R:  # env.py:182-183
            raise ValueError("The number of actions must match the number of robots.")
        H, W, R = self.H, self.W, self.R
        mv = [0] * R
        op = [0] * R
        for i, a in enumerate(actions):
            mv[i], op[i] = split_action(a)
        # proposals, env.py:193-201 (+ compute_new_position :318-334, valid_position :336-345)
        want = []
        holder = {}  # cell -> robot standing on it when the step starts
        for i in range(R):
            r, c = divmod(self.rcell[i], W)
            dr, dc = _DELTA[mv[i]]
            nr, nc = r + dr, c + dc
            if not (0 <= nr < H and 0 <= nc < W) or self.grid[nr, nc] == 1:
                nr, nc = r, c
            want.append(nr * W + nc)
            holder[self.rcell[i]] = i
        # restart-scan resolution, env.py:203-250
        settled = [False] * R
        final = [None] * R
        claimed = set()
        progress = True
        while progress:
            progress = False
            for i in range(R):
                if settled[i]:
                    continue
                tgt = want[i]
                j = holder.get(tgt)
                if j is not None and j != i and not settled[j]:
                    continue  # the cell's holder has not been decided yet
                if tgt in claimed:
                    tgt = self.rcell[i]
                claimed.add(tgt)
                final[i] = tgt
                settled[i] = True
                progress = True
                break
        for i in range(R):
            if not settled[i]:
                final[i] = self.rcell[i]
        # movement cost + commit, env.py:253-257  (reward starts as int 0, float adds in robot order)
        rew = 0
        for i in range(R):
            if mv[i] in (MV_L, MV_R, MV_U, MV_D) and final[i] != self.rcell[i]:
                rew += self.move_cost
            self.rcell[i] = final[i]
        # package ops in robot order, env.py:260-292
        for i in range(R):
            if op[i] == 1:
                if self.carry[i] == 0:
                    for j in range(self.P):
                        if self.p_status[j] == WAITING and self.p_start[j] == self.rcell[i] and self.p_t0[j] <= self.t:
                            self.carry[i] = j + 1
                            self.p_status[j] = IN_TRANSIT
                            break
            elif op[i] == 2:
                if self.carry[i] != 0:
                    j = self.carry[i] - 1
                    if self.rcell[i] == self.p_target[j]:
                        self.p_status[j] = DELIVERED
                        rew += self.delivery_reward if self.t <= self.p_deadline[j] else self.delay_reward
                        self.carry[i] = 0
        self.t += 1  # env.py:295
        self.total_reward += rew
        done = (self.t == self.T) or all(s == DELIVERED for s in self.p_status)  # env.py:308-316, not latched
        self._spawn()  # get_state at env.py:306
        return rew, done

    # ------------------------------------------------------------------ views
    def robots_rc(self):
        return [(c // self.W, c % self.W, k) for c, k in zip(self.rcell, self.carry)]

    def snapshot(self):
        """Immutable copy of everything the featurisers / shaping read."""
        return dict(t=self.t, rcell=list(self.rcell), carry=list(self.carry), status=list(self.p_status))

    def state_dict(self):
        """The reference's state dict (env.py:139-146): 1-indexed tuples, newly spawned packages only."""
        W = self.W
        pk = [(j + 1, self.p_start[j] // W + 1, self.p_start[j] % W + 1, self.p_target[j] // W + 1,
               self.p_target[j] % W + 1, self.p_t0[j], self.p_deadline[j])
              for j in range(self.P) if self.p_t0[j] == self.t]
        return {"time_step": self.t, "map": self.grid.tolist(),
                "robots": [(c // W + 1, c % W + 1, k) for c, k in zip(self.rcell, self.carry)], "packages": pk}


# ---------------------------------------------------------------------- featurisers
def _active(env, status, j, t):
    """'waiting and start_time <= t' or 'in_transit' (MAPPO/helper.py:53,194,224)."""
    return (status[j] == WAITING and env.p_t0[j] <= t) or status[j] == IN_TRANSIT


def observation(env: OracleEnv, snap, i):
    """convert_observation, MAPPO/helper.py:6-66.  The package dict is the env truth restricted to
    waiting / in-transit packages (== the trainers' tracker)."""
    H, W = env.H, env.W
    t, rcell, carry, status = snap["t"], snap["rcell"], snap["carry"], snap["status"]
    o = np.zeros((6, H * W), np.float32)
    o[0] = env.grid.reshape(-1)
    o[1, rcell[i]] = 1.0
    for k in range(env.R):
        if k != i:
            o[2, rcell[k]] = 1.0
    for j in range(env.P):
        if status[j] == WAITING and env.p_t0[j] <= t:
            o[3, env.p_start[j]] = 1.0
        if _active(env, status, j, t):
            o[4, env.p_target[j]] = 1.0
    if carry[i] != 0 and status[carry[i] - 1] == IN_TRANSIT:
        o[5, env.p_target[carry[i] - 1]] = 1.0
    return o.reshape(6, H, W)


def vector_features(env: OracleEnv, snap, i, K_o=100, K_p=100):
    """generate_vector_features, MAPPO/helper.py:68-165 (float64 quotients, float64 sort keys,
    Python's stable sort, one final cast to float32)."""
    H, W, T = env.H, env.W, env.T
    t, rcell, carry, status = snap["t"], snap["rcell"], snap["carry"], snap["status"]
    r0, c0 = divmod(rcell[i], W)
    f = [r0 / H, c0 / W, 1.0 if carry[i] else 0.0]
    if carry[i] and status[carry[i] - 1] == IN_TRANSIT:
        j = carry[i] - 1
        tr, tc = divmod(env.p_target[j], W)
        f += [(tr - r0) / H, (tc - c0) / W, (max(0, env.p_deadline[j] - t) / T) if T > 0 else 0.0]
    else:
        f += [0.0, 0.0, 0.0]
    rows = []
    for k in range(env.R):
        if k == i:
            continue
        r, c = divmod(rcell[k], W)
        row = [(r - r0) / H, (c - c0) / W, 1.0 if carry[k] else 0.0]
        if carry[k] and status[carry[k] - 1] == IN_TRANSIT:
            tr, tc = divmod(env.p_target[carry[k] - 1], W)
            row += [(tr - r) / H, (tc - c) / W]
        else:
            row += [0.0, 0.0]
        rows.append(row)
    rows.sort(key=lambda x: x[0] ** 2 + x[1] ** 2)  # MAPPO/helper.py:139
    for k in range(K_o):
        f += rows[k] if k < len(rows) else [0.0] * 5
    prow = []
    for j in range(env.P):
        if status[j] == WAITING and env.p_t0[j] <= t:
            sr, sc = divmod(env.p_start[j], W)
            tr, tc = divmod(env.p_target[j], W)
            prow.append([(sr - r0) / H, (sc - c0) / W, (tr - r0) / H, (tc - c0) / W,
                         (max(0, env.p_deadline[j] - t) / T) if T > 0 else 0.0])
    prow.sort(key=lambda x: (x[4], x[0] ** 2 + x[1] ** 2))  # MAPPO/helper.py:158
    for k in range(K_p):
        f += prow[k] if k < len(prow) else [0.0] * 5
    f.append(t / T if T > 0 else 0.0)
    return np.array(f, dtype=np.float32)


def global_state(env: OracleEnv, snap, MR=100, MP=100):
    """convert_global_state, MAPPO/helper.py:167-255."""
    H, W, T = env.H, env.W, env.T
    t, rcell, carry, status = snap["t"], snap["rcell"], snap["carry"], snap["status"]
    g = np.zeros((4, H * W), np.float32)
    g[0] = env.grid.reshape(-1)
    for c in rcell:
        g[1, c] = 1.0
    for j in range(env.P):
        if status[j] == WAITING and env.p_t0[j] <= t:
            g[2, env.p_start[j]] = 1.0
        if _active(env, status, j, t):
            g[3, env.p_target[j]] = 1.0
    v = []
    for k in range(MR):
        if k < env.R:
            r, c = divmod(rcell[k], W)
            v += [r / H, c / W, 1.0 if carry[k] else 0.0]
            if carry[k] and status[carry[k] - 1] == IN_TRANSIT:
                j = carry[k] - 1
                tr, tc = divmod(env.p_target[j], W)
                v += [tr / H, tc / W, (max(0, env.p_deadline[j] - t) / T) if T > 0 else 0.0]
            else:
                v += [0.0, 0.0, 0.0]
        else:
            v += [0.0] * 6
    act = [j for j in range(env.P) if _active(env, status, j, t)]  # ascending id
    for k in range(MP):
        if k < len(act):
            j = act[k]
            sr, sc = divmod(env.p_start[j], W)
            tr, tc = divmod(env.p_target[j], W)
            wait = status[j] == WAITING
            v += [sr / H, sc / W] if wait else [0.0, 0.0]
            v += [tr / H, tc / W, (max(0, env.p_deadline[j] - t) / T) if T > 0 else 0.0, 0.0 if wait else 1.0]
            cn = -1.0
            if not wait:
                for ridx in range(env.R):
                    if carry[ridx] == j + 1:
                        cn = ridx / (MR - 1) if MR > 1 else 0.0
                        break
            v.append(cn)
        else:
            v += [0.0] * 7
    v.append(t / T if T > 0 else 0.0)
    return g.reshape(4, H, W), np.array(v, dtype=np.float32)


def bfs_all_pairs(grid):
    """greedyagent.py:6-29 (4-neighbour BFS over free cells) from every free cell -> uint16 [H*W, H*W],
    0xFFFF = unreachable / obstacle.  Used by the greedy-policy checks and the opt-in BFS shaping mode."""
    grid = np.asarray(grid)
    H, W = grid.shape
    out = np.full((H * W, H * W), 0xFFFF, np.uint16)
    for g in range(H * W):
        if grid[g // W, g % W]:
            continue
        d = out[g]
        d[g] = 0
        q = [g]
        for cur in q:
            r, c = divmod(cur, W)
            for nr, nc in ((r - 1, c), (r + 1, c), (r, c - 1), (r, c + 1)):
                if 0 <= nr < H and 0 <= nc < W and grid[nr, nc] == 0 and d[nr * W + nc] == 0xFFFF:
                    d[nr * W + nc] = d[cur] + 1
                    q.append(nr * W + nc)
    return out


def shaped_reward(env: OracleEnv, reward, prev, cur, actions, consts=MAPPO_SHAPING, bfs=None):
    """compute_shaped_rewards, MAPPO/helper.py:257-369: float32 per-agent accumulators in the
    reference's order of terms, numpy float32 `.sum()`, then `python float + np.float32` which numpy>=2
    evaluates in float32.  `prev` / `cur` are snapshots before / after the step (the
    after-snapshot is taken BEFORE any auto-reset, Q10).

    `bfs` (an all-pairs table from bfs_all_pairs) switches the distance function from the reference's
    Manhattan distance (MAPPO/helper.py:284-285) to shortest-path length: the OPT-IN mode of the build
    (BASELINE.json north star vs SURVEY.md Q9).  That mode has no reference behaviour to pin against --
    PARITY UNPINNED for bfs is not None; it is checked for equality with the Manhattan mode on maps
    without inner walls and against this restatement elsewhere."""
    (c_pick, c_ontime, c_late, c_closer, c_wpick, c_wdrop, c_stuck, c_idle, c_away) = consts
    W = env.W
    pstat, pt = prev["status"], prev["t"]
    acc = np.zeros(env.R, dtype=np.float32)

    def md(a, b):
        if bfs is not None:
            return int(bfs[b, a])
        return abs(a // W - b // W) + abs(a % W - b % W)

    def tracked(pid):  # "pid in persistent_packages_at_prev_state"
        return pid != 0 and pstat[pid - 1] in (WAITING, IN_TRANSIT)

    waiting_prev = [j for j in range(env.P) if pstat[j] == WAITING and env.p_t0[j] <= pt]
    for i in range(env.R):
        p0, p1 = prev["rcell"][i], cur["rcell"][i]
        k0, k1 = prev["carry"][i], cur["carry"][i]
        mv, op = split_action(actions[i])
        if k0 == 0 and k1 != 0:
            acc[i] += c_pick
        elif k0 != 0 and k1 == 0:
            if tracked(k0) and p1 == env.p_target[k0 - 1]:
                acc[i] += c_ontime if cur["t"] <= env.p_deadline[k0 - 1] else c_late
        if op == 1:
            if k0 != 0:
                acc[i] += c_wpick
            elif k1 == 0:
                if not any(env.p_start[j] == p1 for j in waiting_prev):
                    acc[i] += c_wpick
        elif op == 2:
            if k0 == 0:
                acc[i] += c_wdrop
            elif k1 != 0:
                if tracked(k0) and p1 != env.p_target[k0 - 1]:
                    acc[i] += c_wdrop
        moved = p0 != p1
        if mv != MV_S and not moved:
            acc[i] += c_stuck
        goal = None
        if tracked(k0):
            goal = env.p_target[k0 - 1]
        elif waiting_prev:
            goal = env.p_start[min(waiting_prev, key=lambda j: md(p0, env.p_start[j]))]
        if goal is not None and moved:
            d0, d1 = md(p0, goal), md(p1, goal)
            if d1 < d0:
                acc[i] += c_closer
            elif d1 > d0:
                acc[i] += c_away
        if (not moved) and mv == MV_S and k0 == 0:
            if any(md(p0, env.p_start[j]) <= 3 for j in waiting_prev):
                acc[i] += c_idle
    return reward + acc.sum()


def shaped_reward_raw(W, reward, prev_robots, cur_robots, prev_t, cur_t, actions, pkgs, consts=MAPPO_SHAPING):
    """Same rule on explicitly given data (for the notebook 'Dummy test' known answers, whose inputs
    are not an env transition).  prev_robots/cur_robots: [(r, c, carry)] 0-indexed; pkgs: {id:
    (start_cell, target_cell, start_time, deadline, status_code)}."""
    P = max(pkgs) if pkgs else 0

    class _E:  # minimal stand-in exposing what shaped_reward reads
        pass

    e = _E()
    e.W, e.R, e.P = W, len(prev_robots), P
    e.p_start = [pkgs.get(j + 1, (0, 0, 0, 0, NONE))[0] for j in range(P)]
    e.p_target = [pkgs.get(j + 1, (0, 0, 0, 0, NONE))[1] for j in range(P)]
    e.p_t0 = [pkgs.get(j + 1, (0, 0, 0, 0, NONE))[2] for j in range(P)]
    e.p_deadline = [pkgs.get(j + 1, (0, 0, 0, 0, NONE))[3] for j in range(P)]
    status = [pkgs.get(j + 1, (0, 0, 0, 0, NONE))[4] for j in range(P)]
    prev = dict(t=prev_t, rcell=[r * W + c for r, c, _ in prev_robots], carry=[k for _, _, k in prev_robots], status=status)
    cur = dict(t=cur_t, rcell=[r * W + c for r, c, _ in cur_robots], carry=[k for _, _, k in cur_robots], status=status)
    return shaped_reward(e, reward, prev, cur, actions, consts)


class OracleVecEnv:
    """MAPPO/env_vectorized.py:1-24 / QMIX/env_vectorized.py:2-48: N sequential envs, seed base+idx."""

    def __init__(self, num_envs, grid, seed=None, **kw):
        self.envs = [OracleEnv(grid, seed=(None if seed is None else seed + i), **kw) for i in range(num_envs)]
        self.num_envs = num_envs

    def reset(self, indices=None):
        for i in (range(self.num_envs) if indices is None else indices):
            self.envs[i].reset()

    def step(self, actions, indices=None):
        idx = range(self.num_envs) if indices is None else indices
        out = [self.envs[i].step(a) for i, a in zip(idx, actions)]
        return [o[0] for o in out], [o[1] for o in out]


def mappo_rollout_step(venv: OracleVecEnv, actions, K_o, K_p, MR=100, MP=100, consts=MAPPO_SHAPING):
    """One iteration of MAPPO/trainer.py:172-286 for every env, minus the networks: step, shaped
    reward from (before, after-before-reset), auto-reset, then the four feature tensors of the state
    the next action is taken from.  convert_global_state is evaluated twice per env exactly like the
    trainer does (:237-248), which is what the CPU baseline is charged for."""
    N = venv.num_envs
    R = venv.envs[0].R
    shaped = np.zeros(N, np.float32)
    dones = np.zeros(N, bool)
    obs, vec, gst, gvec = [], [], [], []
    for n, e in enumerate(venv.envs):
        before = e.snapshot()
        r, d = e.step(actions[n])
        after = e.snapshot()
        shaped[n] = shaped_reward(e, r, before, after, actions[n], consts)
        dones[n] = d
        if d:
            e.reset()
        s = e.snapshot()
        gst.append(global_state(e, s, MR, MP)[0])
        gvec.append(global_state(e, s, MR, MP)[1])
        obs.append(np.stack([observation(e, s, i) for i in range(R)]))
        vec.append(np.stack([vector_features(e, s, i, K_o, K_p) for i in range(R)]))
    return shaped, dones, np.stack(obs), np.stack(vec), np.stack(gst), np.stack(gvec)


def gae_reference(rewards, values, dones, last_value, gamma, gae_lambda):
    """The GAE loop of MAPPO/trainer.py:288-301 restated with torch CPU float32 tensors (same expressions,
    same Python-float scalars) -> (advantages, returns).  The plain-torch reference of the CUDA GAE kernel."""
    import torch

    rewards, values = torch.as_tensor(rewards, dtype=torch.float32), torch.as_tensor(values, dtype=torch.float32)
    dones = torch.as_tensor(dones).bool()
    T = rewards.shape[0]
    advantages = torch.zeros_like(rewards)
    last_gae_lambda = 0
    next_value = torch.as_tensor(last_value, dtype=torch.float32)
    for t in reversed(range(T)):
        next_non_terminal = 1.0 - dones[t].float()
        next_values_step = next_value if t == T - 1 else values[t + 1]
        delta = rewards[t] + gamma * next_values_step * next_non_terminal - values[t]
        advantages[t] = last_gae_lambda = delta + gamma * gae_lambda * next_non_terminal * last_gae_lambda
    return advantages, advantages + values
