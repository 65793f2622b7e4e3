"""CPU ORACLE (test infrastructure, NOT product code): restatement of the reference's GreedyAgents
policy (greedyagent.py:6-174) on the flat-state OracleEnv.

The restatement keeps the agent's bookkeeping literal, because the reference's results depend on it:
  * `init_agents` stores the packages visible at t = 0 and the FIRST `get_actions` call stores the same
    state's packages again (greedyagent.py:52-62,121-129: `is_init` is only set inside get_actions), so
    the agent's package list starts with every t = 0 package twice;
  * the list is indexed by `package_id - 1` everywhere (:143,147,165-168), which after the duplication
    points at the wrong entry for every id beyond the t = 0 packages, and the `packages_free` flag that is
    cleared is the one at `id - 1`, not the entry that was found;
  * a robot's target is only refreshed from the env when it was already carrying (:107-113).
Pinned by tests/test_greedy_oracle.py against the README Greedy row and per-step action traces
recorded from the unmodified reference (tests/golden/kats_greedy.npz).
"""
from __future__ import annotations

from collections import deque

MOVE_CODE = {"D": 0, "L": 1, "R": 2, "S": 3, "U": 4}


def bfs_from(grid, goal):
    """greedyagent.py:6-29: 4-neighbour BFS distances from `goal` over free cells."""
    H, W = grid.shape
    d = {goal: 0}
    q = deque([goal])
    while q:
        cur = q.popleft()
        for dx, dy in ((-1, 0), (1, 0), (0, -1), (0, 1)):
            nx = (cur[0] + dx, cur[1] + dy)
            if nx[0] < 0 or nx[0] >= H or nx[1] < 0 or nx[1] >= W:
                continue
            if nx not in d and grid[nx[0], nx[1]] == 0:
                d[nx] = d[cur] + 1
                q.append(nx)
    return d


class GreedyOracle:
    def __init__(self, env):
        """== GreedyAgents() + init_agents(state) on a freshly reset env (t = 0)."""
        self.env = env
        W = env.W
        self.R = env.R
        self.robots = [(c // W, c % W, 0) for c in env.rcell]
        self.target = [None] * env.R  # 'free'
        self.pkgs = self._spawned_now()  # (id, sr, sc, tr, tc)
        self.free = [True] * len(self.pkgs)
        self._bfs = {}

    def _spawned_now(self):
        e, W = self.env, self.env.W
        return [(j + 1, e.p_start[j] // W, e.p_start[j] % W, e.p_target[j] // W, e.p_target[j] % W)
                for j in range(e.P) if e.p_t0[j] == e.t]

    def _step_towards(self, start, goal):
        """run_bfs (greedyagent.py:6-41): first of U, D, L, R that lowers the BFS distance to goal."""
        if goal not in self._bfs:
            self._bfs[goal] = bfs_from(self.env.grid, goal)
        d = self._bfs[goal]
        if start not in d:
            return "S", 100000
        for mv, (dx, dy) in zip("UDLR", ((-1, 0), (1, 0), (0, -1), (0, 1))):
            nx = (start[0] + dx, start[1] + dy)
            if nx in d and d[nx] == d[start] - 1:
                return mv, d[nx]
        return "S", d[start]

    def _move_to(self, i, idx, phase):
        """update_move_to_target (greedyagent.py:64-100); idx indexes the agent's own list."""
        pk = self.pkgs[idx]
        goal = (pk[1], pk[2]) if phase == "start" else (pk[3], pk[4])
        pos = (self.robots[i][0], self.robots[i][1])
        dist = abs(goal[0] - pos[0]) + abs(goal[1] - pos[1])
        act = 0
        if dist >= 1:
            mv, dist = self._step_towards(pos, goal)
            if dist == 0:
                act = 1 if phase == "start" else 2
        else:
            mv = "S"
            act = 1 if phase == "start" else 2
        return MOVE_CODE[mv] + 5 * act

    def get_actions(self):
        """get_actions (greedyagent.py:119-174) for the env's current state -> action codes."""
        e, W = self.env, self.env.W
        # update_inner_state (:102-117)
        for i in range(self.R):
            prev = self.robots[i]
            self.robots[i] = (e.rcell[i] // W, e.rcell[i] % W, e.carry[i])
            if prev[2] != 0:
                self.target[i] = None if self.robots[i][2] == 0 else self.robots[i][2]
        new = self._spawned_now()
        self.pkgs += new
        self.free += [True] * len(new)
        out = []
        for i in range(self.R):
            if self.target[i] is not None:
                tid = self.target[i]
                out.append(self._move_to(i, tid - 1, "target" if self.robots[i][2] != 0 else "start"))
                continue
            best, best_d = None, 1000000
            for j, pk in enumerate(self.pkgs):
                if not self.free[j]:
                    continue
                d = abs(pk[1] - self.robots[i][0]) + abs(pk[2] - self.robots[i][1])
                if d < best_d:
                    best_d, best = d, pk[0]
            if best is None:
                out.append(MOVE_CODE["S"])
                continue
            self.free[best - 1] = False
            self.target[i] = best
            out.append(self._move_to(i, best - 1, "start"))
        return out
