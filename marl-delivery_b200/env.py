"""Drop-in for the reference's `env.Environment` (env.py:18-345) on top of the CUDA batch.

Same constructor, `reset()`, `step(actions)`, `get_state()` and the attributes callers read
(`.grid .n_rows .n_cols .t .total_reward .robots[i].position/.carrying .packages[j].*`,
evaluation.py:50, main.py:86-103).  The dict / attribute views are the slow, lazily materialised
path (one small D2H copy per call) kept for parity and small N; rollouts use DeliveryBatch tensors.
Rendering (env.py:347-456) is out of scope.
"""
from __future__ import annotations

import numpy as np

from .batch import STATUS_NAMES, DeliveryBatch, encode_action
from .maps import load_map


class Robot:
    """Read-only view with the reference's attribute names (env.py:4-7)."""

    def __init__(self, position, carrying):
        self.position = position
        self.carrying = carrying


class Package:
    """Read-only view with the reference's attribute names (env.py:9-16)."""

    def __init__(self, start, start_time, target, deadline, package_id, status):
        self.start = start
        self.start_time = start_time
        self.target = target
        self.deadline = deadline
        self.package_id = package_id
        self.status = status


class StateDict(dict):
    """The state dict of env.py:139-146 plus a back-reference to the environment slot it was read from, so that
    drop-in policies (greedyagent.GreedyAgents) can run on the device-resident state instead of re-parsing lists."""
    __slots__ = ("_view",)


class EnvView:
    """One environment slot of a DeliveryBatch with the reference Environment's API."""

    def __init__(self, batch: DeliveryBatch, index: int, map_file=None):
        self._b = batch
        self._i = int(index)
        self.map_file = map_file
        g = batch.grids[int(batch.env_map[self._i])]
        self.grid = batch._grid_lists[int(batch.env_map[self._i])]  # list of lists like load_map(); never mutated
        self.n_rows, self.n_cols = g.shape
        self.n_robots = batch.n_robots
        self.n_packages = batch.n_packages
        self.max_time_steps = batch.max_time_steps
        self.done = False
        self.state = None
        self._snap = None
        self._snap_ver = -1
        self._ids = [self._i]

    # -- lazily refreshed snapshot ---------------------------------------------------------------
    def _invalidate(self):
        self._snap = None

    def _s(self):
        if self._snap is None or self._snap_ver != self._b._version:
            self._snap = self._b.get_state(self._i)
            self._snap_ver = self._b._version
        return self._snap

    def _seed(self, snap):
        """A snapshot that was read together with other envs' (VectorizedEnv.step / reset: one round trip for all)."""
        self._snap = snap
        self._snap_ver = self._b._version

    @property
    def t(self):
        return self._s()["t"]

    @property
    def total_reward(self):
        return self._s()["total_reward"]

    @property
    def robots(self):
        s = self._s()
        return [Robot((int(r), int(c)), int(k)) for r, c, k in zip(s["robot_row"], s["robot_col"], s["robot_carry"])]

    @property
    def packages(self):
        s = self._s()
        return [Package((int(s["pkg_start_row"][j]), int(s["pkg_start_col"][j])), int(s["pkg_start_time"][j]),
                        (int(s["pkg_target_row"][j]), int(s["pkg_target_col"][j])), int(s["pkg_deadline"][j]), j + 1,
                        STATUS_NAMES[int(s["pkg_status"][j])]) for j in range(self.n_packages)]

    # -- reference API ---------------------------------------------------------------------------
    def get_state(self):
        """env.py:127-147: 1-indexed tuples; 'packages' lists only those spawning at this time step."""
        s = self._s()
        t = s["t"]
        pk = [(j + 1, int(s["pkg_start_row"][j]) + 1, int(s["pkg_start_col"][j]) + 1, int(s["pkg_target_row"][j]) + 1,
               int(s["pkg_target_col"][j]) + 1, t, int(s["pkg_deadline"][j]))
              for j in np.flatnonzero(s["pkg_start_time"] == t).tolist()]
        out = StateDict({"time_step": t, "map": self.grid,
                         "robots": [(r + 1, c + 1, k) for r, c, k in
                                    zip(s["robot_row"].tolist(), s["robot_col"].tolist(), s["robot_carry"].tolist())],
                         "packages": pk})
        out._view = self
        return out

    def reset(self):
        self._b.reset(self._ids)
        self.done = False
        self.state = None
        self._invalidate()
        return self.get_state()

    def step(self, actions):
        if len(actions) != self.n_robots:
            raise ValueError("The number of actions must match the number of robots.")
        codes = np.array([[encode_action(m, p) for (m, p) in actions]], dtype=np.uint8)
        reward, _, done, rows = self._b.step_host(codes, self._ids)  # one C call: H2D, kernel, snapshot, D2H, one sync
        r = float(reward[0])
        d = bool(done[0])
        self._seed(self._b._rows_to_states(rows)[0])
        infos = {}
        if d:
            infos["total_reward"] = self.total_reward
            infos["total_time_steps"] = self.t
        return self.get_state(), r, d, infos

    def check_terminate(self):
        s = self._s()
        return s["t"] == self.max_time_steps or bool(np.all(s["pkg_status"] == 3))

    def is_free_cell(self, position):
        r, c = position
        if r < 0 or r >= self.n_rows or c < 0 or c >= self.n_cols:
            return False
        return self.grid[r][c] == 0

    valid_position = is_free_cell

    def render(self):  # UI is out of scope
        raise NotImplementedError("rendering is not part of the B200 hot path")

    render_pygame = render


class Environment(EnvView):
    """`Environment(map_file, max_time_steps, n_robots, n_packages, move_cost, delivery_reward,
    delay_reward, seed)` -- env.py:20-43.  Like the reference, the constructor already performs one
    reset (it consumes the first block of the seeded stream)."""

    def __init__(self, map_file, max_time_steps=100, n_robots=5, n_packages=20, move_cost=-0.01,
                 delivery_reward=10., delay_reward=1., seed=2025, device=None):
        grid = load_map(map_file)
        batch = DeliveryBatch([grid], 1, n_robots, n_packages, max_time_steps, move_cost, delivery_reward,
                              delay_reward, base_seed=seed, device=device)
        super().__init__(batch, 0, map_file)
        self.move_cost = move_cost
        self.delivery_reward = delivery_reward
        self.delay_reward = delay_reward
