"""Drop-ins for the MAPPO helper's converters and shaping (MAPPO/helper.py:6-369; QMIX/helper.py is
the same code with other constants) -- same names, same arguments, same np.float32 results, computed
by the CUDA kernels.

These per-call functions exist so that reference code keeps working unchanged; each call injects the
given (state dict, tracker dict) into a cached 1-env scratch batch and runs the batched kernel on it
(one small H2D + D2H per call).  Rollouts should use `VectorizedEnv.rollout_step` /
`DeliveryBatch.featurize`, which produce the same tensors for all envs in one launch.

The tracker dict may be any dict the reference trainers build: keyed by ascending package id (a tracker that is
cleared on reset) it is injected as env truth; in any other insertion order (the MAPPO trainer's never-cleared dict,
SURVEY.md Q7) it is injected, in dict order, into the scratch batch's tracker table (MDX_CFG_STALE_TRACKER).
"""
from __future__ import annotations

import numpy as np

from .batch import MAPPO_SHAPING, QMIX_SHAPING, DeliveryBatch, encode_action  # noqa: F401

_SCRATCH = {}
_STATUS_CODE = {"None": 0, "waiting": 1, "in_transit": 2, "delivered": 3}


def _scratch(grid, R, P, T, K_o, K_p, MR, MP, shaping, stale=False):
    key = (grid.tobytes(), grid.shape, R, P, T, K_o, K_p, MR, MP, tuple(shaping), stale)
    b = _SCRATCH.get(key)
    if b is None:
        if len(_SCRATCH) > 16:
            _SCRATCH.pop(next(iter(_SCRATCH))).close()
        b = DeliveryBatch([grid], 1, R, P, T, k_others=K_o, k_packages=K_p, max_robots_in_state=MR,
                          max_packages_in_state=MP, shaping=shaping, base_seed=0, blank=True, lax_robots=True, stale_tracker=stale)
        _SCRATCH[key] = b
    return b


def _pad_packages(n):
    p = 8
    while p < n:
        p *= 2
    return p


def _ascending(tracker):
    prev = 0
    for pid in tracker:
        if int(pid) <= prev:
            return False
        prev = int(pid)
    return True


def _inject(b: DeliveryBatch, state, tracker):
    """(state dict, tracker dict) -> env 0 of the scratch batch as ONE packed row (one host -> device transfer).  A dict keyed
    by ascending package id (every tracker that is cleared on reset) goes in as the env truth; any other dict -- the MAPPO
    trainer's never-cleared one iterates in insertion order -- goes into the scratch batch's tracker table in dict order."""
    W, P = b.cols_out, b.n_packages
    row = np.zeros(1, dtype=b._row_dtype)
    row["t"] = int(state["time_step"])
    row["robots"][0] = [((int(r) - 1) * W + int(c) - 1) | (int(k) << 16) for (r, c, k) in state["robots"]]
    sched, status = row["sched"][0], row["status"][0]
    sched[2, P:] = 0xffff
    entries = []
    for pid, p in tracker.items():
        pid = int(pid)
        (sr, sc), (tr, tc) = p["start_pos"], p["target_pos"]
        if min(sr, sc, tr, tc) < 0 or max(sr, tr) >= b.rows_out or max(sc, tc) >= W:
            raise ValueError(f"package {pid} outside the map")
        if b.stale_tracker:
            entries.append((pid, sr * W + sc, tr * W + tc, int(p["start_time"]), int(p["deadline"]), _STATUS_CODE[p["status"]]))
        else:
            j = pid - 1
            sched[0, j], sched[1, j], sched[2, j], sched[3, j] = sr * W + sc, tr * W + tc, int(p["start_time"]), int(p["deadline"])
            status[j] = _STATUS_CODE[p["status"]]
    # a carried id the tracker does not know stays 'None' -> "not in dict" branch of the reference
    b.set_rows([0], row)
    if b.stale_tracker:
        b.set_tracker(0, entries)


def _prepare(state, tracker, T, K_o, K_p, MR, MP, shaping=MAPPO_SHAPING, min_pkg=0):
    grid = np.ascontiguousarray(np.asarray(state["map"], dtype=np.uint8))
    R = len(state["robots"])
    ids = [int(k) for k in tracker] + [int(r[2]) for r in state["robots"]] + [min_pkg]
    P = _pad_packages(max(ids + [1]))
    b = _scratch(grid, R, P, max(int(T), 0), K_o, K_p, MR, MP, shaping, stale=not _ascending(tracker))
    _inject(b, state, tracker)
    return b, grid.shape


def convert_observation(env_state_dict, persistent_packages_for_env, current_robot_idx):
    """MAPPO/helper.py:6-66 -> (6, n_rows, n_cols) float32."""
    b, (H, W) = _prepare(env_state_dict, persistent_packages_for_env, 1, 0, 0, 1, 1)
    if not (0 <= current_robot_idx < len(env_state_dict["robots"])):
        obs = np.zeros((6, H, W), np.float32)
        obs[0] = np.asarray(env_state_dict["map"], dtype=np.float32)
        return obs
    o, _, _, _ = b.featurize_host(vec=False, gstate=False, gvec=False)
    return o[0, current_robot_idx]


def generate_vector_features(env_state_dict, persistent_packages_for_env, current_robot_idx, max_time_steps,
                             max_other_robots_to_observe=100, max_packages_to_observe=100):
    """MAPPO/helper.py:68-165 -> (6 + 5*K_o + 5*K_p + 1,) float32."""
    K_o, K_p = int(max_other_robots_to_observe), int(max_packages_to_observe)
    if not (0 <= current_robot_idx < len(env_state_dict["robots"])):
        return np.zeros(6 + 5 * K_o + 5 * K_p + 1, dtype=np.float32)
    b, _ = _prepare(env_state_dict, persistent_packages_for_env, max_time_steps, K_o, K_p, 1, 1)
    _, v, _, _ = b.featurize_host(obs=False, gstate=False, gvec=False)
    return v[0, current_robot_idx]


def convert_global_state(env_state_dict, persistent_packages_for_env, max_time_steps, max_robots_in_state=100,
                         max_packages_in_state=100):
    """MAPPO/helper.py:167-255 -> ((4, n_rows, n_cols) float32, (6*MR + 7*MP + 1,) float32)."""
    b, _ = _prepare(env_state_dict, persistent_packages_for_env, max_time_steps, 0, 0, int(max_robots_in_state),
                    int(max_packages_in_state))
    _, _, g, gv = b.featurize_host(obs=False, vec=False)
    return g[0], gv[0]


def compute_shaped_rewards(global_reward, prev_env_state_dict, current_env_state_dict, actions_taken_for_all_agents,
                           persistent_packages_at_prev_state, num_agents, shaping=MAPPO_SHAPING, n_cols=None):
    """MAPPO/helper.py:257-369 -> np.float32.  The reference needs no map for this function; pass
    `n_cols` when the state dicts carry no 'map' (as in the notebook 'Dummy test')."""
    prev = dict(prev_env_state_dict)
    prev["robots"] = list(prev["robots"])[:num_agents]
    cur_robots = list(current_env_state_dict["robots"])[:num_agents]
    if "map" not in prev:
        span = [max(int(x[0]) for x in prev["robots"] + cur_robots), max(int(x[1]) for x in prev["robots"] + cur_robots)]
        for p in persistent_packages_at_prev_state.values():
            span[0] = max(span[0], p["start_pos"][0] + 1, p["target_pos"][0] + 1)
            span[1] = max(span[1], p["start_pos"][1] + 1, p["target_pos"][1] + 1)
        W = int(n_cols) if n_cols else max(span[1], 2)
        prev["map"] = np.zeros((max(span[0], 2), W), np.uint8)
    b, (H, W) = _prepare(prev, persistent_packages_at_prev_state, 1, 0, 0, 1, 1, shaping=shaping,
                         min_pkg=max(int(r[2]) for r in cur_robots))
    cur = np.array([[((int(r) - 1) * W + (int(c) - 1)) | (int(k) << 16) for (r, c, k) in cur_robots]], dtype=np.uint32)
    acts = np.array([[encode_action(m, p) for (m, p) in actions_taken_for_all_agents[:num_agents]]], dtype=np.uint8)
    out = b.shaped_reward_host(cur, np.array([int(current_env_state_dict["time_step"])], np.int32), acts,
                               np.array([float(global_reward)], np.float64))
    return np.float32(out[0])
