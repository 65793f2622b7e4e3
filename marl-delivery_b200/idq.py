"""Drop-ins for the featurisers / shaping of the IDQ and legacy-qmix trainers:

  convert_state(state, persistent_packages, current_robot_idx)          IDQ/networks.py:112-217
                                                                        (== qmix/networks.py:243-348)
  convert_global_state_to_tensor(state_dict, persistent_packages, shape)  qmix/networks.py:350-468
  reward_shaping(prev_state, cur_state, actions, packages_before, n)    IDQ/networks.py:228-349

Same names, arguments and results, computed by the CUDA kernels (`mdx_featurize_alt`,
`mdx_reward_shaping_idq`).  Like helper.py, each per-call function injects its arguments into a cached
1-env scratch batch; rollouts use `DeliveryBatch.featurize_alt()` / `step_idq()`, which do all envs and
robots in one launch.
"""
from __future__ import annotations

import numpy as np
import torch

from .batch import encode_action
from .helper import _prepare


def convert_state(state, persistent_packages, current_robot_idx):
    """-> (6, n_rows, n_cols) float32."""
    b, (H, W) = _prepare(state, persistent_packages, 1, 0, 0, 1, 1)
    if not (0 <= current_robot_idx < len(state["robots"])):
        out = np.zeros((6, H, W), np.float32)
        out[0] = np.asarray(state["map"], dtype=np.float32)
        return out
    obs6, _ = b.featurize_alt(gstate7=False)
    return obs6[0, current_robot_idx].cpu().numpy()


def convert_global_state_to_tensor(state_dict, persistent_packages, state_tensor_shape):
    """-> (C, n_rows, n_cols) float32, C <= 7.  The reference can also centre / crop the map into a
    differently sized tensor (qmix/networks.py:385-407); its trainer always passes the map's own size
    (qmix/trainer.py:92), and that is what this drop-in supports."""
    C_, H_, W_ = (int(x) for x in state_tensor_shape)
    b, (H, W) = _prepare(state_dict, persistent_packages, 1, 0, 0, 1, 1)
    if (H_, W_) != (H, W) or not 0 < C_ <= 7:
        raise ValueError("state_tensor_shape must be (<= 7, n_rows, n_cols) of the map")
    _, g7 = b.featurize_alt(obs6=False)
    return g7[0, :C_].cpu().numpy()


def reward_shaping(prev_env_state, current_env_state, actions_taken, persistent_packages_before_action, num_agents):
    """-> list of per-agent floats.  Package ops that are ints take the pickup / drop branches; the
    `(move_str, op_str)` tuples the IDQ trainer passes compare unequal to 1 / 2, so only the stay penalty
    fires for them -- as in the reference."""
    prev = dict(prev_env_state)
    prev["robots"] = list(prev["robots"])[:num_agents]
    cur_robots = list(current_env_state["robots"])[:num_agents]
    cur_t = current_env_state["time_step"]
    if "time_step" not in prev:
        prev["time_step"] = cur_t - 1  # IDQ/networks.py:268
    ops = [a[1] for a in actions_taken[:num_agents]]
    as_int = all(isinstance(o, (int, np.integer)) and not isinstance(o, bool) for o in ops)
    b, (H, W) = _prepare(prev, persistent_packages_before_action, 1, 0, 0, 1, 1, min_pkg=max(int(r[2]) for r in cur_robots))
    cur = torch.tensor([[[(int(r) - 1) * W + (int(c) - 1), int(k)] for (r, c, k) in cur_robots]], dtype=torch.int64)
    acts = torch.tensor([[encode_action(m, str(int(o)) if as_int else "0") for (m, o) in actions_taken[:num_agents]]],
                        dtype=torch.uint8)
    out = b.reward_shaping_idq(cur, torch.tensor([int(cur_t)]), acts, ops_as_int=as_int)
    return [float(x) for x in out[0].cpu().tolist()]
