"""Drop-in for `QMIX/helper.py`: the same four functions as the MAPPO helper with QMIX's differences --
`generate_vector_features` has no defaults (QMIX/helper.py:70) and `compute_shaped_rewards` uses the QMIX
constants (QMIX/helper.py:270-278).  Computed by the same CUDA kernels; the constants are parameters of
the kernel, not a second code path.  QMIX/config.yaml sizes: max_robots_in_state 10, max_packages_in_state 20.
"""
from __future__ import annotations

from . import helper as _h
from .batch import QMIX_SHAPING

convert_observation = _h.convert_observation
convert_global_state = _h.convert_global_state


def generate_vector_features(env_state_dict, persistent_packages_for_env, current_robot_idx, max_time_steps,
                             max_other_robots_to_observe, max_packages_to_observe):
    """QMIX/helper.py:70-165 (no default sizes)."""
    return _h.generate_vector_features(env_state_dict, persistent_packages_for_env, current_robot_idx, max_time_steps,
                                       max_other_robots_to_observe, max_packages_to_observe)


def compute_shaped_rewards(global_reward, prev_env_state_dict, current_env_state_dict, actions_taken_for_all_agents,
                           persistent_packages_at_prev_state, num_agents, n_cols=None):
    """QMIX/helper.py:257-368."""
    return _h.compute_shaped_rewards(global_reward, prev_env_state_dict, current_env_state_dict, actions_taken_for_all_agents,
                                     persistent_packages_at_prev_state, num_agents, shaping=QMIX_SHAPING, n_cols=n_cols)
